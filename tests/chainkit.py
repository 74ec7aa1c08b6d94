"""CPU tensors + plain-torch references for the chain programs (tests/test_chain_emulator.py)."""
import torch


def _W(g, o, i, ld):
    w = torch.zeros(o, ld, dtype=torch.bfloat16)
    w[:, :i] = (torch.randn(o, i, generator=g) / i ** 0.5).to(torch.bfloat16)
    return w


def make_tensors(R, seed=0):
    """Everything ActorCritic._chain_tensors() provides, on the CPU, with random contents."""
    g = torch.Generator().manual_seed(seed)
    bf = lambda *s: torch.zeros(*s, dtype=torch.bfloat16)
    rnd = lambda r, c, s=1.0: (torch.randn(r, c, generator=g) * s).to(torch.bfloat16)
    T = {}
    T["Xp"] = bf(R, 24); T["Xp"][:, :18] = (torch.rand(R, 18, generator=g) * 2 - 1).to(torch.bfloat16)
    T["Xac"] = bf(R, 64); T["Xac"][:, :42] = rnd(R, 42)
    T["Xh"] = bf(R, 632); T["Xh"][:, :630] = rnd(R, 630)
    for n, c in (("H1", 256), ("H2", 128), ("Y1", 1024), ("A2", 256), ("A3", 128), ("C2", 256), ("C3", 128), ("D1", 256), ("D2", 32)):
        T[n] = bf(R, c)
        T["d" + n] = bf(R, c)
    T["dLat"] = bf(R, 24)
    T["dmean"] = bf(R, 16); T["dmean"][:, :12] = rnd(R, 12, 0.1)
    T["dvalue"] = bf(R, 8); T["dvalue"][:, :1] = rnd(R, 1, 0.1)
    T["dpred"] = bf(R, 24); T["dpred"][:, :18] = rnd(R, 18, 0.1)
    shapes = dict(e1=(256, 18), e2=(128, 256), e3=(18, 128), cat=(1024, 60), a2=(256, 512), a3=(128, 256), a4=(12, 128),
                  c2=(256, 512), c3=(128, 256), c4=(1, 128), d1=(256, 630), d2=(32, 256), d3=(18, 32))
    pad8 = lambda n: (n + 7) // 8 * 8
    off = 3          # deliberately not 16 B aligned
    for k, (o, i) in shapes.items():
        w = _W(g, o, i, pad8(i))
        T["W" + k] = w
        wt = torch.zeros(i, pad8(o), dtype=torch.bfloat16)
        wt[:, :o] = w[:, :i].t()
        T["W" + k + "t"] = wt
        T["b_" + k] = off
        off += o
    T["Wcat_t"] = T.pop("Wcatt")
    T["params"] = torch.randn(off, generator=g) * 0.1
    T["mean"], T["value"], T["pred"] = torch.zeros(R, 12), torch.zeros(R, 1), torch.zeros(R, 18)
    T["num_obs"] = 42
    return T


f32 = lambda t: t.float()
bfr = lambda x: x.to(torch.bfloat16).float()
elu = lambda y: torch.where(y > 0, y, torch.exp(y) - 1)
delu = lambda y: torch.where(y > 0, torch.ones_like(y), y + 1)        # ELU' from the ELU OUTPUT


ACT = {"elu": (elu, delu), "tanh": (torch.tanh, lambda y: 1 - y * y)}      # (activation, derivative from the OUTPUT)
_act = ["elu"]


def set_activation(name):
    """The hidden-layer activation of the ref_* functions below ("elu" | "tanh": the high_level_policy networks)."""
    _act[0] = name


def lin(T, x, name, act=True):
    W = f32(T["W" + name])[:, :x.shape[1]]
    y = x @ W.t() + T["params"][T["b_" + name]:T["b_" + name] + W.shape[0]]
    return ACT[_act[0]][0](y) if act else y


def ref_teacher(T):
    xp = f32(T["Xp"])[:, :18]
    h1 = bfr(lin(T, xp, "e1")); h2 = bfr(lin(T, h1, "e2"))
    lat = bfr(lin(T, h2, "e3", act=False))
    xac = f32(T["Xac"]).clone(); xac[:, 42:60] = lat
    y1 = bfr(lin(T, xac[:, :60], "cat"))
    a2 = bfr(lin(T, y1[:, :512], "a2")); a3 = bfr(lin(T, a2, "a3"))
    c2 = bfr(lin(T, y1[:, 512:], "c2")); c3 = bfr(lin(T, c2, "c3"))
    return dict(H1=h1, H2=h2, Xac=xac, Y1=y1, A2=a2, A3=a3, C2=c2, C3=c3, mean=lin(T, a3, "a4", act=False),
                value=lin(T, c3, "c4", act=False))


def ref_adaptation(T):
    d1 = bfr(lin(T, f32(T["Xh"])[:, :630], "d1")); d2 = bfr(lin(T, d1, "d2"))
    return dict(D1=d1, D2=d2, pred=lin(T, d2, "d3", act=False))


def ref_trunk_backward(T):
    """Needs the saved activations already in T (run ref_teacher and copy them in)."""
    W = lambda n: f32(T["W" + n])
    out = {}
    dy1 = torch.zeros(T["Y1"].shape[0], 1024)
    for tag, d_out, off, n_out in (("a", "dmean", 0, 12), ("c", "dvalue", 512, 1)):
        g = f32(T[d_out])[:, :n_out]
        s3, s2 = ("A3", "A2") if tag == "a" else ("C3", "C2")
        d3 = bfr((g @ W(tag + "4")[:, :128]) * ACT[_act[0]][1](f32(T[s3])))
        d2 = bfr((d3 @ W(tag + "3")[:, :256]) * ACT[_act[0]][1](f32(T[s2])))
        d1 = bfr((d2 @ W(tag + "2")[:, :512]) * ACT[_act[0]][1](f32(T["Y1"])[:, off:off + 512]))
        dy1[:, off:off + 512] = d1
        out["d" + s3], out["d" + s2] = d3, d2
    out["dY1"] = dy1
    dlat = bfr(dy1 @ W("cat")[:, 42:60])
    out["dLat"] = dlat
    dh2 = bfr((dlat @ W("e3")[:, :128]) * ACT[_act[0]][1](f32(T["H2"])))
    out["dH2"] = dh2
    out["dH1"] = bfr((dh2 @ W("e2")[:, :256]) * ACT[_act[0]][1](f32(T["H1"])))
    return out


def ref_adaptation_backward(T):
    g = f32(T["dpred"])[:, :18]
    dd2 = bfr((g @ f32(T["Wd3"])[:, :32]) * ACT[_act[0]][1](f32(T["D2"])))
    dd1 = bfr((dd2 @ f32(T["Wd2"])[:, :256]) * ACT[_act[0]][1](f32(T["D1"])))
    return dict(dD2=dd2, dD1=dd1)
