"""ActorCritic on the tcgen05 GEMM kernel.

Drop-in for mini_gym_learn/ppo/actor_critic.py:23-173: same constructor, sub-modules
(`env_factor_encoder` = `encoder`, `adaptation_module`, `actor_body`, `critic_body`), `std` parameter and
state_dict keys (so `ac_weights_*.pt` checkpoints load), `act / evaluate / get_actions_log_prob /
act_student / act_teacher / act_inference / act_expert`, `action_mean / action_std / entropy`.

What differs is underneath: all parameters live in ONE flat fp32 buffer (the nn.Parameters are views
of it) with flat gradient / Adam-moment twins, bf16 shadow copies (and transposes) feed the tensor
cores, and every Linear(+ELU) is one launch of csrc/gemm_tc.cu with the bias / ELU fused in its
epilogue.  The first layers of the actor and the critic read the same input and run as one
concatenated [1024 x 60] GEMM; the encoder is evaluated once per batch, not twice (reference quirk
10: `act` and `evaluate` both call it).  There is no eager fallback.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from .. import _lib

EPI_F32, EPI_ATOMIC, EPI_BIAS_ELU_BF16, EPI_BIAS_F32, EPI_DELU_BF16, EPI_BF16, EPI_BIAS_BF16 = range(7)


class AC_Args:
    """actor_critic.py:9-20."""
    init_noise_std = 1.0
    actor_hidden_dims = [512, 256, 128]
    critic_hidden_dims = [512, 256, 128]
    activation = "elu"
    adaptation_module_branch_hidden_dims = [[256, 32]]
    env_factor_encoder_branch_input_dims = [18]
    env_factor_encoder_branch_latent_dims = [18]
    env_factor_encoder_branch_hidden_dims = [[256, 128]]


def _pad8(n):
    return (n + 7) // 8 * 8


def _mlp(dims, activation="elu"):
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(nn.ELU() if activation == "elu" else nn.Tanh())
    return nn.Sequential(*layers)


class _Layer:
    """One Linear: fp32 master views + bf16 shadows + gradient views."""
    __slots__ = ("w", "b", "gw", "gb", "wb", "wbt", "out", "inp", "ld_wb", "ld_wbt")


class ActorCritic(nn.Module):
    is_recurrent = False

    # the configuration namespace and the latent switch of this learner family: the high_level_policy variant
    # (rapid_locomotion_rl_b200/high_level_policy: tanh networks, USE_LATENT = False) overrides both in a subclass
    ac_args = AC_Args
    use_latent = True

    def __init__(self, num_obs, num_privileged_obs, num_obs_history, num_actions, device="cuda:0", **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs.keys())))
        super().__init__()
        AC_Args = self.ac_args          # noqa: N806 (shadows the module-level namespace on purpose)
        self.activation = AC_Args.activation
        if self.activation not in ("elu", "tanh"):
            raise NotImplementedError("the fused epilogues cover ELU and tanh (AC_Args.activation=%r)" % AC_Args.activation)
        if num_actions != 12 or num_privileged_obs != 18 or AC_Args.env_factor_encoder_branch_latent_dims != [18]:
            raise NotImplementedError("the fused loss kernel is specialised for 12 actions and an 18-d latent")
        if len(AC_Args.actor_hidden_dims) != 3 or len(AC_Args.critic_hidden_dims) != 3 or \
                AC_Args.actor_hidden_dims[0] != AC_Args.critic_hidden_dims[0]:
            raise NotImplementedError("actor/critic must have 3 hidden layers with equal first widths")
        self.device_ = torch.device(device)
        if self.device_.type != "cuda":
            raise _lib.RlError("ActorCritic needs a CUDA device: the learner has no CPU fallback")
        self._lib = _lib.lib()
        self.num_obs, self.num_priv, self.num_hist, self.num_actions = num_obs, num_privileged_obs, num_obs_history, num_actions
        lat = AC_Args.env_factor_encoder_branch_latent_dims[0]
        self.latent_dim = lat
        eh = AC_Args.env_factor_encoder_branch_hidden_dims[0]
        ah = AC_Args.adaptation_module_branch_hidden_dims[0]
        act_fn = self.activation
        self.env_factor_encoder = _mlp([AC_Args.env_factor_encoder_branch_input_dims[0]] + eh + [lat], act_fn)
        self.add_module("encoder", self.env_factor_encoder)           # alias, as in the reference (:56)
        self.adaptation_module = _mlp([num_obs_history] + ah + [lat], act_fn)
        self.actor_body = _mlp([lat + num_obs] + AC_Args.actor_hidden_dims + [num_actions], act_fn)
        self.critic_body = _mlp([lat + num_obs] + AC_Args.critic_hidden_dims + [1], act_fn)
        self.std = nn.Parameter(AC_Args.init_noise_std * torch.ones(num_actions))
        self.distribution = None
        if not self.use_latent:
            # high_level_policy with USE_LATENT = False (high_level_policy/ppo/actor_critic.py:39, :146-150, :188-192): the
            # bodies see the observations only.  The fused passes keep their [obs | latent] input box; the latent is pinned to
            # exactly zero instead - the encoder's last layer and the latent columns of both first layers start at zero, and
            # zero they stay: the latent columns' weight gradient is dY^T * latent = 0, the encoder's gradient comes through
            # those zero columns, and Adam leaves a parameter whose gradient was always zero untouched.  The observation
            # columns are drawn like nn.Linear(num_obs, .) draws them (bound 1 / sqrt(num_obs)).
            with torch.no_grad():
                last = [m for m in self.env_factor_encoder if isinstance(m, nn.Linear)][-1]
                last.weight.zero_(); last.bias.zero_()
                bound = 1.0 / float(num_obs) ** 0.5
                for body in (self.actor_body, self.critic_body):
                    body[0].weight[:, :num_obs].uniform_(-bound, bound)
                    body[0].bias.uniform_(-bound, bound)
                    body[0].weight[:, num_obs:].zero_()
        self.to(self.device_)
        self._flatten()
        self._ws_rows = 0
        self._ws_gen = 0          # bumped whenever the workspace is reallocated: captured graphs hold its addresses
        self._cache_key = None
        self._split = None        # (stream, fork event, join event) of forward_teacher's two-launch form
        self._act_step = 0
        self.seed = 0
        # fused MLP chains (csrc/chain.cu): one persistent kernel per network pass instead of one GEMM per layer
        self.use_chain = os.environ.get("RL_USE_CHAIN", "1") != "0"
        self._chains = {}

    # ------------------------------------------------------------------------------------------------
    def _linears(self, seq):
        return [m for m in seq if isinstance(m, nn.Linear)]

    def _flatten(self):
        """Move every parameter into one flat fp32 buffer: [encoder | actor+critic | std | adaptation].
        Actor / critic first-layer weights (and biases) are adjacent so they form one [1024, 60] layer."""
        enc, ada = self._linears(self.env_factor_encoder), self._linears(self.adaptation_module)
        act, cri = self._linears(self.actor_body), self._linears(self.critic_body)
        order = []
        for l in enc:
            order += [l.weight, l.bias]
        order += [act[0].weight, cri[0].weight, act[0].bias, cri[0].bias]
        for la, lc in zip(act[1:], cri[1:]):
            order += [la.weight, la.bias, lc.weight, lc.bias]
        order += [self.std]
        n_main = sum(p.numel() for p in order)
        for l in ada:
            order += [l.weight, l.bias]
        n = sum(p.numel() for p in order)
        dev = self.device_
        self.flat = torch.zeros(n, device=dev)
        # + 8 tail floats: the loss statistics ride in the same all-reduce as the gradients (multi-GPU)
        self._grad_store = torch.zeros(n + 8, device=dev)
        self.flat_grad = self._grad_store[:n]
        self.flat_m = torch.zeros(n, device=dev)
        self.flat_v = torch.zeros(n, device=dev)
        self.n_main, self.n_total = n_main, n
        off = 0
        self._grad_view = {}
        for p in order:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            self._grad_view[id(p)] = self.flat_grad[off:off + k].view(p.shape)
            off += k
        # layer table
        H = AC_Args.actor_hidden_dims[0]

        def mk(w, b, gw, gb):
            L = _Layer()
            L.w, L.b, L.gw, L.gb = w, b, gw, gb
            L.out, L.inp = w.shape
            L.ld_wb, L.ld_wbt = _pad8(L.inp), _pad8(L.out)
            L.wb = torch.zeros(L.out, L.ld_wb, dtype=torch.bfloat16, device=dev)
            L.wbt = torch.zeros(L.inp, L.ld_wbt, dtype=torch.bfloat16, device=dev)
            return L
        g = lambda p: self._grad_view[id(p)]
        self.L_enc = [mk(l.weight.data, l.bias.data, g(l.weight), g(l.bias)) for l in enc]
        self.L_ada = [mk(l.weight.data, l.bias.data, g(l.weight), g(l.bias)) for l in ada]
        self.L_act = [mk(l.weight.data, l.bias.data, g(l.weight), g(l.bias)) for l in act[1:]]
        self.L_cri = [mk(l.weight.data, l.bias.data, g(l.weight), g(l.bias)) for l in cri[1:]]
        # concatenated first layer: rows [0,H) actor, [H,2H) critic
        w_off = self._offset_of(act[0].weight)
        b_off = self._offset_of(act[0].bias)
        kin = act[0].weight.shape[1]
        self.L_cat = mk(self.flat[w_off:w_off + 2 * H * kin].view(2 * H, kin), self.flat[b_off:b_off + 2 * H],
                        self.flat_grad[w_off:w_off + 2 * H * kin].view(2 * H, kin), self.flat_grad[b_off:b_off + 2 * H])
        self.std_grad = g(self.std)
        self._all_layers = self.L_enc + [self.L_cat] + self.L_act + self.L_cri + self.L_ada
        self.refresh_shadows()

    def rebind_grad_store(self, store):
        """Moves the flat gradient (and every view of it: per-parameter, per-layer, std) into `store`, a float32
        device tensor with at least n_total + 8 elements - e.g. memory that the other ranks have mapped."""
        old = self.flat_grad
        n = self.n_total
        assert store.dtype == torch.float32 and store.numel() >= n + 8 and store.is_contiguous()
        store[:n + 8].copy_(self._grad_store)
        new_flat = store[:n]

        def move(v):
            off = (v.data_ptr() - old.data_ptr()) // 4
            assert 0 <= off and off + v.numel() <= n
            return new_flat[off:off + v.numel()].view(v.shape)
        self._grad_view = {k: move(v) for k, v in self._grad_view.items()}
        for L in self._all_layers:
            L.gw, L.gb = move(L.gw), move(L.gb)
        self.std_grad = move(self.std_grad)
        self._grad_store, self.flat_grad = store[:n + 8], new_flat

    def _offset_of(self, p):
        return (p.data.data_ptr() - self.flat.data_ptr()) // 4

    def refresh_shadows(self, layers=None):
        """bf16 copies (and transposes) of the master weights: the tensor-core operands."""
        Ls = self._all_layers if layers is None else layers
        n = len(Ls)
        arr_p = (C.c_void_p * n)
        arr_i = (C.c_int32 * n)
        _lib.check(self._lib.rl_refresh_shadows(
            arr_p(*[L.w.data_ptr() for L in Ls]), arr_p(*[L.wb.data_ptr() for L in Ls]),
            arr_p(*[L.wbt.data_ptr() for L in Ls]), arr_i(*[L.out for L in Ls]), arr_i(*[L.inp for L in Ls]),
            arr_i(*[L.ld_wb for L in Ls]), arr_i(*[L.ld_wbt for L in Ls]), n, _lib.current_stream()))
        self._cache_key = None

    def adam_shadows(self, start, n, grad, ctrl, lr_fixed, use_ctrl, step_dev, layers, grad_ptr=None):
        """One launch: Adam on flat[start : start + n] (rl_adam's arguments) and the bf16 operands of `layers`, whose
        master weights lie inside that range (csrc/ppo.cu adam_shadows_kernel).  grad_ptr: address of the range's
        gradient when it does not live in `grad` (the flat gradient buffer) at the same offset."""
        key = (start, tuple(id(L) for L in layers))
        tab = self._adam_tabs.get(key) if hasattr(self, "_adam_tabs") else None
        if tab is None:
            if not hasattr(self, "_adam_tabs"):
                self._adam_tabs = {}
            k = len(layers)
            arr_p, arr_i, arr_l = (C.c_void_p * k), (C.c_int32 * k), (C.c_int64 * k)
            tab = (arr_l(*[self._offset_of_data(L.w) - start for L in layers]), arr_p(*[L.wb.data_ptr() for L in layers]),
                   arr_p(*[L.wbt.data_ptr() for L in layers]), arr_i(*[L.out for L in layers]), arr_i(*[L.inp for L in layers]),
                   arr_i(*[L.ld_wb for L in layers]), arr_i(*[L.ld_wbt for L in layers]), k)
            self._adam_tabs[key] = tab
        off = 4 * start
        _lib.check(self._lib.rl_adam_shadows(
            self.flat.data_ptr() + off, (grad.data_ptr() + off) if grad_ptr is None else grad_ptr,
            self.flat_m.data_ptr() + off, self.flat_v.data_ptr() + off, n,
            ctrl, float(lr_fixed), int(use_ctrl), 0.9, 0.999, 1e-8, 0, 1.0, step_dev, *tab, _lib.current_stream()))
        self._cache_key = None

    def _offset_of_data(self, t):
        return (t.data_ptr() - self.flat.data_ptr()) // 4

    def split_adaptation_grad(self):
        """Moves the adaptation module's gradient out of the flat buffer into one of its own (returned; same internal
        layout as flat_grad[n_main : n_total]): with several GPUs it is reduced by a collective of its own, on the side
        branch, while the policy gradient's all-reduce runs on the main path (PPO.minibatch_step)."""
        if getattr(self, "ada_grad", None) is not None:
            return self.ada_grad
        n_ad = self.n_total - self.n_main
        g = torch.zeros(n_ad, device=self.device_)
        base = self.flat_grad.data_ptr() + 4 * self.n_main

        def move(v):
            off = (v.data_ptr() - base) // 4
            assert 0 <= off and off + v.numel() <= n_ad
            return g[off:off + v.numel()].view(v.shape)
        g.copy_(self.flat_grad[self.n_main:self.n_total])
        self.flat_grad[self.n_main:self.n_total].zero_()
        for L in self.L_ada:
            L.gw, L.gb = move(L.gw), move(L.gb)
        for k, v in list(self._grad_view.items()):
            if base <= v.data_ptr() < base + 4 * n_ad:
                self._grad_view[k] = move(v)
        self.ada_grad = g
        return g

    def load_state_dict(self, state_dict, strict=True):
        if not self.use_latent:
            # the reference's no-latent class has `actor_body` / `critic_body` / `std` only, first layers [., num_obs]
            # (high_level_policy/ppo/actor_critic.py:86-110): pad them with the (zero) latent columns
            sd = dict(super().state_dict())
            for k, v in state_dict.items():
                if k in ("actor_body.0.weight", "critic_body.0.weight") and v.shape[1] == self.num_obs:
                    full = torch.zeros_like(sd[k])
                    full[:, :self.num_obs] = v
                    v = full
                sd[k] = v
            state_dict = sd
        out = super().load_state_dict(state_dict, strict=strict)
        self.refresh_shadows()
        return out

    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        if not self.use_latent:
            keep = type(sd)()
            for k, v in sd.items():
                if k.startswith(("actor_body.", "critic_body.")) or k == "std":
                    keep[k] = v[:, :self.num_obs] if k in ("actor_body.0.weight", "critic_body.0.weight") else v
            if hasattr(sd, "_metadata"):
                keep._metadata = sd._metadata
            return keep
        return sd

    # ------------------------------------------------------------------------------------------------
    # workspace
    # ------------------------------------------------------------------------------------------------
    def workspace(self, rows, backward=False):
        """bf16 activation (and gradient) buffers for `rows` rows, grown on demand."""
        need_new = rows > self._ws_rows or (backward and not getattr(self, "_ws_bwd", False))
        if need_new:
            R, dev = max(rows, self._ws_rows), self.device_
            bf = lambda c: torch.zeros(R, c, dtype=torch.bfloat16, device=dev)
            f32 = lambda *s: torch.zeros(R, *s, device=dev)
            H = AC_Args.actor_hidden_dims
            eh = AC_Args.env_factor_encoder_branch_hidden_dims[0]
            ah = AC_Args.adaptation_module_branch_hidden_dims[0]
            w = {}
            w["Xp"], w["Xac"], w["Xh"] = bf(_pad8(self.num_priv)), bf(_pad8(self.num_obs + self.latent_dim)), bf(_pad8(self.num_hist))
            w["H1"], w["H2"] = bf(eh[0]), bf(eh[1])
            w["Y1"] = bf(2 * H[0])
            w["A2"], w["A3"], w["C2"], w["C3"] = bf(H[1]), bf(H[2]), bf(H[1]), bf(H[2])
            w["D1"], w["D2"] = bf(ah[0]), bf(ah[1])
            w["mean"], w["value"], w["pred"] = f32(self.num_actions), f32(1), f32(self.latent_dim)
            if backward or getattr(self, "_ws_bwd", False):
                w["Lrow"] = f32(40)
                w["dmean"], w["dvalue"], w["dpred"], w["dLat"] = bf(16), bf(8), bf(24), bf(24)
                w["dA3"], w["dA2"], w["dC3"], w["dC2"], w["dY1"] = bf(H[2]), bf(H[1]), bf(H[2]), bf(H[1]), bf(2 * H[0])
                w["dH2"], w["dH1"], w["dD2"], w["dD1"] = bf(eh[1]), bf(eh[0]), bf(ah[1]), bf(ah[0])
                self._ws_bwd = True
            self._ws, self._ws_rows = w, R
            self._ws_gen += 1
            self._cache_key = None
            self._chains = {}          # chain programs bake the workspace pointers in
        return self._ws

    # ------------------------------------------------------------------------------------------------
    # fused chains
    # ------------------------------------------------------------------------------------------------
    def _chain_tensors(self):
        """Workspace buffers, bf16 shadow weights and bias offsets (floats into self.flat) by the names
        ppo/chain.py uses."""
        w = self._ws
        off = lambda L: (L.b.data_ptr() - self.flat.data_ptr()) // 4
        e, a, c, d = self.L_enc, self.L_act, self.L_cri, self.L_ada
        T = {k: w[k] for k in w}
        T.update(params=self.flat, num_obs=self.num_obs,
                 We1=e[0].wb, We2=e[1].wb, We3=e[2].wb, Wcat=self.L_cat.wb, Wa2=a[0].wb, Wa3=a[1].wb, Wa4=a[2].wb,
                 Wc2=c[0].wb, Wc3=c[1].wb, Wc4=c[2].wb, Wd1=d[0].wb, Wd2=d[1].wb, Wd3=d[2].wb,
                 We2t=e[1].wbt, We3t=e[2].wbt, Wcat_t=self.L_cat.wbt, Wa2t=a[0].wbt, Wa3t=a[1].wbt, Wa4t=a[2].wbt,
                 Wc2t=c[0].wbt, Wc3t=c[1].wbt, Wc4t=c[2].wbt, Wd2t=d[1].wbt, Wd3t=d[2].wbt,
                 b_e1=off(e[0]), b_e2=off(e[1]), b_e3=off(e[2]), b_cat=off(self.L_cat), b_a2=off(a[0]), b_a3=off(a[1]),
                 b_a4=off(a[2]), b_c2=off(c[0]), b_c3=off(c[1]), b_c4=off(c[2]), b_d1=off(d[0]), b_d2=off(d[1]),
                 b_d3=off(d[2]))
        return T

    def prepare_update_chains(self):
        """Compiles every chain PPO.minibatch_step launches (rl_chain_create allocates and copies, which is
        not allowed while a CUDA graph is being captured)."""
        if not self.use_chain:
            return
        from . import chain
        self._chain(("teacher", True, True, True), lambda T: chain.teacher_forward(T, save=True))
        self._chain(("trunk_backward",), chain.trunk_backward)
        self._chain(("encoder",), lambda T: chain.teacher_forward(T, save=False, trunk=False))
        self._chain(("adaptation", True), lambda T: chain.adaptation_forward_program(T, save=True))
        self._chain(("adaptation_backward",), chain.adaptation_backward_program)

    def prepare_rollout_chains(self, rows):
        """Workspace + policy chain(s) for `act` / `evaluate` on `rows` envs, compiled ahead of a graph capture."""
        self.workspace(rows)
        if self.use_chain:
            from . import chain
            self._chain(("teacher", False, True, True), lambda T: chain.teacher_forward(T, save=False))
            if self._split_ok(rows):
                for wm, wv in ((True, False), (False, True)):
                    self._chain(("teacher", False, wm, wv),
                                lambda T, wm=wm, wv=wv: chain.teacher_forward(T, save=False, want_mean=wm, want_value=wv))
                self._split_streams()

    @staticmethod
    def _split_ok(rows):
        """Small batches leave most SMs idle (one 128-row tile per CTA, and a tile's pass is a serial chain of layers:
        ~40 us whatever the batch): the actor and the critic then run as TWO concurrent launches, each with its own
        encoder pass, on disjoint SMs - 23 us instead of 40 us for 4000 envs.  Worth it while both fit in one wave."""
        return os.environ.get("RL_CHAIN_SPLIT", "1") != "0" and 2 * ((rows + 127) // 128) <= 148

    def _split_streams(self):
        if self._split is None:
            self._split = (torch.cuda.Stream(device=self.device_), torch.cuda.Event(), torch.cuda.Event())
        return self._split

    def _chain(self, key, build):
        prog = self._chains.get(key)
        if prog is None:
            prog = build(self._chain_tensors())
            prog.activation = self.activation          # "tanh": hidden-layer epilogues packed as the tanh modes
            prog = prog.compile()
            self._chains[key] = prog
        return prog

    # ------------------------------------------------------------------------------------------------
    # GEMM plumbing
    # ------------------------------------------------------------------------------------------------
    def _gemm(self, A, lda, B, ldb, Cp, ldc, M, N, K, epi, transposed=0, bias=None, aux=None, ld_aux=0, db=None, split_k=1):
        if self.activation != "elu" and epi in (EPI_BIAS_ELU_BF16, EPI_DELU_BF16):
            raise NotImplementedError("the per-layer path (RL_USE_CHAIN=0) fuses ELU only; tanh networks run on the chain kernels")
        _lib.check(self._lib.rl_gemm_bf16(A, B, Cp, bias, aux, db, M, N, K, lda, ldb, ldc, ld_aux, transposed, epi,
                                          split_k, _lib.current_stream()))

    @staticmethod
    def _p(t, elem_off=0):
        return t.data_ptr() + elem_off * t.element_size()

    def _fwd(self, L, X, x_off, ldx, Y, y_off, ldy, rows, epi):
        self._gemm(self._p(X, x_off), ldx, L.wb.data_ptr(), L.ld_wb, self._p(Y, y_off), ldy, rows, L.out, L.inp, epi,
                   bias=L.b.data_ptr())

    def forward_encoder(self, rows):
        """env_factor_encoder(priv) -> bf16 latent written straight into the latent slot of Xac."""
        if self.use_chain:
            from . import chain
            self._chain(("encoder",), lambda T: chain.teacher_forward(T, save=False, trunk=False)).run(rows)
            return
        w, e = self._ws, self.L_enc
        self._fwd(e[0], w["Xp"], 0, w["Xp"].shape[1], w["H1"], 0, w["H1"].shape[1], rows, EPI_BIAS_ELU_BF16)
        self._fwd(e[1], w["H1"], 0, w["H1"].shape[1], w["H2"], 0, w["H2"].shape[1], rows, EPI_BIAS_ELU_BF16)
        self._fwd(e[2], w["H2"], 0, w["H2"].shape[1], w["Xac"], self.num_obs, w["Xac"].shape[1], rows, EPI_BIAS_BF16)

    def forward_teacher(self, rows, want_value=True, want_mean=True, save=False, tiles=None, loss=None):
        """encoder -> [actor | critic] on the staged inputs Xp / Xac of the workspace (tiles: only these 128-row tiles).
        loss (chain path, save=True): an `_lib.RlChainPpoLoss` - the critic's output epilogue also evaluates the PPO loss of its
        rows (dmean / dvalue / dstd / statistics as rl_ppo_loss writes them), so no loss launch follows."""
        if self.use_chain:
            from . import chain
            prog = lambda wm, wv: self._chain(("teacher", save, wm, wv), lambda T: chain.teacher_forward(
                T, save=save, want_mean=wm, want_value=wv))
            if loss is not None or (save and want_mean and want_value):
                # (the setting is part of the launch parameters: always (re)stated for the program the update uses)
                assert loss is None or (save and want_mean and want_value)
                prog(True, True).set_ppo_loss(loss)
            if tiles is not None:
                prog(want_mean, want_value).run(rows, tiles=tiles)
                return
            if want_mean and want_value and not save and self._split_ok(rows):
                side, fork, join = self._split_streams()
                critic, actor = prog(False, True), prog(True, False)      # (compiled before the fork: graph-capture safe)
                fork.record()
                with torch.cuda.stream(side):
                    side.wait_event(fork)
                    critic.run(rows)
                    join.record()
                actor.run(rows)
                torch.cuda.current_stream().wait_event(join)
                return
            prog(want_mean, want_value).run(rows)
            return
        self.forward_encoder(rows)
        self._trunk(rows, want_value, want_mean)

    def _trunk(self, rows, want_value=True, want_mean=True):
        w = self._ws
        H = AC_Args.actor_hidden_dims[0]
        ldy = w["Y1"].shape[1]
        self._fwd(self.L_cat, w["Xac"], 0, w["Xac"].shape[1], w["Y1"], 0, ldy, rows, EPI_BIAS_ELU_BF16)
        if want_mean:
            a = self.L_act
            self._fwd(a[0], w["Y1"], 0, ldy, w["A2"], 0, w["A2"].shape[1], rows, EPI_BIAS_ELU_BF16)
            self._fwd(a[1], w["A2"], 0, w["A2"].shape[1], w["A3"], 0, w["A3"].shape[1], rows, EPI_BIAS_ELU_BF16)
            self._fwd(a[2], w["A3"], 0, w["A3"].shape[1], w["mean"], 0, self.num_actions, rows, EPI_BIAS_F32)
        if want_value:
            c = self.L_cri
            self._fwd(c[0], w["Y1"], H, ldy, w["C2"], 0, w["C2"].shape[1], rows, EPI_BIAS_ELU_BF16)
            self._fwd(c[1], w["C2"], 0, w["C2"].shape[1], w["C3"], 0, w["C3"].shape[1], rows, EPI_BIAS_ELU_BF16)
            self._fwd(c[2], w["C3"], 0, w["C3"].shape[1], w["value"], 0, 1, rows, EPI_BIAS_F32)

    def forward_adaptation(self, rows, into_latent_slot=False, save=False):
        """adaptation_module(obs_history) -> pred (fp32) or, for the student policy, the latent slot of Xac."""
        if self.use_chain and not into_latent_slot:
            from . import chain
            self._chain(("adaptation", save), lambda T: chain.adaptation_forward_program(T, save=save)).run(rows)
            return
        w, a = self._ws, self.L_ada
        self._fwd(a[0], w["Xh"], 0, w["Xh"].shape[1], w["D1"], 0, w["D1"].shape[1], rows, EPI_BIAS_ELU_BF16)
        self._fwd(a[1], w["D1"], 0, w["D1"].shape[1], w["D2"], 0, w["D2"].shape[1], rows, EPI_BIAS_ELU_BF16)
        if into_latent_slot:
            self._fwd(a[2], w["D2"], 0, w["D2"].shape[1], w["Xac"], self.num_obs, w["Xac"].shape[1], rows, EPI_BIAS_BF16)
        else:
            self._fwd(a[2], w["D2"], 0, w["D2"].shape[1], w["pred"], 0, self.latent_dim, rows, EPI_BIAS_F32)

    def _stage(self, src, dst, cols, dst_col0=0, pad_to=None):
        src = src.to(self.device_, torch.float)
        if src.stride(-1) != 1:
            src = src.contiguous()
        rows = src.shape[0]
        pad_to = cols if pad_to is None else pad_to
        _lib.check(self._lib.rl_cast_bf16(src.data_ptr(), src.stride(0), dst.data_ptr(), dst.shape[1], rows, cols, dst_col0,
                                          pad_to, _lib.current_stream()))

    def _stage_obs(self, observations, rows):
        w = self._ws
        # obs into columns [0, num_obs); the tail pad columns beyond the latent slot stay zero
        self._stage(observations, w["Xac"], self.num_obs)

    # ------------------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------------------
    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    @property
    def action_mean(self):
        return self._mean

    @property
    def action_std(self):
        return self._mean * 0.0 + self.std

    @property
    def entropy(self):
        per_dim = 0.5 + 0.5 * torch.log(torch.tensor(2 * torch.pi, device=self.device_)) + torch.log(self.std)
        return per_dim.sum().expand(self._mean.shape[0])

    def update_distribution(self, observations, privileged_observations, reuse=False):
        """encoder + actor + critic in one pass.  `reuse=True` (only `evaluate` right after `act` on the
        same tensors, the PPO.act pattern ppo.py:64-65) consumes the cached pass instead of repeating it;
        the cache is single-use because env buffers are rewritten in place by kernels torch cannot see."""
        key = (observations.data_ptr(), 0 if privileged_observations is None else privileged_observations.data_ptr(),
               observations.shape[0])
        if reuse and key == self._cache_key:
            self._cache_key = None
            return
        rows = observations.shape[0]
        w = self.workspace(rows)
        self._stage_obs(observations, rows)
        if privileged_observations is not None and self.use_latent:
            self._stage(privileged_observations, w["Xp"], self.num_priv, 0, w["Xp"].shape[1])
        # (no-latent family: the encoder's input box stays at its zeros and its output is pinned to zero anyway)
        self.forward_teacher(rows)
        self._mean = w["mean"][:rows]
        self._value = w["value"][:rows]
        self._cache_key = key

    def act(self, observations, privileged_observations, inject_normal=None, **kwargs):
        """actor_critic.py:142-144: sample from Normal(mean, std)."""
        self.update_distribution(observations, privileged_observations)
        rows = observations.shape[0]
        self._actions = torch.empty(rows, self.num_actions, device=self.device_)
        self._logp = torch.empty(rows, device=self.device_)
        self._mu = torch.empty(rows, self.num_actions, device=self.device_)
        self._sigma = torch.empty(rows, self.num_actions, device=self.device_)
        self._act_step += 1
        _lib.check(self._lib.rl_policy_sample(self._mean.data_ptr(), self.std.data_ptr(), rows, self.seed, self._act_step,
                                              _lib.ptr(inject_normal), self._actions.data_ptr(), self._logp.data_ptr(),
                                              self._mu.data_ptr(), self._sigma.data_ptr(), _lib.current_stream()))
        return self._actions

    def get_actions_log_prob(self, actions):
        if actions is self._actions or (actions.data_ptr() == self._actions.data_ptr() and actions.shape == self._actions.shape):
            return self._logp
        var = self.std ** 2
        return (-((actions - self._mean) ** 2) / (2 * var) - torch.log(self.std) - 0.9189385332046727).sum(dim=-1)

    def evaluate(self, critic_observations, privileged_observations, **kwargs):
        self.update_distribution(critic_observations, privileged_observations, reuse=True)
        return self._value.clone()

    def act_teacher(self, observations, privileged_info, policy_info={}):
        self.update_distribution(observations, privileged_info)
        return self._mean.clone()

    def act_student(self, observations, observation_history, policy_info={}):
        if not self.use_latent:        # high_level_policy/ppo/actor_critic.py:175: actor_body(observations)
            self.update_distribution(observations, None)
            return self._mean.clone()
        rows = observations.shape[0]
        w = self.workspace(rows)
        self._stage_obs(observations, rows)
        self._stage(observation_history, w["Xh"], self.num_hist, 0, w["Xh"].shape[1])
        self.forward_adaptation(rows, into_latent_slot=True)
        self._trunk(rows, want_value=False)
        self._cache_key = None
        return w["mean"][:rows].clone()

    def act_expert(self, ob, policy_info={}):
        return self.act_teacher(ob["obs"], ob["privileged_obs"])

    def act_inference(self, ob, policy_info={}):
        return self.act_student(ob["obs"], ob["obs_history"])
