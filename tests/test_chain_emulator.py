"""Chain programs (rapid_locomotion_rl_b200/ppo/chain.py) on the CPU emulator: every schedule must run to
completion under adversarial scheduling without deadlock, phase-parity aliasing or shared-memory hazards,
and reproduce the plain-torch network maths (bf16 activations, fp32 accumulation).  Also checks that the
emulator really detects broken schedules (mutation tests)."""
import pytest
import torch

import chainkit as ck
from rapid_locomotion_rl_b200.ppo import chain

ROWS = 300          # 3 tiles, the last one partial


def _close(T, ref, keys, tol=0.03):
    for k in keys:
        got, want = T[k].float()[:, :ref[k].shape[1]], ref[k]
        err = (got - want).abs().max().item()
        assert err <= tol * max(1.0, want.abs().max().item()), (k, err)


@pytest.mark.parametrize("seed,n_ctas", [(0, 1), (1, 2)])
def test_teacher_forward(seed, n_ctas):
    T = ck.make_tensors(ROWS, seed)
    ref = ck.ref_teacher(T)
    prog = chain.teacher_forward_program(T)
    chain.Emulator(prog, ROWS, n_ctas=n_ctas, seed=seed).run()
    _close(T, ref, ("H1", "H2", "Xac", "Y1", "A2", "A3", "C2", "C3", "mean", "value"))


def test_teacher_forward_variants():
    for kw in (dict(save=False), dict(trunk=False), dict(want_value=False), dict(want_mean=False, save=False)):
        T = ck.make_tensors(ROWS, 3)
        ref = ck.ref_teacher(T)
        chain.Emulator(chain.teacher_forward_program(T, **kw), ROWS, seed=4).run()
        keys = ["Xac"]
        if kw.get("trunk", True):
            keys += (["mean"] if kw.get("want_mean", True) else []) + (["value"] if kw.get("want_value", True) else [])
        _close(T, ref, keys)
        if not kw.get("save", True):
            assert float(T["H1"].float().abs().max()) == 0.0        # nothing stored


@pytest.mark.parametrize("seed,n_ctas", [(0, 1), (5, 2)])
def test_trunk_backward(seed, n_ctas):
    T = ck.make_tensors(ROWS, seed)
    fwd = ck.ref_teacher(T)
    for k in ("H1", "H2", "Y1", "A2", "A3", "C2", "C3"):
        T[k].copy_(fwd[k].to(torch.bfloat16))
    ref = ck.ref_trunk_backward(T)
    prog = chain.trunk_backward_program(T)
    chain.Emulator(prog, ROWS, n_ctas=n_ctas, seed=seed).run()
    _close(T, ref, ("dA3", "dA2", "dC3", "dC2", "dY1", "dLat", "dH2", "dH1"))


def test_adaptation_forward_and_backward():
    T = ck.make_tensors(ROWS, 7)
    ref = ck.ref_adaptation(T)
    chain.Emulator(chain.adaptation_forward_program(T), ROWS, seed=1).run()
    _close(T, ref, ("D1", "D2", "pred"))
    refb = ck.ref_adaptation_backward(T)
    chain.Emulator(chain.adaptation_backward_program(T), ROWS, n_ctas=2, seed=2).run()
    _close(T, refb, ("dD2", "dD1"))


def _mutants():
    def no_acc_free(p):
        for o in p.mmas:
            o["waits"] = [w for w in o["waits"] if ".free" not in p.bar_name[w.bar] or "acc" not in p.bar_name[w.bar]]

    def no_stage_empty(p):
        for o in p.loads:
            if "stage" in p.bar_name[o["wait"].bar]:
                o["wait"] = None

    def no_box_ready(p):
        for o in p.mmas:
            o["waits"] = [w for w in o["waits"] if "ready" not in p.bar_name[w.bar]]

    def no_store_wait(p):
        for o in p.epis:
            o["store_wait_pending"] = -1
    return [no_acc_free, no_stage_empty, no_box_ready, no_store_wait]


@pytest.mark.parametrize("mutant", _mutants(), ids=lambda m: m.__name__)
def test_emulator_detects_broken_schedules(mutant):
    caught = 0
    for seed in range(4):
        T = ck.make_tensors(ROWS, seed)
        prog = chain.teacher_forward_program(T)
        mutant(prog)
        try:
            chain.Emulator(prog, ROWS, seed=seed).run()
            ref = ck.ref_teacher(T)
            caught += any((T[k].float() - ref[k]).abs().max().item() > 0.05 for k in ("mean", "value", "Y1", "A2"))
        except chain.ChainHazard:
            caught += 1
    assert caught >= 1


def test_emulator_detects_accumulator_aliasing():
    """trunk_backward lets two accumulator names share tensor-memory columns (SB inside D2) and relies on program
    order for safety.  Swapping the ping-pong order makes super-chunk 0 overwrite columns of the 256-wide
    gradient the epilogue has not read yet: the emulated result must come out wrong."""
    import inspect
    code = inspect.getsource(chain.trunk_backward_program)
    assert '"SA" if k % 2 == 0 else "SB"' in code
    ns = dict(vars(chain))
    exec(code.replace('"SA" if k % 2 == 0 else "SB"', '"SB" if k % 2 == 0 else "SA"'), ns)
    caught = 0
    for seed in range(6):
        T = ck.make_tensors(ROWS, seed)
        fwd = ck.ref_teacher(T)
        for k in ("H1", "H2", "Y1", "A2", "A3", "C2", "C3"):
            T[k].copy_(fwd[k].to(torch.bfloat16))
        ref = ck.ref_trunk_backward(T)
        try:
            chain.Emulator(ns["trunk_backward_program"](T), ROWS, seed=seed).run()
            caught += any((T[k].float()[:, :ref[k].shape[1]] - ref[k]).abs().max().item() > 0.03 * max(1.0, ref[k].abs().max().item())
                          for k in ("dA2", "dC2", "dY1", "dLat"))
        except chain.ChainHazard:
            caught += 1
    assert caught >= 1


def test_packed_program_layout():
    """The ctypes op arrays carry what the builder decided (spot checks) and respect the device limits."""
    T = ck.make_tensors(ROWS, 0)
    prog = chain.teacher_forward_program(T)
    d = prog.pack()
    L, M, E, _ = prog._packed
    assert d.n_units <= 14 and d.n_barriers <= 64 and d.n_tensors <= 32
    assert d.n_loads == len(prog.loads) and d.n_mmas == len(prog.mmas) and d.n_epis == len(prog.epis)
    for i, o in enumerate(prog.mmas):
        assert M[i].n == o["n"] and M[i].tmem_col == o["tmem_col"] and M[i].tmem_col + M[i].n <= 512
        assert M[i].a_off % 1024 == 0 and M[i].b_off % 1024 == 0
    for i, o in enumerate(prog.loads):
        assert L[i].expect_bytes == o["bytes"] and L[i].smem_off == o["smem_off"]
    assert sum(1 for o in prog.epis if o["store_tensor"] != chain.NONE) == sum(prog.n_stores)


@pytest.mark.parametrize("a_depth,n_ctas", [(2, 1), (4, 2), (6, 1)])
def test_adaptation_forward_ring_depths(a_depth, n_ctas):
    """Two TMA rings (input boxes `a_depth` deep, weights in what is left): every split of the ring units must run
    to completion and give the same numbers - the prefetch distance is a performance knob, not a correctness one."""
    T = ck.make_tensors(ROWS, 11)
    ref = ck.ref_adaptation(T)
    prog = chain.adaptation_forward_program(T, a_depth=a_depth)
    assert set(prog.rings) == {"w", "x"} and prog.rings["x"]["n"] == a_depth and prog.n_units <= 14
    chain.Emulator(prog, ROWS, n_ctas=n_ctas, seed=a_depth).run()
    _close(T, ref, ("D1", "D2", "pred"))


def test_emulator_detects_load_order_deadlock():
    """The LOAD role issues in program order.  A load that waits for a ring stage whose release depends on a LATER
    load can never be issued: the emulator must report the deadlock instead of spinning.  Built by emitting one
    more input box ahead than the input ring has stages."""
    T = ck.make_tensors(ROWS, 0)
    p = chain.ChainProgram(n_pool=2, n_stages=0, n_inputs=0, regions={"BIG": (0, 256)}, name="bad_prefetch",
                           rings=[("w", 1, 2), ("x", 2, 1)])
    p.params = T["params"]
    tX, tW = p.tensor(T["Xh"], 128), p.tensor(T["Wd1"], 256)
    acc = p.acc("BIG")
    xs = [p.load_stage(tX, col0=64 * j, row0=0, tile_rows=True, ring="x") for j in range(3)]     # 3 boxes, 2 stages
    for j in range(3):
        b = p.load_stage(tW, col0=64 * j, row0=0, ring="w")       # never reached: the third x load blocks the role
        p.mma(xs[j], b, n=256, acc=acc, k_steps=4, accumulate=j > 0, acc_last=(j == 2), a_release=True)
    p.epi_box(acc, 0, chain.EPI_BIAS_ELU, bias_off=T["b_d1"], last=True, has_reader=False)
    with pytest.raises(chain.ChainHazard, match="deadlock"):
        chain.Emulator(p, 128, seed=0).run()


def test_timing_model_tracks_measured_periods():
    """profiles/chain_model.py (the cycle model used to explore schedules without a GPU) must run every shipped
    program to completion and stay within 20 % of the per-tile periods measured on the B200 (196608 rows)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles"))
    import chain_model as cm
    T = ck.make_tensors(512, 0)
    for prog in (chain.teacher_forward_program(T), chain.trunk_backward_program(T), chain.adaptation_forward_program(T),
                 chain.adaptation_backward_program(T)):
        m = cm.Model(prog, tiles=5).run()
        r = m.report()
        assert all(e > 0 for e in m.tile_end), prog.name                  # every tile finished: no modelled deadlock
        assert abs(r["period"] / cm.MEASURED[prog.name] - 1) < 0.20, (prog.name, r["period"], cm.MEASURED[prog.name])


def test_three_and_four_worker_teacher_forward():
    """Programs for three (the shipped teacher forward: three chunk accumulators, the 128-column layer accumulator
    shares the columns of two of them) and four epilogue workers (teacher_forward_program4: 256-column first-layer
    staging, every narrow accumulator aliased into it, alias waits) must be hazard free and numerically right, for
    every variant the product builds."""
    for seed, n_ctas in ((0, 1), (1, 2)):
        for nw, build in ((3, lambda T: chain.teacher_forward_program(T, n_workers=3)), (4, chain.teacher_forward_program4)):
            T = ck.make_tensors(ROWS, seed)
            ref = ck.ref_teacher(T)
            prog = build(T)
            assert {o["worker"] for o in prog.epis} == set(range(nw))
            chain.Emulator(prog, ROWS, n_ctas=n_ctas, seed=seed).run()
            _close(T, ref, ("H1", "H2", "Xac", "Y1", "A2", "A3", "C2", "C3", "mean", "value"))
    for kw in (dict(save=False), dict(trunk=False), dict(want_value=False), dict(want_mean=False, save=False)):
        for build in (lambda T, **k: chain.teacher_forward_program(T, n_workers=3, **k), chain.teacher_forward_program4):
            T = ck.make_tensors(ROWS, 3)
            ref = ck.ref_teacher(T)
            chain.Emulator(build(T, **kw), ROWS, seed=4).run()
            keys = ["Xac"]
            if kw.get("trunk", True):
                keys += (["mean"] if kw.get("want_mean", True) else []) + (["value"] if kw.get("want_value", True) else [])
            _close(T, ref, keys)
    # more than four workers: refused loudly (argument validation precedes every CUDA call)
    from rapid_locomotion_rl_b200 import _lib
    T = ck.make_tensors(512, 0)
    prog = chain.teacher_forward_program(T, n_workers=3)
    prog.epis[0]["worker"] = 4
    with pytest.raises(_lib.RlError, match="worker"):
        prog.compile()


def test_tanh_programs():
    """The high_level_policy networks (high_level_policy/ppo/actor_critic.py:15): the same programs with `activation = "tanh"`
    - the emulator evaluates tanh / (1 - y^2) in the hidden-layer epilogues and the packed ops carry the kernel's tanh modes."""
    ck.set_activation("tanh")
    try:
        T = ck.make_tensors(ROWS, 7)
        ref = ck.ref_teacher(T)
        prog = chain.teacher_forward_program(T)
        prog.activation = "tanh"
        chain.Emulator(prog, ROWS, seed=7).run()
        _close(T, ref, ("H1", "H2", "Xac", "Y1", "A2", "A3", "C2", "C3", "mean", "value"))
        prog.pack()
        E = prog._packed[2]
        modes = {E[i].mode for i in range(len(prog.epis))}
        assert chain.EPI_BIAS_TANH in modes and chain.EPI_BIAS_ELU not in modes
        T = ck.make_tensors(ROWS, 8)
        fwd = ck.ref_teacher(T)
        for k in ("H1", "H2", "Y1", "A2", "A3", "C2", "C3"):
            T[k].copy_(fwd[k].to(torch.bfloat16))
        ref = ck.ref_trunk_backward(T)
        prog = chain.trunk_backward_program(T)
        prog.activation = "tanh"
        chain.Emulator(prog, ROWS, seed=8).run()
        _close(T, ref, ("dA3", "dA2", "dC3", "dC2", "dY1", "dLat", "dH2", "dH1"))
        prog.pack()
        E = prog._packed[2]
        assert chain.EPI_DTANH in {E[i].mode for i in range(len(prog.epis))}
    finally:
        ck.set_activation("elu")
