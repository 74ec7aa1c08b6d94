"""No-GPU checks of the C-ABI library: it builds for sm_100a, loads, exports every symbol the
header declares, its struct layouts match the generated ctypes binding, and argument validation
returns error codes (no compute calls are made here)."""
import ctypes as C

import pytest


def test_library_exports_every_declared_symbol():
    from rapid_locomotion_rl_b200 import _lib
    lib = _lib.lib()
    declared = _lib.declared_symbols()
    assert set(declared) == set(_lib.SIGNATURES), "binding table and header disagree"
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.rl_version()


def test_struct_layouts_match():
    from rapid_locomotion_rl_b200 import _lib
    lib = _lib.lib()
    for name, st in _lib.STRUCTS.items():
        assert lib.rl_sizeof(name.encode()) == C.sizeof(st), name
    assert lib.rl_sizeof(b"nope") == -1


def test_argument_validation_without_gpu():
    from rapid_locomotion_rl_b200 import _lib
    lib = _lib.lib()
    assert lib.rl_gae(None, None, None, None, None, None, 24, 10, 0.99, 0.95, None, None) == _lib.DEFINES["RL_ERR_BAD_ARG"]
    assert b"null" in lib.rl_last_error()
    assert lib.rl_history_push(None, None, 1, 1, 1, 0, None) == _lib.DEFINES["RL_ERR_BAD_ARG"]
    cfg = _lib.RlEnvCfg(); bufs = _lib.RlEnvBuffers()
    assert lib.rl_env_step_fused(C.byref(cfg), C.byref(bufs), 0, 0, None) == _lib.DEFINES["RL_ERR_BAD_CFG"]
    with pytest.raises(_lib.RlError):
        _lib.check(-1)
    assert lib.rl_gae_workspace_bytes(4000) >= 64 + 63 * 16


def test_product_refuses_cpu_device():
    """The product path must fail loudly without CUDA: no CPU fallback."""
    from rapid_locomotion_rl_b200 import _lib
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    from rapid_locomotion_rl_b200.ppo import RolloutStorage
    from cases import build_case
    cfg, robot, terrain = build_case("mc_flat", 8)
    with pytest.raises(_lib.RlError):
        LeggedRobot(cfg, sim_device="cpu", terrain=terrain)
    with pytest.raises(_lib.RlError):
        RolloutStorage(8, 4, [42], [18], [630], [12], device="cpu")


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    import os
    import re
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rapid_locomotion_rl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(root, f)
