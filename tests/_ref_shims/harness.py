"""Harness that imports the UNMODIFIED reference from /root/reference through the
shims in this directory and builds a vectorised env on synthetic simulator state.

Test infrastructure only.  It exists to (a) pin the oracle in oracle/ against the
reference itself and (b) generate the golden fixtures under tests/golden/.  It
needs /root/reference, so nothing that runs on the GPU box may import it.
"""
import os
import sys

import numpy as np

_SHIMS = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(os.path.dirname(os.path.dirname(_SHIMS)), "oracle", "_ref")      # oracle/make_ref.py


def _reference_root():
    """RL_REFERENCE_ROOT, else the source tree of the build container, else the copy staged by oracle/make_ref.py
    (the only one that exists on the GPU box)."""
    env = os.environ.get("RL_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isdir(os.path.join(cand, "mini_gym")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _reference_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mini_gym"))


def install():
    """Put the shims and the reference on sys.path (idempotent)."""
    if not hasattr(np, "int"):
        np.int = int  # legged_robot.py:1065 uses the removed alias
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)


class SyntheticTerrain:
    """Replacement for mini_gym.utils.terrain.Terrain (needs closed terrain_utils).

    Provides exactly the attributes the env core reads: `cfg` (+ env_length,
    env_width, x_offset, env_origins), tot_rows/tot_cols, heightsamples int16,
    vertices/triangles placeholders (terrain.py:24-41,62-70,166-184).
    """
    height_fn = None  # optional callable(rows, cols) -> int16 array

    def __init__(self, cfg, num_robots, eval_cfg=None, num_eval_robots=0):
        self.cfg = cfg
        self.type = cfg.mesh_type
        cfg.env_length = cfg.terrain_length
        cfg.env_width = cfg.terrain_width
        cfg.x_offset = 0
        cfg.rows_offset = 0
        per_env_w = int(cfg.terrain_length / cfg.horizontal_scale)
        per_env_l = int(cfg.terrain_width / cfg.horizontal_scale)
        border = int(cfg.border_size / cfg.horizontal_scale)
        self.tot_cols = int(cfg.num_cols * per_env_w) + 2 * border
        self.tot_rows = int(cfg.num_rows * per_env_l) + 2 * border
        if SyntheticTerrain.height_fn is not None:
            self.heightsamples = SyntheticTerrain.height_fn(self.tot_rows, self.tot_cols)
        else:
            self.heightsamples = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
        self.height_field_raw = self.heightsamples
        origins = np.zeros((cfg.num_rows, cfg.num_cols, 3))
        for i in range(cfg.num_rows):
            for j in range(cfg.num_cols):
                sx, ex = border + i * per_env_l, border + (i + 1) * per_env_l
                sy, ey = border + j * per_env_w, border + (j + 1) * per_env_w
                origins[i, j] = [(i + 0.5) * cfg.terrain_length, (j + 0.5) * cfg.terrain_width,
                                 np.max(self.heightsamples[sx:ex, sy:ey]) * cfg.vertical_scale]
        cfg.env_origins = origins
        self.vertices = np.zeros((3, 3), dtype=np.float32)
        self.triangles = np.zeros((1, 3), dtype=np.uint32)
        self.eval_cfg = eval_cfg
        if eval_cfg is not None:
            # utils/terrain.py:49-57, 166-184: the evaluation tiles are appended below the training tiles
            ev = SyntheticTerrain(eval_cfg, num_eval_robots)
            eval_cfg.x_offset = self.tot_rows
            eval_cfg.rows_offset = cfg.num_rows
            eval_cfg.env_origins[:, :, 0] += eval_cfg.x_offset * eval_cfg.horizontal_scale
            hs = np.zeros((self.tot_rows + ev.tot_rows, max(self.tot_cols, ev.tot_cols)), dtype=np.int16)
            hs[:self.tot_rows, :self.tot_cols] = self.heightsamples
            hs[self.tot_rows:, :ev.tot_cols] = ev.heightsamples
            self.heightsamples = self.height_field_raw = hs
            self.tot_rows, self.tot_cols = hs.shape


def clone_cfg(c, name=None):
    """A second, independent Cfg class tree (the reference's Cfg is one process-global class): what a caller passes as
    `eval_cfg`."""
    import copy
    from params_proto.neo_proto import Meta, PrefixProto
    ns = {}
    for k, v in vars(c).items():
        ns[k] = clone_cfg(v, k) if isinstance(v, Meta) else copy.deepcopy(v)
    return Meta(name or c.__name__, (PrefixProto,), ns, cli=False)


def make_reference_env(robot="mini_cheetah", num_envs=64, device="cpu", rough=False, height_fn=None,
                       cfg_hook=None, history=True, eval_hook=None):
    """Build VelocityTrackingEasyEnv (+HistoryWrapper) of the reference on the fake backend.

    NOTE: the reference's `Cfg` is a process-global class mutated by config_*();
    call this once per process per configuration.
    """
    install()
    import torch  # noqa: F401
    from mini_gym.envs.base.legged_robot_config import Cfg
    import mini_gym.envs.base.legged_robot as lr
    SyntheticTerrain.height_fn = height_fn
    lr.Terrain = SyntheticTerrain
    if robot == "mini_cheetah":
        from mini_gym.envs.mini_cheetah.mini_cheetah_config import config_mini_cheetah
        config_mini_cheetah(Cfg)
    elif robot == "go1":
        from mini_gym.envs.go1.go1_config import config_go1
        config_go1(Cfg)
    else:
        raise ValueError(robot)
    Cfg.env.num_envs = num_envs
    Cfg.env.record_video = False
    if rough:
        Cfg.terrain.mesh_type = 'heightfield'
        Cfg.terrain.measure_heights = True
        Cfg.terrain.curriculum = True
        Cfg.terrain.terrain_proportions = [0.1, 0.1, 0.35, 0.25, 0.2]
        Cfg.env.num_observations = 42 + 17 * 11
    if cfg_hook is not None:
        cfg_hook(Cfg)
    from mini_gym.envs.mini_cheetah.velocity_tracking import VelocityTrackingEasyEnv
    eval_cfg = None
    if eval_hook is not None:          # train / eval split: `eval_hook` edits a copy of the finished training Cfg
        eval_cfg = clone_cfg(Cfg)
        eval_hook(eval_cfg)
    env = VelocityTrackingEasyEnv(sim_device=device, headless=True, cfg=Cfg, eval_cfg=eval_cfg)
    if eval_cfg is not None:
        env.eval_cfg_used = eval_cfg
    if history:
        from mini_gym.envs.wrappers.history_wrapper import HistoryWrapper
        env = HistoryWrapper(env)
    return env, Cfg
