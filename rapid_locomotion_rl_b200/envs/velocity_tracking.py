"""VelocityTrackingEasyEnv: the concrete env the training script builds.

Mirror of mini_gym/envs/mini_cheetah/velocity_tracking/velocity_tracking_easy_env.py:10-69:
`step(actions) -> (obs, rew, done, extras)` with `privileged_obs` inside `extras`.  The reference
also copies 11 arrays to host numpy on EVERY step (:50-61, eleven blocking device->host syncs);
here those entries are materialised lazily, only when a caller actually reads them.
"""
import torch

from .legged_robot import LeggedRobot


class LazyExtras(dict):
    """dict whose registered keys are computed on first access (then cached until the next step)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._lazy = {}

    def set_lazy(self, key, fn):
        self._lazy[key] = fn
        dict.pop(self, key, None)

    def __missing__(self, key):
        if key in self._lazy:
            val = self._lazy[key]()
            dict.__setitem__(self, key, val)
            return val
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def keys(self):
        return list(dict.keys(self)) + [k for k in self._lazy if not dict.__contains__(self, k)]

    def get(self, key, default=None):
        return self[key] if key in self else default


class VelocityTrackingEasyEnv(LeggedRobot):
    def __init__(self, sim_device="cuda:0", headless=True, num_envs=None, prone=False, deploy=False, cfg=None,
                 eval_cfg=None, initial_dynamics_dict=None, physics_engine="SIM_PHYSX", **kw):
        if num_envs is not None:
            cfg.env.num_envs = num_envs
        if prone:
            cfg.init_state.rot = [0.0, 1.0, 0.0, 0.0]
            cfg.init_state.pos = [0.0, 0.0, 0.15]
            cfg.asset.fix_base_link = True
        if deploy:
            raise NotImplementedError("deploy=True reconfigures the closed-source simulator terrain")
        super().__init__(cfg, None, physics_engine, sim_device, headless, eval_cfg, initial_dynamics_dict, **kw)
        self.extras = LazyExtras()

    def step(self, actions):
        obs, priv, rew, reset, _ = super().step(actions)
        ex = self.extras
        ex["privileged_obs"] = priv
        nb = self.num_bodies
        feet = self.feet_indices
        ex.set_lazy("joint_pos", lambda: self.dof_pos.cpu().numpy())
        ex.set_lazy("joint_vel", lambda: self.dof_vel.cpu().numpy())
        ex.set_lazy("joint_pos_target", lambda: self.joint_pos_target.cpu().numpy())
        ex.set_lazy("joint_vel_target", lambda: torch.zeros(12))
        ex.set_lazy("body_linear_vel", lambda: self.base_lin_vel.cpu().numpy())
        ex.set_lazy("body_angular_vel", lambda: self.base_ang_vel.cpu().numpy())
        ex.set_lazy("body_linear_vel_cmd", lambda: self.commands.cpu().numpy()[:, 0:2])
        ex.set_lazy("body_angular_vel_cmd", lambda: self.commands.cpu().numpy()[:, 2:])
        ex.set_lazy("contact_states", lambda: (self.contact_forces[:, feet, 2] > 1.0).cpu().numpy().copy())
        ex.set_lazy("foot_positions",
                    lambda: self.rigid_body_state.view(self.num_envs, nb, 13)[:, feet, 0:3].cpu().numpy().copy())
        ex.set_lazy("body_pos", lambda: self.root_states[:, 0:3].cpu().numpy())
        ex.set_lazy("torques", lambda: self.torques.cpu().numpy())
        return obs, rew, reset, ex

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device))
        return obs
