"""Simulator boundary for runs without Isaac Gym.

The closed-source Isaac Gym / PhysX integrator stays behind the gymapi boundary and is not
reimplemented (BASELINE.json north_star).  What the env core needs from it are four state
tensors in the PhysX AoS layout (legged_robot.py:939-971):

    actor_root_state [A,13]  pos3, quat xyzw4, lin_vel3, ang_vel3
    dof_state        [A*12,2] (pos, vel)
    net_contact_force[A*NB,3]
    rigid_body_state [A*NB,13]

`SyntheticSim` owns such tensors on the device and can fill them with the synthetic state
distribution of SURVEY.md section 8(d); a real gymapi integration hands in zero-copy views of
PhysX memory instead.
"""
import numpy as np
import torch

from .robots import RobotSpec


class SyntheticSim:
    def __init__(self, robot: RobotSpec, num_envs: int, device="cuda:0"):
        self.robot = robot
        self.num_envs = num_envs
        self.device = torch.device(device)
        nb, nd = robot.num_bodies, robot.num_dof
        self.root_states = torch.zeros(num_envs, 13, device=self.device)
        self.root_states[:, 6] = 1.0
        self.dof_state = torch.zeros(num_envs * nd, 2, device=self.device)
        self.contact_forces = torch.zeros(num_envs * nb, 3, device=self.device)
        self.rigid_body_state = torch.zeros(num_envs * nb, 13, device=self.device)

    # gymapi-like no-ops so the env core reads the same either way
    def simulate(self):
        pass

    def refresh(self):
        pass

    def set_dof_actuation_force_tensor(self, torques):
        pass

    def fill_random(self, seed, default_dof_pos, feet_idx, term_idx, z0=0.30, teleport_band_frac=0.0,
                    uniform_rotations=False):
        """Synthetic state of SURVEY.md 8(d): generated on the CPU with a seeded numpy generator so
        the oracle and the kernels see identical inputs, then copied to the device."""
        n, nb, nd = self.num_envs, self.robot.num_bodies, self.robot.num_dof
        st = synthetic_state(seed, n, nb, nd, np.asarray(default_dof_pos, np.float32).reshape(-1), feet_idx,
                             term_idx, z0, teleport_band_frac, uniform_rotations)
        self.root_states.copy_(torch.from_numpy(st["root_states"]))
        self.dof_state.copy_(torch.from_numpy(st["dof_state"]).view(-1, 2))
        self.contact_forces.copy_(torch.from_numpy(st["contact_forces"]).view(-1, 3))
        self.rigid_body_state.copy_(torch.from_numpy(st["rigid_body_state"]).view(-1, 13))
        return st


def synthetic_state(seed, n, nb, nd, default_dof_pos, feet_idx, term_idx, z0=0.30, teleport_band_frac=0.0,
                    uniform_rotations=False):
    """numpy float32 arrays: root_states [n,13], dof_state [n,nd,2], contact_forces [n,nb,3],
    rigid_body_state [n,nb,13] (SURVEY.md 8(d) 'Synthetic inputs')."""
    rng = np.random.default_rng(seed)
    root = np.zeros((n, 13), np.float32)
    root[:, 0] = rng.uniform(10.0, 70.0, n)
    root[:, 1] = rng.uniform(10.0, 150.0, n)
    if teleport_band_frac > 0:
        k = max(1, int(n * teleport_band_frac))
        idx = rng.choice(n, k, replace=False)
        root[idx[: k // 2], 0] = rng.uniform(0.0, 1.9, k // 2)
        root[idx[k // 2:], 1] = rng.uniform(158.1, 160.0, k - k // 2)
    root[:, 2] = z0 + rng.normal(0.0, 0.02, n)
    if uniform_rotations:
        q = rng.normal(0.0, 1.0, (n, 4))
    else:
        q = np.concatenate([rng.normal(0.0, 0.1, (n, 3)), np.ones((n, 1))], axis=1)
    root[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    root[:, 7:13] = rng.normal(0.0, 0.5, (n, 6))
    dof = np.zeros((n, nd, 2), np.float32)
    dof[:, :, 0] = default_dof_pos[None, :] + rng.normal(0.0, 0.3, (n, nd))
    dof[:, :, 1] = rng.normal(0.0, 3.0, (n, nd))
    con = rng.normal(0.0, 0.5, (n, nb, 3)).astype(np.float32)
    big = rng.random((n, len(term_idx))) < 0.02
    for k, b in enumerate(term_idx):
        con[big[:, k], b, :] *= 20.0
    for b in feet_idx:
        con[:, b, 2] = np.abs(rng.normal(0.0, 30.0, n)) * (rng.random(n) < 0.5)
    rb = rng.normal(0.0, 1.0, (n, nb, 13)).astype(np.float32)
    return dict(root_states=root.astype(np.float32), dof_state=dof.astype(np.float32),
                contact_forces=con.astype(np.float32), rigid_body_state=rb)


def synthetic_heightfield(rows, cols, seed=0):
    """int16 heightfield: smooth noise of +-20 counts plus stair bands."""
    rng = np.random.default_rng(seed)
    coarse = rng.uniform(-20.0, 20.0, (rows // 16 + 2, cols // 16 + 2))
    yy = np.arange(rows) / 16.0
    xx = np.arange(cols) / 16.0
    y0, x0 = yy.astype(int), xx.astype(int)
    fy, fx = (yy - y0)[:, None], (xx - x0)[None, :]
    a = coarse[y0][:, x0]; b = coarse[y0][:, x0 + 1]; c = coarse[y0 + 1][:, x0]; d = coarse[y0 + 1][:, x0 + 1]
    field = a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx
    stairs = ((np.arange(cols) // 31) % 6) * 10.0
    band = ((np.arange(rows) // 80) % 3 == 1)[:, None]
    field = field + band * stairs[None, :]
    return np.round(field).astype(np.int16)


class GymApiSim:
    """Boundary to a LIVE Isaac Gym simulation: everything LeggedRobot does through `self.gym` on the
    hot path (mini_gym/envs/base/legged_robot.py:116-126, :143-160, :303-312, :700-745, :951-971, :1266-1277), kept
    behind one object so that the env core only ever sees four state tensors in the PhysX layout.

        sim = GymApiSim(gym, gymtorch, sim_handle, env_handles, actor_handles, actor_name, body_names, num_dof, device)
        env = LeggedRobot(cfg, sim=sim, ...)

    * state tensors: `acquire_*_tensor` + `gymtorch.wrap_tensor` once (:951-971).  The reference then GATHERS the robot's
      rows through three index tables (actor / DOF / rigid-body indices in the simulation domain, :1266-1277 -> :156,
      :124, :165-170) every step, because an env may hold other actors.  Here the tables are built the same way and
      checked: when they are the identity (one actor per env - always true in this snapshot, `WorldAsset` is disabled)
      `root_states` / `dof_state` / `contact_forces` / `rigid_body_state` ARE the simulator's tensors (zero copy);
      otherwise they are persistent gathered copies refreshed by `refresh()` (index_select into the same storage, so
      kernel arguments and captured graphs stay valid) and scattered back before an indexed set.
    * stepping (:116-126): `apply_torques_and_step(torques)` = set_dof_actuation_force_tensor -> simulate ->
      fetch_results -> refresh_dof_state_tensor (+ gather).
    * resets (:700-745): `push_dof_state(env_ids)` / `push_root_state(env_ids)` hand the simulator the int32 actor ids of
      the reset envs (`set_dof_state_tensor_indexed` / `set_actor_root_state_tensor_indexed`); the id list is a device
      gather from the actor table - no per-env Python loop over `find_actor_index` (the reference's :701, :723)."""
    live = True

    def __init__(self, gym, gymtorch, sim, envs, actor_handles, actor_name, body_names, num_dof, device, domain_sim=None):
        self.gym, self.gymtorch, self.sim, self.device = gym, gymtorch, sim, torch.device(device)
        self.num_envs, nd, nb = len(envs), int(num_dof), len(body_names)
        dom = domain_sim
        actor_idx, dof_idx, rb_idx = [], [], []
        for env, ah in zip(envs, actor_handles):                     # :1266-1273
            actor_idx.append(gym.find_actor_index(env, actor_name, dom))
            dof_idx.extend(gym.get_actor_dof_index(env, ah, d, dom) for d in range(nd))
            rb_idx.extend(gym.find_actor_rigid_body_index(env, ah, n, dom) for n in body_names)
        long = lambda x: torch.tensor(x, dtype=torch.long, device=self.device)
        self.actor_indices, self.dof_indices, self.rb_indices = long(actor_idx), long(dof_idx), long(rb_idx)
        self.actor_ids_int32 = self.actor_indices.to(torch.int32)
        for fn in ("refresh_dof_state_tensor", "refresh_actor_root_state_tensor", "refresh_net_contact_force_tensor",
                   "refresh_rigid_body_state_tensor"):
            getattr(gym, fn)(sim)
        wrap = gymtorch.wrap_tensor
        self.all_root_states = wrap(gym.acquire_actor_root_state_tensor(sim))
        self.all_dof_state = wrap(gym.acquire_dof_state_tensor(sim))
        self.all_contact_forces = wrap(gym.acquire_net_contact_force_tensor(sim))
        self.all_rigid_body_state = wrap(gym.acquire_rigid_body_state_tensor(sim))
        for t in (self.all_root_states, self.all_dof_state, self.all_contact_forces, self.all_rigid_body_state):
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("simulator tensors must be contiguous float32 on %s (use_gpu_pipeline)" % self.device)

        def ident(idx, total):
            return idx.numel() == total and bool(torch.equal(idx, torch.arange(total, device=self.device)))
        self.identity = (ident(self.actor_indices, self.all_root_states.shape[0]) and
                         ident(self.dof_indices, self.all_dof_state.shape[0]) and
                         ident(self.rb_indices, self.all_contact_forces.shape[0]))
        if self.identity:
            self.root_states, self.dof_state = self.all_root_states, self.all_dof_state
            self.contact_forces, self.rigid_body_state = self.all_contact_forces, self.all_rigid_body_state
        else:
            self.root_states = self.all_root_states[self.actor_indices].contiguous()
            self.dof_state = self.all_dof_state[self.dof_indices].contiguous()
            self.contact_forces = self.all_contact_forces[self.rb_indices].contiguous()
            self.rigid_body_state = self.all_rigid_body_state[self.rb_indices].contiguous()

    # ---- per-step traffic -----------------------------------------------------------------------------------
    def _gather(self, what):
        if self.identity:
            return
        pairs = {"dof": (self.dof_state, self.all_dof_state, self.dof_indices),
                 "root": (self.root_states, self.all_root_states, self.actor_indices),
                 "contact": (self.contact_forces, self.all_contact_forces, self.rb_indices),
                 "rb": (self.rigid_body_state, self.all_rigid_body_state, self.rb_indices)}
        for k in what:
            dst, src, idx = pairs[k]
            torch.index_select(src, 0, idx, out=dst)

    def apply_torques_and_step(self, torques):
        """One physics sub-step (:117-124)."""
        g, s = self.gym, self.sim
        g.set_dof_actuation_force_tensor(s, self.gymtorch.unwrap_tensor(torques))
        g.simulate(s)
        g.fetch_results(s, True)
        g.refresh_dof_state_tensor(s)
        self._gather(("dof",))

    def refresh(self):
        """:143-146 + the gathers of :156, :165-170."""
        g, s = self.gym, self.sim
        g.refresh_actor_root_state_tensor(s)
        g.refresh_dof_state_tensor(s)
        g.refresh_net_contact_force_tensor(s)
        g.refresh_rigid_body_state_tensor(s)
        self._gather(("root", "dof", "contact", "rb"))

    # ---- resets ---------------------------------------------------------------------------------------------------
    def push_dof_state(self, env_ids):
        """:713-717 after the reset kernel wrote the DOF rows of `env_ids`."""
        ids = self.actor_ids_int32[env_ids].contiguous()
        if not self.identity:
            self.all_dof_state[self.dof_indices] = self.dof_state
        self.gym.set_dof_state_tensor_indexed(self.sim, self.gymtorch.unwrap_tensor(self.all_dof_state),
                                              self.gymtorch.unwrap_tensor(ids), len(ids))
        return ids

    def push_root_state(self, env_ids):
        """:739-741 (teleport :789-791 and push :765-766 use the same call)."""
        ids = self.actor_ids_int32[env_ids].contiguous()
        if not self.identity:
            self.all_root_states[self.actor_indices] = self.root_states
        self.gym.set_actor_root_state_tensor_indexed(self.sim, self.gymtorch.unwrap_tensor(self.all_root_states),
                                                     self.gymtorch.unwrap_tensor(ids), len(ids))
        return ids
