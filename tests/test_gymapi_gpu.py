"""gymapi boundary: LeggedRobot on a LIVE simulator through `GymApiSim` - the reference's call sequence
per step (mini_gym/envs/base/legged_robot.py:116-126, :143-146), the index-table gather path for envs that hold more than
one actor (:1266-1277 -> :156, :124, :165-170), and int32 actor-id lists for the indexed state writes of a reset
(:700-745).  The simulator is a recording stand-in with the gymapi surface the path uses (tensors on the device)."""
import numpy as np
import pytest
import torch

from cases import build_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _GymTorch:
    @staticmethod
    def wrap_tensor(h):
        return h

    @staticmethod
    def unwrap_tensor(t):
        return t


class RecordingGym:
    """`extra_actors` one-body actors precede the robot inside every env, so the robot's rows are NOT the identity."""

    def __init__(self, n_envs, n_bodies, n_dof, extra_actors=0, seed=0):
        self.n, self.nb, self.nd, self.x = n_envs, n_bodies, n_dof, extra_actors
        g = torch.Generator().manual_seed(seed)
        per_env_actors, per_env_bodies = 1 + extra_actors, n_bodies + extra_actors
        self.root = torch.zeros(n_envs * per_env_actors, 13); self.root[:, 6] = 1.0
        self.dof = torch.zeros(n_envs * n_dof, 2)
        self.contact = torch.zeros(n_envs * per_env_bodies, 3)
        self.rb = torch.randn(n_envs * per_env_bodies, 13, generator=g)
        self.root, self.dof, self.contact, self.rb = (t.to(DEV) for t in (self.root, self.dof, self.contact, self.rb))
        self.calls, self.torques_seen, self.indexed = [], [], []
        self.envs = list(range(n_envs))

    # index queries (:1266-1273)
    def find_actor_index(self, env, name, domain):
        return env * (1 + self.x) + self.x

    def get_actor_dof_index(self, env, actor, d, domain):
        return env * self.nd + d

    def find_actor_rigid_body_index(self, env, actor, name, domain):
        return env * (self.nb + self.x) + self.x + name

    def acquire_actor_root_state_tensor(self, sim): return self.root
    def acquire_dof_state_tensor(self, sim): return self.dof
    def acquire_net_contact_force_tensor(self, sim): return self.contact
    def acquire_rigid_body_state_tensor(self, sim): return self.rb

    def set_dof_actuation_force_tensor(self, sim, t):
        self.calls.append("set_dof_actuation_force_tensor"); self.torques_seen.append(t.clone())
        self._tau = t

    def simulate(self, sim):
        self.calls.append("simulate")
        d = self.dof.view(self.n, self.nd, 2)
        d[..., 1] += 0.01 * self._tau.view(self.n, self.nd)          # a deterministic stand-in integrator
        d[..., 0] += 0.005 * d[..., 1]

    def fetch_results(self, sim, wait): self.calls.append("fetch_results")
    def refresh_dof_state_tensor(self, sim): self.calls.append("refresh_dof_state_tensor")
    def refresh_actor_root_state_tensor(self, sim): self.calls.append("refresh_actor_root_state_tensor")
    def refresh_net_contact_force_tensor(self, sim): self.calls.append("refresh_net_contact_force_tensor")
    def refresh_rigid_body_state_tensor(self, sim): self.calls.append("refresh_rigid_body_state_tensor")

    def set_dof_state_tensor_indexed(self, sim, t, ids, n):
        self.calls.append("set_dof_state_tensor_indexed"); self.indexed.append(("dof", t, ids.clone(), n))

    def set_actor_root_state_tensor_indexed(self, sim, t, ids, n):
        self.calls.append("set_actor_root_state_tensor_indexed"); self.indexed.append(("root", t, ids.clone(), n))


def _env_on(gym, case, n):
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    from rapid_locomotion_rl_b200.sim import GymApiSim
    cfg, robot, terrain = build_case(case, n)
    sim = GymApiSim(gym, _GymTorch, "sim", gym.envs, [0] * n, "robot", list(range(robot.num_bodies)), 12, DEV)
    env = LeggedRobot(cfg, sim=sim, sim_device=DEV, headless=True, terrain=terrain, seed=9)
    return env, sim, robot


def _fill(gym, sim, robot, n, seed=1):
    """The same per-robot state whatever the actor layout."""
    from rapid_locomotion_rl_b200.sim import synthetic_state
    st = synthetic_state(seed, n, robot.num_bodies, 12, np.zeros(12, np.float32), [3, 6, 9, 12], [0], z0=0.3)
    gym.root[sim.actor_indices] = torch.from_numpy(st["root_states"]).to(DEV)
    gym.dof[sim.dof_indices] = torch.from_numpy(st["dof_state"]).view(-1, 2).to(DEV)
    gym.contact[sim.rb_indices] = torch.from_numpy(st["contact_forces"]).view(-1, 3).to(DEV)
    sim.refresh()          # the state was written behind the adapter's back: re-gather (a no-op for identity tables)


@pytest.mark.parametrize("extra", [0, 2])
def test_live_step_call_sequence_and_gather_path(extra):
    n = 96
    results = {}
    for x in sorted({0, extra}):
        cfg0, robot, _ = build_case("mc_flat", n)
        gym = RecordingGym(n, robot.num_bodies, 12, extra_actors=x)
        env, sim, robot = _env_on(gym, "mc_flat", n)
        assert sim.identity == (x == 0)
        if x == 0:
            assert env.root_states.data_ptr() == gym.root.data_ptr() and env.dof_state.data_ptr() == gym.dof.data_ptr()   # zero copy
        _fill(gym, sim, robot, n)
        other_rows = None
        if x:
            mask = torch.ones(gym.root.shape[0], dtype=torch.bool, device=DEV); mask[sim.actor_indices] = False
            other_rows = (mask, gym.root[mask].clone())
        env.commands[:, :3] = torch.tensor([0.5, 0.0, 0.2], device=DEV)
        gym.calls.clear(); gym.torques_seen.clear()
        a = torch.from_numpy(np.random.default_rng(3).normal(0, 1, (n, 12)).astype(np.float32)).to(DEV)
        env._inject = dict(noise_u=torch.full((n, 42), 0.5, device=DEV))
        obs, priv, rew, reset, _ = env.step(a)
        torch.cuda.synchronize()
        dec = cfg0.control.decimation
        sub = ["set_dof_actuation_force_tensor", "simulate", "fetch_results", "refresh_dof_state_tensor"]
        want = sub * dec + ["refresh_actor_root_state_tensor", "refresh_dof_state_tensor", "refresh_net_contact_force_tensor",
                            "refresh_rigid_body_state_tensor"]
        assert gym.calls[:len(want)] == want, gym.calls                      # :116-126 then :143-146
        assert len(gym.torques_seen) == dec and torch.equal(gym.torques_seen[-1], env.torques)
        assert not torch.equal(gym.torques_seen[0], gym.torques_seen[-1])    # the torque kernel saw the DOF state move
        assert torch.equal(env.last_dof_vel, env.dof_vel)
        if other_rows is not None:
            assert torch.equal(gym.root[other_rows[0]], other_rows[1])       # rows of the other actors untouched
        results[x] = [t.clone() for t in (obs, priv, rew, reset, env.torques)]
    if extra:
        for u, v in zip(results[0], results[extra]):
            assert torch.equal(u, v)                                         # gathered path == zero-copy path


def test_reset_hands_int32_actor_ids_to_the_simulator():
    n, extra = 64, 1
    cfg0, robot, _ = build_case("mc_flat", n)
    gym = RecordingGym(n, robot.num_bodies, 12, extra_actors=extra)
    env, sim, robot = _env_on(gym, "mc_flat", n)
    _fill(gym, sim, robot, n)
    gym.calls.clear(); gym.indexed.clear()
    ids = torch.tensor([3, 10, 11, 40], device=DEV)
    env.reset_idx(ids)
    torch.cuda.synchronize()
    kinds = [k for k, *_ in gym.indexed]
    assert kinds == ["dof", "root"]
    for kind, tensor, got_ids, cnt in gym.indexed:
        assert got_ids.dtype == torch.int32 and cnt == 4
        assert got_ids.tolist() == [(1 + extra) * i + extra for i in ids.tolist()]            # find_actor_index of each env
        assert tensor.data_ptr() == (gym.dof if kind == "dof" else gym.root).data_ptr()       # the simulator's own tensor
    # the reset rows reached the simulator tensor through the scatter (:711), the others kept their state
    d = gym.dof.view(n, 12, 2)
    assert torch.equal(d[ids][..., 0], env.default_dof_pos.expand(4, 12)) and float(d[ids][..., 1].abs().max()) == 0.0
    assert float(d[0][..., 1].abs().max()) > 0
    root = gym.root[sim.actor_indices]
    torch.testing.assert_close(root[ids, 2], env.env_origins[ids, 2] + cfg0.init_state.pos[2])
