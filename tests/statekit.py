"""Moves env state between the three implementations under test, in one canonical format.

Canonical state = dict of numpy arrays in the reference's AoS shapes:
  root_states [N,13], dof_state [N,12,2], contact_forces [N,NB,3], commands [N,4],
  last_actions/last_dof_vel/Kp_factors/Kd_factors/motor_strengths/torques/joint_pos_target [N,12],
  last_root_vel [N,6], friction_coeffs/restitutions/payloads [N], com_displacements [N,3],
  feet_air_time [N,4], last_contacts [N,4] bool, episode_length_buf [N] int64,
  base_lin_vel/base_ang_vel/projected_gravity [N,3], env_origins [N,3], terrain_levels/terrain_types [N],
  episode_sums/<name> [N], command_sums/<name> [N].
Sources/sinks: the reference env (through the shims), oracle.env_oracle.OracleEnv, and the
product rapid_locomotion_rl_b200.envs.LeggedRobot.
"""
import numpy as np
import torch

STATE_KEYS = ["root_states", "dof_state", "contact_forces", "commands", "last_actions", "last_dof_vel",
              "last_root_vel", "Kp_factors", "Kd_factors", "motor_strengths", "friction_coeffs", "restitutions",
              "payloads", "com_displacements", "feet_air_time", "last_contacts", "episode_length_buf",
              "env_origins", "terrain_levels", "terrain_types"]
DERIVED_KEYS = ["torques", "joint_pos_target", "base_lin_vel", "base_ang_vel", "projected_gravity"]


def _np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy().copy()
    return np.asarray(t).copy()


def _sums(prefix, d):
    return {"%s/%s" % (prefix, k): _np(v) for k, v in d.items()}


def state_from_reference(e):
    """`e` is the reference LeggedRobot (unwrapped).  In this fork root_states / dof_state are gathered
    copies (legged_robot.py:156,124), so after a step they hold the state the step actually used."""
    N = e.num_envs
    s = dict(
        root_states=_np(e.root_states), dof_state=_np(e.dof_state).reshape(N, e.num_dof, 2),
        contact_forces=_np(e.all_contact_forces).reshape(N, e.num_bodies, 3), commands=_np(e.commands),
        last_actions=_np(e.last_actions), last_dof_vel=_np(e.last_dof_vel), last_root_vel=_np(e.last_root_vel),
        Kp_factors=_np(e.Kp_factors), Kd_factors=_np(e.Kd_factors), motor_strengths=_np(e.motor_strengths),
        friction_coeffs=_np(e.friction_coeffs), restitutions=_np(e.restitutions), payloads=_np(e.payloads),
        com_displacements=_np(e.com_displacements), feet_air_time=_np(e.feet_air_time),
        last_contacts=_np(e.last_contacts), episode_length_buf=_np(e.episode_length_buf),
        env_origins=_np(e.env_origins), terrain_levels=_np(e.terrain_levels), terrain_types=_np(e.terrain_types),
        torques=_np(e.torques), base_lin_vel=_np(e.base_lin_vel), base_ang_vel=_np(e.base_ang_vel),
        projected_gravity=_np(e.projected_gravity))
    if hasattr(e, "joint_pos_target"):
        s["joint_pos_target"] = _np(e.joint_pos_target)
    s.update(_sums("episode_sums", e.episode_sums))
    s.update(_sums("command_sums", e.command_sums))
    return s


def state_from_oracle(o):
    s = {k: _np(getattr(o, k)) for k in STATE_KEYS}
    for k in DERIVED_KEYS:
        if hasattr(o, k):
            s[k] = _np(getattr(o, k))
    s.update(_sums("episode_sums", o.episode_sums))
    s.update(_sums("command_sums", o.command_sums))
    return s


def apply_to_oracle(o, s):
    for k in STATE_KEYS:
        if k in s:
            cur = getattr(o, k)
            setattr(o, k, torch.from_numpy(np.asarray(s[k])).to(cur.dtype).to(cur.device).clone())
    for name in o.episode_sums:
        key = "episode_sums/" + name
        if key in s:
            o.episode_sums[name] = torch.from_numpy(s[key]).to(o.dtype).clone()
    for name in o.command_sums:
        key = "command_sums/" + name
        if key in s:
            o.command_sums[name] = torch.from_numpy(s[key]).to(o.dtype).clone()


def state_from_product(e):
    N = e.num_envs
    s = dict(
        root_states=_np(e.root_states), dof_state=_np(e.dof_state).reshape(N, e.num_dof, 2),
        contact_forces=_np(e.contact_forces), commands=_np(e.commands), last_actions=_np(e.last_actions),
        last_dof_vel=_np(e.last_dof_vel), last_root_vel=_np(e.last_root_vel), Kp_factors=_np(e.Kp_factors),
        Kd_factors=_np(e.Kd_factors), motor_strengths=_np(e.motor_strengths), friction_coeffs=_np(e.friction_coeffs),
        restitutions=_np(e.restitutions), payloads=_np(e.payloads), com_displacements=_np(e.com_displacements),
        feet_air_time=_np(e.feet_air_time), last_contacts=_np(e.last_contacts),
        episode_length_buf=_np(e.episode_length_buf), env_origins=_np(e.env_origins),
        terrain_levels=_np(e.terrain_levels), terrain_types=_np(e.terrain_types), torques=_np(e.torques),
        joint_pos_target=_np(e.joint_pos_target), base_lin_vel=_np(e.base_lin_vel), base_ang_vel=_np(e.base_ang_vel),
        projected_gravity=_np(e.projected_gravity))
    s.update(_sums("episode_sums", e.episode_sums))
    s.update(_sums("command_sums", e.command_sums))
    return s


def apply_to_product(e, s):
    """Copy a canonical state into the product env's device buffers (views write through to SoA)."""
    dev = e.device

    def put(dst, key, shape=None):
        if key in s:
            src = torch.from_numpy(np.asarray(s[key]))
            if shape is not None:
                src = src.reshape(shape)
            dst.copy_(src.to(dst.dtype).to(dev))
    N = e.num_envs
    put(e.root_states, "root_states"); put(e.dof_state, "dof_state", (N * e.num_dof, 2))
    put(e.all_contact_forces, "contact_forces", (N * e.num_bodies, 3)); put(e.commands, "commands")
    for k in ("last_actions", "last_dof_vel", "last_root_vel", "Kp_factors", "Kd_factors", "motor_strengths",
              "friction_coeffs", "restitutions", "payloads", "com_displacements", "feet_air_time",
              "episode_length_buf", "env_origins", "terrain_levels", "terrain_types"):
        put(getattr(e, k), k)
    if "last_contacts" in s:
        e._last_contacts_u8.copy_(torch.from_numpy(np.asarray(s["last_contacts"]).astype(np.uint8)).to(dev))
    for name, row in e.episode_sums.items():
        put(row, "episode_sums/" + name)
    for name, row in e.command_sums.items():
        put(row, "command_sums/" + name)


def assert_state_close(got, want, rtol, atol, exact_keys=(), skip=(), label=""):
    """Float keys within tolerance, integer/bool keys and `exact_keys` exactly."""
    bad = []
    for k, w in want.items():
        if k in skip or k not in got:
            continue
        g = np.asarray(got[k])
        w = np.asarray(w)
        if g.shape != w.shape:
            bad.append("%s: shape %s vs %s" % (k, g.shape, w.shape))
            continue
        if w.dtype.kind in "biu" or k in exact_keys:
            if not np.array_equal(g, w):
                bad.append("%s: %d mismatching entries (exact)" % (k, int((g != w).sum())))
        else:
            if not np.allclose(g, w, rtol=rtol, atol=atol):
                err = np.abs(g.astype(np.float64) - w.astype(np.float64))
                bad.append("%s: max abs err %.3e (max |ref| %.3e)" % (k, err.max(), np.abs(w).max()))
    assert not bad, "%s state mismatch:\n  " % label + "\n  ".join(bad)
