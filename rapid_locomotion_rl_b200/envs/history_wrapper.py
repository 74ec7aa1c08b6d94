"""HistoryWrapper with a ring buffer instead of a shift-concatenate.

Mirror of mini_gym/envs/wrappers/history_wrapper.py:6-41.  The reference rebuilds the whole
[N, 15*num_obs] history with torch.cat every step (2.35 KB read + 2.52 KB written per env).  Here
each env owns 2*H slots; a push writes the new observation twice (slots k and k+H), so the newest H
observations are always one contiguous span of the row, oldest first - the element order of the
reference - and `obs_history` is a zero-copy row-strided view of it.
"""
import torch

from .. import _lib


class HistoryWrapper:
    def __init__(self, env):
        self.env = env
        self.obs_history_length = env.cfg.env.num_observation_history
        self.num_obs_history = self.obs_history_length * env.num_obs
        self._ring = torch.zeros(env.num_envs, 2 * self.num_obs_history, dtype=torch.float, device=env.device)
        self._slot = self.obs_history_length - 1   # so that the first push lands in slot 0
        self._lib = _lib.lib()

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def obs_history(self):
        w = self.env.num_obs
        s = self._slot
        return self._ring[:, (s + 1) * w:(s + 1 + self.obs_history_length) * w]

    def _push(self, obs):
        self._slot = (self._slot + 1) % self.obs_history_length
        _lib.check(self._lib.rl_history_push(self._ring.data_ptr(), obs.data_ptr(), self.env.num_envs,
                                             self.env.num_obs, self.obs_history_length, self._slot,
                                             _lib.current_stream()))

    def step(self, action):
        obs, rew, done, info = self.env.step(action)
        privileged_obs = info["privileged_obs"]
        self._push(obs)
        return {"obs": obs, "privileged_obs": privileged_obs, "obs_history": self.obs_history}, rew, done, info

    def get_observations(self):
        obs = self.env.get_observations()
        privileged_obs = self.env.get_privileged_observations()
        self._push(obs)  # the reference shifts the history here too (:29)
        return {"obs": obs, "privileged_obs": privileged_obs, "obs_history": self.obs_history}

    def reset_idx(self, env_ids):
        return self.env.reset_idx(env_ids, obs_history=self._ring)

    def reset(self):
        ret = self.env.reset()
        privileged_obs = self.env.get_privileged_observations()
        self._ring.zero_()
        return {"obs": ret, "privileged_obs": privileged_obs, "obs_history": self.obs_history}
