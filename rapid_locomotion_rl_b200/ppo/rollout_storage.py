"""RolloutStorage with a fused GAE.

Mirror of mini_gym_learn/ppo/rollout_storage.py:7-139: same constructor, tensor attributes
(`observations ... env_bins`, [T,N,.]), `add_transitions`, `compute_returns`, `mini_batch_generator`,
`clear`.  `compute_returns` (:76-90) runs as two launches of csrc/gae.cu instead of 24 x ~8 eager
ops and two global reductions.
"""
import torch

from .. import _lib


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = None
            self.privileged_observations = None
            self.observation_histories = None
            self.critic_observations = None
            self.actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.env_bins = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, obs_history_shape,
                 actions_shape, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RlError("RolloutStorage needs a CUDA device: compute_returns has no CPU fallback")
        self._lib = _lib.lib()
        self.obs_shape, self.privileged_obs_shape = obs_shape, privileged_obs_shape
        self.obs_history_shape, self.actions_shape = obs_history_shape, actions_shape
        T, N, dev = num_transitions_per_env, num_envs, self.device

        def z(*s, **k):
            return torch.zeros(T, N, *s, device=dev, **k)
        self.observations = z(*obs_shape)
        self.privileged_observations = z(*privileged_obs_shape)
        self.observation_histories = z(*obs_history_shape)
        self.rewards = z(1)
        self.actions = z(*actions_shape)
        self.dones = z(1, dtype=torch.uint8)
        self.actions_log_prob = z(1)
        self.values = z(1)
        self.returns = z(1)
        self.advantages = z(1)
        self.mu = z(*actions_shape)
        self.sigma = z(*actions_shape)
        self.env_bins = z(1)
        self.num_transitions_per_env, self.num_envs = T, N
        self.step = 0
        self._gae_ws = torch.zeros(int(self._lib.rl_gae_workspace_bytes(N)) // 8 + 1, dtype=torch.float64, device=dev)
        self._gae_stats = torch.zeros(3, dtype=torch.float64, device=dev)

    def store_observations(self, obs, privileged_obs, obs_history):
        """Rows of rollout_storage.py:57-60 for the CURRENT step, written when the action is taken (PPO.act): the env's
        observation buffers are rewritten in place by env.step, before add_transitions runs."""
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        self.observations[s].copy_(obs)
        self.privileged_observations[s].copy_(privileged_obs)
        self.observation_histories[s].copy_(obs_history)
        self._obs_stored_step = s

    def add_transitions(self, transition):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        stored = transition.observations is None and getattr(self, "_obs_stored_step", -1) == s
        if transition.observations is None and not stored:
            raise AssertionError("transition without observations: call PPO.act (store_observations) first")
        self._obs_stored_step = -1
        if self._fused_add(transition, s, stored):
            self.step += 1
            return
        if not stored:
            self.observations[s].copy_(transition.observations)
            self.privileged_observations[s].copy_(transition.privileged_observations)
            self.observation_histories[s].copy_(transition.observation_histories)
        self.actions[s].copy_(transition.actions)
        self.rewards[s].copy_(transition.rewards.view(-1, 1))
        self.dones[s].copy_(transition.dones.view(-1, 1))
        self.values[s].copy_(transition.values)
        self.actions_log_prob[s].copy_(transition.actions_log_prob.view(-1, 1))
        self.mu[s].copy_(transition.action_mean)
        self.sigma[s].copy_(transition.action_sigma)
        self.env_bins[s].copy_(transition.env_bins.view(-1, 1))
        self.step += 1

    def _fused_add(self, t, s, stored=False):
        """One launch for the eleven per-step copies (csrc/history.cu rl_storage_add).  Falls back to the
        copy_ sequence (still on the device) only for layouts the kernel does not take: non-fp32 sources,
        rows that are not unit-stride."""
        N = self.num_envs
        rows = () if stored else (t.observations, t.privileged_observations, t.observation_histories)
        f32 = rows + (t.actions, t.action_mean, t.action_sigma, t.rewards, t.values, t.actions_log_prob, t.env_bins)
        if any(x.dtype != torch.float32 or x.device != self.device or x.shape[0] != N for x in f32):
            return False
        if any(x.dim() != 2 or x.stride(1) != 1 for x in rows):
            return False
        small = [x if x.is_contiguous() else x.contiguous() for x in f32[len(rows):]]
        dones = t.dones
        if dones.dtype == torch.bool:
            dones = dones.view(torch.uint8)
        if dones.dtype != torch.uint8 or not dones.is_contiguous() or dones.numel() != N:
            return False
        q = _lib.RlStorageAdd()
        P = lambda x: x.data_ptr()
        if stored:
            q.obs = q.priv = q.hist = None
        else:
            q.obs, q.priv, q.hist = P(rows[0]), P(rows[1]), P(rows[2])
        q.actions, q.mu, q.sigma, q.rewards, q.values, q.logp, q.bins = (P(x) for x in small)
        q.dones = P(dones)
        q.dst_obs, q.dst_priv, q.dst_hist = (None, None, None) if stored else \
            (P(self.observations[s]), P(self.privileged_observations[s]), P(self.observation_histories[s]))
        q.dst_actions, q.dst_mu, q.dst_sigma = P(self.actions[s]), P(self.mu[s]), P(self.sigma[s])
        q.dst_rewards, q.dst_values, q.dst_logp = P(self.rewards[s]), P(self.values[s]), P(self.actions_log_prob[s])
        q.dst_bins, q.dst_dones = P(self.env_bins[s]), P(self.dones[s])
        dims = (self.observations.shape[-1], self.privileged_observations.shape[-1], self.observation_histories.shape[-1])
        q.ld_obs, q.ld_priv, q.ld_hist = dims if stored else (rows[0].stride(0), rows[1].stride(0), rows[2].stride(0))
        q.N, q.obs_dim, q.priv_dim, q.hist_dim, q.act_dim = N, dims[0], dims[1], dims[2], small[0].shape[-1]
        import ctypes as C
        self._keep = (small, dones)
        _lib.check(self._lib.rl_storage_add(C.byref(q), _lib.current_stream()))
        return True

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        """rollout_storage.py:76-90 - reverse-time GAE scan + unbiased-std normalisation, on device.
        With torch.distributed initialised the (sum, sumsq, count) statistics are all-reduced between
        the two phases, so the normalisation spans the env shards of every rank."""
        from ..sharding import all_reduce_sum_
        T, N = self.num_transitions_per_env, self.num_envs
        last_values = last_values.to(self.device, torch.float).contiguous()
        P, st = _lib.ptr, _lib.current_stream()
        _lib.check(self._lib.rl_gae_scan(P(self.rewards), P(self.values), P(self.dones), P(last_values),
                                         P(self.returns), P(self.advantages), T, N, float(gamma), float(lam),
                                         P(self._gae_ws), P(self._gae_stats), st))
        all_reduce_sum_(self._gae_stats)
        _lib.check(self._lib.rl_gae_normalize(P(self.advantages), T, N, P(self._gae_stats), st))

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """rollout_storage.py:100-139: ONE permutation reused by every epoch (:103,:119)."""
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        indices = torch.randperm(num_mini_batches * mini_batch_size, requires_grad=False, device=self.device)
        flat = lambda t: t.flatten(0, 1)
        obs, priv, hist = flat(self.observations), flat(self.privileged_observations), flat(self.observation_histories)
        actions, values, returns = flat(self.actions), flat(self.values), flat(self.returns)
        logp, adv, mu, sigma, bins = flat(self.actions_log_prob), flat(self.advantages), flat(self.mu), \
            flat(self.sigma), flat(self.env_bins)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                idx = indices[i * mini_batch_size:(i + 1) * mini_batch_size]
                yield obs[idx], obs[idx], priv[idx], hist[idx], actions[idx], values[idx], adv[idx], returns[idx], \
                    logp[idx], mu[idx], sigma[idx], None, bins[idx]
