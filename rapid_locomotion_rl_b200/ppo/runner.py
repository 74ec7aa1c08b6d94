"""Runner: the rollout / update loop that drives the hot path.

Mirror of mini_gym_learn/ppo/__init__.py:47-265 (`RunnerArgs`, `Runner(env, device)`, `learn(...)`): same
call order - `alg.act` -> `env.step` -> `alg.process_env_step` for `num_steps_per_env` steps, then
`alg.compute_returns` and `alg.update` - on the kernels of this package.  What is NOT here is the
reference's control plane (ml_logger uploads, video recording, TorchScript export, curriculum caches): a
`log` callback receives the per-iteration numbers instead.

SURVEY.md 8(f1): after the kernels are fast the Python between them is the bottleneck (24 steps x ~10
launches + dict handling).  With `graph_rollout=True` one whole rollout (24 x [policy chain, Normal
sample, fused env step, history push, 11 storage writes]) is captured ONCE into a CUDA graph and replayed
per iteration: no Python, no launch latency between kernels.  Everything inside is replay-safe: the env
kernel keys its RNG with a device-side step counter, the action noise comes from torch's graph-aware
Philox generator, and all tensors are the persistent buffers of the env / storage.
"""
import time

import torch

from .actor_critic import ActorCritic
from .ppo import PPO


class RunnerArgs:
    """mini_gym_learn/ppo/__init__.py:47-62."""
    algorithm_class_name = "PPO"
    num_steps_per_env = 24
    max_iterations = 1500
    save_interval = 400
    save_video_interval = 100
    log_freq = 10
    resume = False
    load_run = -1
    checkpoint = -1
    resume_path = None


class Runner:
    def __init__(self, env, device="cuda:0", graph_rollout=False, physics=None):
        """env: HistoryWrapper(VelocityTrackingEasyEnv(...)).  physics: optional callable run after every
        env.step (stands in for the simulator advancing its state tensors; synthetic in the tests)."""
        self.device = device
        self.env = env
        actor_critic = ActorCritic(env.num_obs, env.num_privileged_obs, env.num_obs_history, env.num_actions, device=device)
        self.alg = PPO(actor_critic, device=device)
        self.num_steps_per_env = RunnerArgs.num_steps_per_env
        self.alg.init_storage(env.num_train_envs, self.num_steps_per_env, [env.num_obs], [env.num_privileged_obs],
                              [env.num_obs_history], [env.num_actions])
        self.tot_timesteps = 0
        self.tot_time = 0
        self.current_learning_iteration = 0
        self.last_recording_it = 0
        self.graph_rollout = graph_rollout
        self.physics = physics
        self._graphs, self._warm = {}, False
        self.env.reset()

    # ------------------------------------------------------------------------------------------------
    def _rollout_steps(self, obs, privileged_obs, obs_history):
        """mini_gym_learn/ppo/__init__.py:127-141 for the training envs (the eval split is 8(f3) 'next')."""
        n = self.env.num_train_envs
        alg = self.alg
        for _ in range(self.num_steps_per_env):
            z = torch.randn(n, self.env.num_actions, device=self.device)         # graph-safe Philox stream
            alg.transition.actions = alg.actor_critic.act(obs[:n], privileged_obs[:n], inject_normal=z).detach()
            alg.transition.values = alg.actor_critic.evaluate(obs[:n], privileged_obs[:n]).detach()
            t = alg.transition
            t.actions_log_prob = alg.actor_critic.get_actions_log_prob(t.actions).detach()
            t.action_mean, t.action_sigma = alg.actor_critic._mu, alg.actor_critic._sigma
            t.observations, t.critic_observations = obs[:n], obs[:n]
            t.privileged_observations, t.observation_histories = privileged_obs[:n], obs_history[:n]
            obs_dict, rewards, dones, infos = self.env.step(t.actions)
            if self.physics is not None:
                self.physics(self.env)
            obs, privileged_obs, obs_history = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
            alg.process_env_step(rewards[:n], dones[:n], {"env_bins": infos["env_bins"]} if "env_bins" in infos else
                                 {"env_bins": torch.zeros(n, device=self.device)})
        return obs, privileged_obs, obs_history

    def _rollout(self, obs, privileged_obs, obs_history):
        if not self.graph_rollout:
            return self._rollout_steps(obs, privileged_obs, obs_history)
        T = self.num_steps_per_env
        inner = getattr(self.env, "env", self.env)
        if not self._warm:
            # first rollout runs eagerly: it compiles the policy chain and makes every lazy allocation
            self._warm = True
            inner.use_device_step_counter(True)
            return self._rollout_steps(obs, privileged_obs, obs_history)
        # The history ring advances T mod H slots per rollout, and the `obs_history` view (an address) and the
        # push slots are frozen inside a graph: one graph per distinct start slot (lcm(T, H) / T of them: 5 for
        # T = 24, H = 15), captured on first use, then replayed round robin.
        key = getattr(self.env, "_slot", 0)
        if key not in self._graphs:
            host = (key, inner.common_step_counter, self.alg.actor_critic._act_step)
            self.alg.actor_critic.prepare_rollout_chains(self.env.num_train_envs)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._rollout_steps(obs, privileged_obs, obs_history)
            slot_after = getattr(self.env, "_slot", 0)
            # capture ran the Python (host counters moved) but no kernel: rewind, the replay below does the work
            if hasattr(self.env, "_slot"):
                self.env._slot = host[0]
            inner.common_step_counter, self.alg.actor_critic._act_step = host[1], host[2]
            self.alg.storage.clear()
            self._graphs[key] = (g, out, slot_after)
        g, out, slot_after = self._graphs[key]
        g.replay()
        if hasattr(self.env, "_slot"):
            self.env._slot = slot_after
        inner.common_step_counter += T
        self.alg.storage.step = T
        return out

    # ------------------------------------------------------------------------------------------------
    def learn(self, num_learning_iterations, init_at_random_ep_len=False, eval_freq=100, eval_expert=False, log=None):
        """mini_gym_learn/ppo/__init__.py:92-265 without the logger / checkpoint plumbing.  Returns the list
        of per-iteration dicts that `log` (if given) also receives."""
        env = self.env
        if init_at_random_ep_len:
            env.episode_length_buf.copy_(torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length)))
        obs_dict = env.get_observations()
        obs, privileged_obs, obs_history = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
        self.alg.actor_critic.train()
        history = []
        tot_iter = self.current_learning_iteration + num_learning_iterations
        n = env.num_train_envs
        for it in range(self.current_learning_iteration, tot_iter):
            start = time.time()
            with torch.inference_mode():
                obs, privileged_obs, obs_history = self._rollout(obs, privileged_obs, obs_history)
                self.alg.compute_returns(obs[:n], privileged_obs[:n])
            mean_value_loss, mean_surrogate_loss, mean_adaptation_module_loss = self.alg.update()
            self.tot_timesteps += self.num_steps_per_env * env.num_envs
            rec = dict(iteration=it, time_iter=time.time() - start, adaptation_loss=mean_adaptation_module_loss,
                       mean_value_loss=mean_value_loss, mean_surrogate_loss=mean_surrogate_loss,
                       timesteps=self.tot_timesteps, learning_rate=self.alg.learning_rate)
            history.append(rec)
            if log is not None:
                log(rec)
        self.current_learning_iteration += num_learning_iterations
        return history

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        return self.alg.actor_critic.act_inference

    def get_expert_policy(self, device=None):
        self.alg.actor_critic.eval()
        return self.alg.actor_critic.act_expert
