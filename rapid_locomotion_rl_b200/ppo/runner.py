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


def export_policy(actor_critic, directory, iteration=0):
    """The files the reference's Runner writes every `save_interval` iterations and at the end of learn()
    (mini_gym_learn/ppo/__init__.py:222-242, :248-265), under `directory`/checkpoints/:
        ac_weights_{iteration:06d}.pt, ac_weights_last.pt   the state_dict (35 keys, loadable by the reference class)
        adaptation_module_latest.jit, body_latest.jit       TorchScript of the adaptation module and the actor body on
                                                             the CPU - what scripts/play.py and the robot deploy.
    (The reference uploads them through ml_logger; here they are plain files.)  Returns the paths."""
    import os

    import torch.nn as nn
    ck = os.path.join(directory, "checkpoints")
    os.makedirs(ck, exist_ok=True)
    sd = {k: v.detach().cpu().clone() for k, v in actor_critic.state_dict().items()}
    paths = [os.path.join(ck, "ac_weights_%06d.pt" % iteration), os.path.join(ck, "ac_weights_last.pt")]
    for p_ in paths:
        torch.save(sd, p_)

    def fresh(prefix, seq):
        # an ordinary fp32 module with its own storage (the live parameters are views of one flat device buffer)
        layers = []
        for m in seq:
            if isinstance(m, nn.Linear):
                width = sd[prefix + ".%d.weight" % len(layers)].shape[1]       # (no-latent family: first layer [., num_obs])
                layers.append(nn.Linear(width, m.out_features))
            else:
                layers.append(nn.ELU() if actor_critic.activation == "elu" else nn.Tanh())
        mod = nn.Sequential(*layers)
        mod.load_state_dict({k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")})
        return mod.eval()
    exports = [("body_latest.jit", "actor_body", actor_critic.actor_body)]
    if actor_critic.use_latent:        # (high_level_policy/ppo/__init__.py:237-242: only with USE_LATENT)
        exports.insert(0, ("adaptation_module_latest.jit", "adaptation_module", actor_critic.adaptation_module))
    for name, prefix, seq in exports:
        path = os.path.join(ck, name)
        torch.jit.script(fresh(prefix, seq)).save(path)
        paths.append(path)
    return paths


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
    # the learner family (overridden by rapid_locomotion_rl_b200.high_level_policy.ppo.Runner)
    runner_args = RunnerArgs
    actor_critic_class = ActorCritic
    ppo_class = PPO

    def __init__(self, env, device="cuda:0", graph_rollout=False, physics=None, fused_rollout=True):
        """env: HistoryWrapper(VelocityTrackingEasyEnv(...)).  physics: optional callable run after every
        env.step (stands in for the simulator advancing its state tensors; synthetic in the tests)."""
        self.device = device
        self.env = env
        actor_critic = self.actor_critic_class(env.num_obs, env.num_privileged_obs, env.num_obs_history, env.num_actions,
                                               device=device)
        self.alg = self.ppo_class(actor_critic, device=device)
        self.num_steps_per_env = self.runner_args.num_steps_per_env
        self.alg.init_storage(env.num_train_envs, self.num_steps_per_env, [env.num_obs], [env.num_privileged_obs],
                              [env.num_obs_history], [env.num_actions])
        self.tot_timesteps = 0
        self.tot_time = 0
        self.current_learning_iteration = 0
        self.last_recording_it = 0
        self.graph_rollout = graph_rollout
        self.physics = physics
        self._graphs, self._warm, self._graphs_ws_gen = {}, False, -1
        self.fused_rollout = fused_rollout
        self._act_buf = self._step_state = self._zero_bins = None
        self._fused_device_steps = False
        # test hooks (eager rollouts): standard-normal draws to use instead of the generator ([n, 12] tensor), and a
        # callable(t, action_buffer, storage_slice_or_None) that may overwrite the actions between the policy and env.step
        self.inject_normal = None
        self.action_hook = None
        self.eval_expert = False          # learn(eval_expert=...): evaluation envs act with the teacher instead of the student
        self.env.reset()

    # ------------------------------------------------------------------------------------------------
    def _rollout_steps(self, obs, privileged_obs, obs_history):
        """mini_gym_learn/ppo/__init__.py:127-141: the training envs act through PPO.act and fill the storage; evaluation envs
        (train / eval split) act with the student - or, `eval_expert`, the teacher - policy and are only stepped."""
        n, n_all = self.env.num_train_envs, self.env.num_envs
        alg = self.alg
        for _ in range(self.num_steps_per_env):
            actions_eval = None
            if n_all > n:
                # (:130-135; evaluated first and copied: the training pass below re-uses the learner's workspace)
                ac = alg.actor_critic
                actions_eval = (ac.act_teacher(obs[n:], privileged_obs[n:]) if self.eval_expert
                                else ac.act_student(obs[n:], obs_history[n:])).detach().clone()
            z = self.inject_normal if self.inject_normal is not None else \
                torch.randn(n, self.env.num_actions, device=self.device)         # graph-safe Philox stream
            alg.transition.actions = alg.actor_critic.act(obs[:n], privileged_obs[:n], inject_normal=z).detach()
            alg.transition.values = alg.actor_critic.evaluate(obs[:n], privileged_obs[:n]).detach()
            t = alg.transition
            t.actions_log_prob = alg.actor_critic.get_actions_log_prob(t.actions).detach()
            t.action_mean, t.action_sigma = alg.actor_critic._mu, alg.actor_critic._sigma
            # (stored now: env.step rewrites these buffers in place - see PPO.act)
            if self.action_hook is not None:
                self.action_hook(alg.storage.step, t.actions, None)
            alg.storage.store_observations(obs[:n], privileged_obs[:n], obs_history[:n])
            t.observations = t.critic_observations = t.privileged_observations = t.observation_histories = None
            obs_dict, rewards, dones, infos = self.env.step(t.actions if actions_eval is None else
                                                            torch.cat((t.actions, actions_eval), dim=0))      # :136
            if self.physics is not None:
                self.physics(self.env)
            obs, privileged_obs, obs_history = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
            # the reference passes `infos` through (time_outs feeds the gamma * V bootstrap, ppo.py:81-83); only a
            # missing env_bins is defaulted
            step_infos = {"env_bins": infos["env_bins"][:n] if "env_bins" in infos else torch.zeros(n, device=self.device)}
            if "time_outs" in infos:
                step_infos["time_outs"] = infos["time_outs"][:n]
            alg.process_env_step(rewards[:n], dones[:n], step_infos)
        return obs, privileged_obs, obs_history

    # ------------------------------------------------------------------------------------------------
    def _fused_ok(self):
        """The fused step glue (csrc/rollout.cu) takes the standard stack: HistoryWrapper over a LeggedRobot whose
        observation row fits a warp pair, the chain learner, no eval split."""
        env, ac = self.env, self.alg.actor_critic
        inner = getattr(env, "env", None)
        return (self.fused_rollout and inner is not None and hasattr(env, "_ring") and hasattr(inner, "_reset_u8") and
                ac.use_chain and env.num_obs <= 64 and env.num_privileged_obs <= 32 and env.num_actions == 12 and
                env.num_train_envs == env.num_envs and ac.use_latent)

    def _rollout_steps_fused(self, obs, privileged_obs, obs_history):
        """The same loop as `_rollout_steps` with the per-step glue fused: per step ONE boundary kernel (closes
        transition t-1: reward / bootstrap / done / bin + history push; opens transition t: obs / priv / history rows
        into the storage + bf16 staging of the policy inputs), the policy pass, ONE act kernel (Normal sample,
        log-prob, action / mu / sigma / log-prob / value slices) and the fused env step: 4-5 launches instead of ~25.
        What lands in the storage is what `_rollout_steps` stores (tests/test_runner_gpu.py)."""
        import ctypes as C
        from .. import _lib
        from .ppo import PPO_Args
        env, alg = self.env, self.alg
        inner, ac, st = env.env, alg.actor_critic, alg.storage
        lib, P = _lib.lib(), _lib.ptr
        n, T, H, W = env.num_train_envs, self.num_steps_per_env, env.obs_history_length, env.num_obs
        w = ac.workspace(n)
        if self._act_buf is None or self._act_buf.shape[0] != n:
            self._act_buf = torch.zeros(n, env.num_actions, device=self.device)
            self._step_state = torch.zeros(2, dtype=torch.int64, device=self.device)
            self._zero_bins = torch.zeros(n, device=self.device)
        if st.step != 0:
            raise AssertionError("Rollout buffer overflow")

        def boundary(t_close, t_open):
            q = _lib.RlRolloutBoundary()
            q.obs, q.priv, q.ring = P(inner.obs_buf), P(inner.privileged_obs_buf), P(env._ring)
            q.N, q.obs_dim, q.priv_dim, q.H = n, W, env.num_privileged_obs, H
            q.gamma = PPO_Args.gamma
            q.do_post, q.do_pre = int(t_close is not None), int(t_open is not None)
            if t_close is not None:
                env._slot = (env._slot + 1) % H
                q.push_slot = env._slot
                ex = inner.extras
                tmo = ex["time_outs"] if "time_outs" in ex else None
                bins = ex["env_bins"] if "env_bins" in ex else None
                q.rew, q.dones = P(inner.rew_buf), P(inner._reset_u8)
                q.time_outs = P(tmo.view(torch.uint8)) if tmo is not None else None
                q.values_prev = P(st.values[t_close])
                q.bins = P(bins) if bins is not None else None
                q.dst_rewards, q.dst_dones, q.dst_bins = P(st.rewards[t_close]), P(st.dones[t_close]), P(st.env_bins[t_close])
                self._keep = (tmo, bins)
            if t_open is not None:
                q.hist_slot = env._slot + 1
                q.dst_obs, q.dst_priv, q.dst_hist = P(st.observations[t_open]), P(st.privileged_observations[t_open]), \
                    P(st.observation_histories[t_open])
                q.Xac, q.ld_xac, q.Xp, q.ld_xp = P(w["Xac"]), w["Xac"].shape[1], P(w["Xp"]), w["Xp"].shape[1]
            _lib.check(lib.rl_rollout_boundary(C.byref(q), _lib.current_stream()))

        for t in range(T):
            boundary(t - 1 if t > 0 else None, t)
            ac.forward_teacher(n)                      # encoder -> actor mean / critic value on the staged rows
            ac._cache_key = None
            a = _lib.RlRolloutAct()
            a.mean, a.value, a.std = P(w["mean"]), P(w["value"]), P(ac.std.data)
            a.inj_normal, a.actions_out, a.logp_out = P(self.inject_normal), P(self._act_buf), None
            a.dst_actions, a.dst_mu, a.dst_sigma = P(st.actions[t]), P(st.mu[t]), P(st.sigma[t])
            a.dst_logp, a.dst_values = P(st.actions_log_prob[t]), P(st.values[t])
            ac._act_step += 1
            a.step_state = P(self._step_state) if self._fused_device_steps else None
            a.seed, a.step, a.N = ac.seed, (0 if self._fused_device_steps else ac._act_step), n
            _lib.check(lib.rl_rollout_act(C.byref(a), _lib.current_stream()))
            if self.action_hook is not None:
                self.action_hook(t, self._act_buf, st.actions[t])
            inner.step(self._act_buf)
            if self.physics is not None:
                self.physics(env)
        boundary(T - 1, None)
        st.step = T
        return inner.obs_buf, inner.privileged_obs_buf, env.obs_history

    def _rollout(self, obs, privileged_obs, obs_history):
        steps = self._rollout_steps_fused if self._fused_ok() else self._rollout_steps
        if not self.graph_rollout:
            return steps(obs, privileged_obs, obs_history)
        T = self.num_steps_per_env
        inner = getattr(self.env, "env", self.env)
        if not self._warm:
            # first rollout runs eagerly: it compiles the policy chain and makes every lazy allocation
            self._warm = True
            inner.use_device_step_counter(True)
            out = steps(obs, privileged_obs, obs_history)
            if self._fused_ok():
                # from here on the act kernel keys its Philox draws with a device-side counter (graph replay freezes
                # kernel arguments); it continues from the host count
                self._fused_device_steps = True
                self._step_state[0] = self.alg.actor_critic._act_step + 1
            return out
        # The history ring advances T mod H slots per rollout, and the `obs_history` view (an address) and the
        # push slots are frozen inside a graph: one graph per distinct start slot (lcm(T, H) / T of them: 5 for
        # T = 24, H = 15), captured on first use, then replayed round robin.
        key = getattr(self.env, "_slot", 0)
        ws_gen = getattr(self.alg.actor_critic, "_ws_gen", 0)
        if ws_gen != self._graphs_ws_gen:        # the learner's workspace moved: the captured addresses are stale
            self._graphs, self._graphs_ws_gen = {}, ws_gen
        if key not in self._graphs:
            host = (key, inner.common_step_counter, self.alg.actor_critic._act_step)
            self.alg.actor_critic.prepare_rollout_chains(self.env.num_train_envs)
            self._graphs_ws_gen = getattr(self.alg.actor_critic, "_ws_gen", 0)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = steps(obs, privileged_obs, obs_history)
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
    def learn(self, num_learning_iterations, init_at_random_ep_len=False, eval_freq=100, eval_expert=False, log=None,
              save_dir=None):
        """mini_gym_learn/ppo/__init__.py:92-265 without the logger plumbing.  Returns the list of per-iteration dicts
        that `log` (if given) also receives.  save_dir: write the reference's checkpoint / TorchScript files there every
        RunnerArgs.save_interval iterations and at the end (export_policy)."""
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
                self.eval_expert = bool(eval_expert)
                obs, privileged_obs, obs_history = self._rollout(obs, privileged_obs, obs_history)
                if it % eval_freq == 0:
                    self.env.reset_evaluation_envs()                     # :194-195 (no-op without evaluation envs)
                self.alg.compute_returns(obs[:n], privileged_obs[:n])
            mean_value_loss, mean_surrogate_loss, mean_adaptation_module_loss = self.alg.update()
            self.tot_timesteps += self.num_steps_per_env * env.num_envs
            rec = dict(iteration=it, time_iter=time.time() - start, adaptation_loss=mean_adaptation_module_loss,
                       mean_value_loss=mean_value_loss, mean_surrogate_loss=mean_surrogate_loss,
                       timesteps=self.tot_timesteps, learning_rate=self.alg.learning_rate)
            history.append(rec)
            if log is not None:
                log(rec)
            if save_dir is not None and it % self.runner_args.save_interval == 0:
                export_policy(self.alg.actor_critic, save_dir, it)
        self.current_learning_iteration += num_learning_iterations
        if save_dir is not None and num_learning_iterations > 0:
            export_policy(self.alg.actor_critic, save_dir, it)
        return history

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        return self.alg.actor_critic.act_inference

    def get_expert_policy(self, device=None):
        self.alg.actor_critic.eval()
        return self.alg.actor_critic.act_expert
