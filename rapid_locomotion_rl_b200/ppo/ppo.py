"""PPO on the fused kernels.

Drop-in for mini_gym_learn/ppo/ppo.py:15-178: `PPO_Args`, `PPO(actor_critic, device)`, `init_storage`,
`act`, `process_env_step`, `compute_returns`, `update() -> (mean_value_loss, mean_surrogate_loss,
mean_adaptation_module_loss)`.

One minibatch step = gather (1 launch) -> 14 forward GEMMs -> fused loss/gradient kernel -> 25 backward
GEMMs (dgrad with the ELU derivative in the epilogue, split-K wgrad with the bias gradient from a
ones-MMA) -> [NCCL all-reduce of the flat gradient when torch.distributed is initialised] -> gradient
norm / clip / KL-adaptive learning rate on the device -> fused Adam (policy) -> fused Adam (adaptation
module) -> bf16 shadow refresh.  No host synchronisation happens inside update(); the three returned
means are read back once at the end (the reference syncs >= 5 times per minibatch).
"""
import ctypes as C
import os

import torch

from .. import _lib
from . import chain
from .actor_critic import (AC_Args, ActorCritic, EPI_ATOMIC, EPI_BF16, EPI_DELU_BF16)
from .rollout_storage import RolloutStorage


class PPO_Args:
    """ppo.py:15-30."""
    value_loss_coef = 1.0
    use_clipped_value_loss = True
    clip_param = 0.2
    entropy_coef = 0.01
    num_learning_epochs = 5
    num_mini_batches = 4
    learning_rate = 1.e-3
    adaptation_module_learning_rate = 1.e-3
    num_adaptation_module_substeps = 1
    schedule = "adaptive"
    gamma = 0.99
    lam = 0.95
    desired_kl = 0.01
    max_grad_norm = 1.


class PPO:
    actor_critic: ActorCritic

    def __init__(self, actor_critic, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RlError("PPO needs a CUDA device: the learner has no CPU fallback")
        self._lib = _lib.lib()
        self.actor_critic = actor_critic
        self.storage = None
        self.transition = RolloutStorage.Transition()
        self.learning_rate = PPO_Args.learning_rate
        dev = self.device
        self._ctrl = torch.tensor([PPO_Args.learning_rate, 1.0, 0.0], device=dev)      # lr, clip coef, kl
        self._stats = torch.zeros(4, dtype=torch.float64, device=dev)
        self._fin_ws = torch.zeros(2, dtype=torch.float64, device=dev)
        self._loss_acc = torch.zeros(4, dtype=torch.float64, device=dev)
        # The adaptation module's forward pass depends only on the gathered history rows and its own weights, not
        # on the policy update: it runs on a side stream (a parallel branch of the captured graph) next to the
        # teacher path, where its CTAs fill the SMs that a 188-tile minibatch leaves idle in its second wave.
        self.overlap_adaptation = os.environ.get("RL_PPO_OVERLAP", "1") != "0"
        self._side = self._ev_fork = self._ev_join = self._ev_grads = self._ev_loss = self._hp = None
        self._graph_lag = self._graph_rest = self._graph_reduce = None
        self._graph_ws_gen = -1
        self._chunk_stream = self._ev_chunk_fork = self._ev_chunk_join = None
        self._fused_adam = os.environ.get("RL_PPO_FUSED_ADAM", "1") != "0"      # Adam + bf16 operands in one launch
        self._ada_g = None          # several GPUs over NCCL: the adaptation gradient in a buffer of its own (enable_side_comm)
        self._steps = torch.zeros(4, dtype=torch.int32, device=dev)    # {main count, ticket, adaptation count, ticket}
        self._graph = None          # CUDA graph of one minibatch step (single-GPU path)
        self._graph_B = 0
        self._idx_buf = None
        self.use_cuda_graph = True
        self.group_wgrads = os.environ.get("RL_GROUP_WGRADS", "1") != "0"
        self._peer = None           # sharding.PeerComm once enable_peer_allreduce() has mapped the ranks' buffers
        self._g_red = None
        self._wq = []
        _lib.check(self._lib.rl_gemm_init())
        ac = actor_critic
        self._main_layers = ac.L_enc + [ac.L_cat] + ac.L_act + ac.L_cri

    def enable_side_comm(self):
        """Several GPUs, NCCL (opt-in, RL_PPO_SIDE_COMM=1): the adaptation module's gradient gets a buffer of its own and
        is summed over a SECOND communicator from the side branch, concurrently with the policy path - so that the one
        all-reduce per minibatch need not wait for the side branch (adaptation forward / loss / dgrad / wgrad of the
        previous minibatch, running at low priority in whatever the policy kernels leave free).  Identical results
        (tests/multigpu_worker.py), but measured slower than the single collective: see update().  Collective (creates
        the group and warms its communicator up outside any graph capture)."""
        from ..sharding import side_all_reduce_sum_
        if self._ada_g is None:
            assert self._peer is None, "the peer all-reduce keeps the gradient in one mapped buffer"
            self._ada_g = self.actor_critic.split_adaptation_grad()
            warm = torch.zeros(8, device=self.device)
            side_all_reduce_sum_(warm)
            torch.cuda.synchronize()

    def enable_peer_allreduce(self):
        """Multi-GPU: moves the gradient accumulation buffer into memory every rank has mapped and switches the
        gradient all-reduce of update() from NCCL to the fused NVLink peer kernel.  Collective: call on every rank."""
        from ..sharding import PeerComm
        ac = self.actor_critic
        self._peer = PeerComm(ac.n_total + 8, self.device)
        ac.rebind_grad_store(self._peer.grad)
        self._g_red = torch.zeros(self._peer.n, device=self.device)
        self._graph = None

    def release_graphs(self):
        """Drops the captured CUDA graphs of update().  With several GPUs the graphs hold NCCL kernels: call this (or
        let the PPO object go out of scope) BEFORE torch.distributed.destroy_process_group(), which otherwise waits
        for the communicator the graphs still reference."""
        self._graph = self._graph_rest = None
        self._graph_B, self._graph_lag = 0, None

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, privileged_obs_shape, obs_history_shape,
                     action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, privileged_obs_shape,
                                      obs_history_shape, action_shape, self.device)

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    def act(self, obs, privileged_obs, obs_history):
        """ppo.py:62-74."""
        ac, t = self.actor_critic, self.transition
        t.actions = ac.act(obs, privileged_obs).detach()
        t.values = ac.evaluate(obs, privileged_obs).detach()
        t.actions_log_prob = ac.get_actions_log_prob(t.actions).detach()
        t.action_mean = ac._mu
        t.action_sigma = ac._sigma
        # "need to record obs and critic_obs before env.step()" (ppo.py:69): the reference's env returns NEW tensors every
        # step, so holding references is enough there.  This env rewrites obs_buf / privileged_obs_buf / the history ring
        # in place, so the three rows go into their [step] slices of the storage NOW; add_transitions finds them stored.
        if self.storage is not None:
            self.storage.store_observations(obs, privileged_obs, obs_history)
            t.observations = t.critic_observations = t.privileged_observations = t.observation_histories = None
        else:
            t.observations = t.critic_observations = obs
            t.privileged_observations, t.observation_histories = privileged_obs, obs_history
        return t.actions

    def process_env_step(self, rewards, dones, infos):
        """ppo.py:76-88."""
        t = self.transition
        self.actor_critic._cache_key = None      # the env buffers behind the cached pass have been rewritten
        t.rewards = rewards.clone()
        t.dones = dones
        # (the high_level_policy learner stores no bins: high_level_policy/ppo/ppo.py:81)
        t.env_bins = infos["env_bins"] if "env_bins" in infos else torch.zeros_like(t.rewards)
        if "time_outs" in infos:
            t.rewards += PPO_Args.gamma * torch.squeeze(t.values * infos["time_outs"].unsqueeze(1).to(self.device), 1)
        self.storage.add_transitions(t)
        t.clear()
        self.actor_critic.reset(dones)

    def compute_returns(self, last_critic_obs, last_critic_privileged_obs):
        last_values = self.actor_critic.evaluate(last_critic_obs, last_critic_privileged_obs).detach()
        self.storage.compute_returns(last_values, PPO_Args.gamma, PPO_Args.lam)

    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _split(M, N, bn_threshold, rows):
        tiles = ((M + 127) // 128) * ((N + (127 if N > 64 else 63)) // (128 if N > 64 else 64))
        total_kb = (rows + 63) // 64
        return max(1, min(total_kb, (296 + tiles - 1) // tiles))

    def _wgrad(self, L, dY, dy_off, ld_dy, X, x_off, ld_x, rows):
        """dW[out,in] += dY^T X (split-K, atomics into the flat gradient) and db from the ones-MMA.
        Problems are queued and launched together by _wgrad_flush (one grid for all layers)."""
        ac = self.actor_critic
        if self.group_wgrads:
            q = _lib.RlWgradProblem()
            q.dY, q.X, q.dW, q.db = ac._p(dY, dy_off), ac._p(X, x_off), L.gw.data_ptr(), L.gb.data_ptr()
            q.M, q.N, q.K, q.ld_dy, q.ld_x, q.ld_dw = L.out, L.inp, rows, ld_dy, ld_x, L.inp
            self._wq.append(q)
            return
        split = self._split(L.out, L.inp, 64, rows)
        ac._gemm(ac._p(dY, dy_off), ld_dy, ac._p(X, x_off), ld_x, L.gw.data_ptr(), L.inp, L.out, L.inp, rows,
                 EPI_ATOMIC, transposed=1, db=L.gb.data_ptr(), split_k=split)

    def _wgrad_flush(self):
        """Launches the queued weight-gradient problems as one grouped grid (csrc/gemm_tc.cu), largest first."""
        if not self._wq:
            return
        qs = sorted(self._wq, key=lambda q: -(q.M * q.N))
        self._wq = []
        if os.environ.get("RL_WGRAD_PERSISTENT", "1") != "0":
            # persistent kernel (csrc/wgrad_persistent.cu): the library sizes the work items itself
            for q in qs:
                q.split_k = 0
            arr = (_lib.RlWgradProblem * len(qs))(*qs)
            _lib.check(self._lib.rl_wgrad_grouped(arr, len(qs), _lib.current_stream()))
            return
        tiles = sum(((q.M + 127) // 128) * ((q.N + (127 if q.N > 64 else 63)) // (128 if q.N > 64 else 64)) for q in qs)
        total_kb = (qs[0].K + 63) // 64
        # every CTA reduces the same number of 64-row k-blocks (equal durations, no straggler problem):
        # about six waves of CTAs over the 148 SMs, at most 64 k-blocks each
        kb_per_cta = int(max(4, min(64, total_kb * tiles / (148.0 * 6))))
        for q in qs:
            q.split_k = max(1, (total_kb + kb_per_cta - 1) // kb_per_cta)
        arr = (_lib.RlWgradProblem * len(qs))(*qs)
        _lib.check(self._lib.rl_wgrad_grouped(arr, len(qs), _lib.current_stream()))

    def _dgrad(self, L, dY, dy_off, ld_dy, dX, dx_off, ld_dx, rows, aux=None, aux_off=0, ld_aux=0, col0=0, ncols=None):
        """dX = (dY W[:, col0:col0+ncols]) (* elu'(aux)): B operand = the transposed shadow, rows col0.."""
        ac = self.actor_critic
        ncols = L.inp if ncols is None else ncols
        ac._gemm(ac._p(dY, dy_off), ld_dy, ac._p(L.wbt, col0 * L.ld_wbt), L.ld_wbt, ac._p(dX, dx_off), ld_dx, rows, ncols,
                 L.out, EPI_DELU_BF16 if aux is not None else EPI_BF16,
                 aux=None if aux is None else ac._p(aux, aux_off), ld_aux=ld_aux)

    def minibatch_step(self, idx, world=1, allreduce=None, lag=False, pending=False):
        """One PPO minibatch on the rows `idx` (int64 device tensor) of the flattened storage.

        lag=False: the reference's order inside one call (ppo.py:99-170): policy forward / loss / backward /
        [all-reduce] / clip + KL-adaptive lr + Adam, then the adaptation module's regression step against the
        latent of the UPDATED encoder [+ its own all-reduce].
        lag=True (update()'s CUDA-graph schedule): the adaptation module runs ONE CALL BEHIND on a side stream.
        Call i's side branch does the adaptation forward / loss / dgrad / wgrad of minibatch i-1 (`pending`; its
        history rows were gathered by call i-1, its regression target - the refreshed encoder latent - was copied out
        of the [obs | latent] box by call i-1) and then gathers the history rows of minibatch i; the main branch joins
        it before the ONE all-reduce of [policy gradient | adaptation gradient | KL sum], then steps both optimisers.
        Every dependent pair of operations keeps the reference's order (the policy path never reads the adaptation
        module; the adaptation forward of minibatch i-1 sees the weights left by the step of minibatch i-2), so the
        parameters are those of the serial schedule.  update() flushes the last pending minibatch."""
        ac, st, A = self.actor_critic, self.storage, PPO_Args
        B = int(idx.numel())
        w = ac.workspace(B, backward=True)
        P = _lib.ptr
        ld = lambda k: w[k].shape[1]
        flat = lambda t: t.flatten(0, 1)
        assert not lag or (ac.use_chain and A.num_adaptation_module_substeps == 1)
        debug = getattr(self, "debug_keep_grad", False)
        acc0 = self._loss_acc.clone() if debug else None
        peer = self._peer if allreduce == "peer" else None
        g_used = self._g_red if peer is not None else ac.flat_grad
        if lag:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
                self._ev_fork, self._ev_join, self._ev_grads, self._ev_loss = (torch.cuda.Event() for _ in range(4))
            # The adaptation FORWARD of minibatch i runs at the END of call i's side branch (behind the history gather that
            # produces its input and the module's own optimiser step of minibatch i-1), not at the start of call i+1's: there
            # it overlaps the main path's finalize / Adam / encoder-refresh tail, during which the side branch used to idle,
            # instead of holding every SM (a 148-CTA chain launch) when the next call's policy forward wants them.  Only where
            # the module's optimiser step is on the side branch too (one GPU, or the side communicator), and only for batches of
            # several waves of row tiles: A/B in one gpurun call (profiles/jobs/r2_job63.sh) 32768 envs (1536 tiles) 38.02 ->
            # 37.34 ms per update, 4000 envs (188 tiles) 7.30 -> 7.44 - there the extra 148-CTA launch lands on the ragged
            # chunk's backward (timelines gpurun_out/r2_ppo_timeline_adafwd0/1.txt).  RL_PPO_ADA_FWD_EARLY=0 / 1 forces it.
            knob = os.environ.get("RL_PPO_ADA_FWD_EARLY", "auto")
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            self._ada_pre = pre = (allreduce is None or self._ada_g is not None) and \
                (knob == "1" or (knob != "0" and (B + 127) // 128 > 3 * sms))
            self._ev_fork.record()
            with torch.cuda.stream(self._side):
                self._side.wait_event(self._ev_fork)
                if pending:
                    self._adapt_grads(B, world, lagged=True, loss_event=self._ev_loss, forward=not pre)
                    if allreduce is None or self._ada_g is not None:
                        # the adaptation module's optimiser step stays on the side branch too (nothing on the policy path
                        # reads its weights, gradient or step counter); several GPUs: its gradient is summed here, over
                        # the side communicator (enable_side_comm)
                        if allreduce is not None:
                            from ..sharding import side_all_reduce_sum_
                            side_all_reduce_sum_(self._ada_g)
                        self._adapt_step(g_used)
                    self._ev_grads.record()
                _lib.check(self._lib.rl_ppo_gather_history(P(flat(st.observation_histories)), P(idx), B, ac.num_hist, P(w["Xh"]),
                                                           ld("Xh"), _lib.current_stream()))
                if pre:
                    ac.forward_adaptation(B, save=True)
                self._ev_join.record()
        stream = _lib.current_stream()
        _lib.check(self._lib.rl_ppo_gather(
            P(flat(st.observations)), P(flat(st.privileged_observations)), P(flat(st.observation_histories)),
            P(flat(st.actions)), P(flat(st.values)), P(flat(st.returns)), P(flat(st.actions_log_prob)),
            P(flat(st.advantages)), P(flat(st.mu)), P(flat(st.sigma)), P(idx), B, ac.num_obs, ac.num_priv, ac.num_hist,
            P(w["Xp"]), ld("Xp"), P(w["Xac"]), ld("Xac"), None if lag else P(w["Xh"]), ld("Xh"), P(w["Lrow"]), stream))
        # ---- forward -> loss + output gradients -> dgrad.  Each 128-row tile's three passes depend only on that tile,
        # and a tile's pass is a serial chain of layers (~40 us whatever else runs): a batch that fills the SMs 1.27
        # times (24000 rows = 188 tiles on 148 SMs) would run every pass as two waves, the second 27 % full.  The tiles
        # of the ragged last wave therefore form a second CHUNK that goes through the same three launches on its own
        # stream: its forward runs next to the first chunk's backward instead of after the forward's first wave. ----
        inv_gb = 1.0 / (B * world)
        kl_slot = ac._grad_store[ac.n_total:ac.n_total + 1] if allreduce is not None else None
        H = AC_Args.actor_hidden_dims[0]
        a, c, e = ac.L_act, ac.L_cri, ac.L_enc

        # RL_PPO_FUSED_LOSS=1: the loss rides in the forward chain's last epilogue op (rl_chain_set_ppo_loss) - one launch
        # and one kernel boundary less between the forward and the backward of a chunk.  Bit-identical output gradients
        # (tests/test_ppo_gpu.py::test_fused_loss_equals_loss_kernel), but measured SLOWER inside the update's graph (4000
        # envs: 7.53 - 7.57 vs 7.34 ms per update, A/B in one gpurun call, profiles/jobs/r2_job54.sh): the loss launch was a
        # scheduling point at which the ragged chunk's forward got its SMs; without it the first chunk's backward takes
        # every SM the moment its forward ends and the ragged chunk finishes ~25 us later (profiles/r02_ppo_timeline_fused.txt).
        # Kept as an opt-in.
        fused_loss = None
        if ac.use_chain and os.environ.get("RL_PPO_FUSED_LOSS", "0") == "1" and ac.activation == "elu":
            fl = fused_loss = _lib.RlChainPpoLoss()
            fl.Lrow, fl.std, fl.dmean, fl.dvalue = P(w["Lrow"]), P(ac.std.data), P(w["dmean"]), P(w["dvalue"])
            fl.dstd, fl.stats, fl.kl_slot = P(ac.std_grad), P(self._stats), P(kl_slot)
            fl.clip, fl.value_coef, fl.entropy_coef, fl.inv_global_B = A.clip_param, A.value_loss_coef, A.entropy_coef, inv_gb
            fl.use_clipped_value, fl.mean_out, fl.value_out = int(A.use_clipped_value_loss), 0, 1

        def passes(t0, t1):
            r0, r1 = 128 * t0, min(B, 128 * t1)
            ac.forward_teacher(B, save=True, tiles=(t0, t1), loss=fused_loss)
            # (the statistics are zeroed by the previous call's finalize kernel)
            if fused_loss is None:
                _lib.check(self._lib.rl_ppo_loss(
                    P(w["mean"][r0:]), P(w["value"][r0:]), None, P(w["Xac"][r0:]), ld("Xac"), ac.num_obs, P(w["Lrow"][r0:]),
                    P(ac.std.data), r1 - r0, A.clip_param, A.value_loss_coef, A.entropy_coef, int(A.use_clipped_value_loss),
                    inv_gb, P(w["dmean"][r0:]), P(w["dvalue"][r0:]), P(w["dpred"][r0:]), P(ac.std_grad), P(self._stats),
                    P(kl_slot), _lib.current_stream()))
            # dgrad of actor + critic + encoder in ONE persistent kernel (csrc/chain.cu)
            ac._chain(("trunk_backward",), chain.trunk_backward).run(B, tiles=(t0, t1))
        chunks = self._chunks(B) if ac.use_chain else None
        if chunks:
            if self._chunk_stream is None:
                # HIGHER priority than the main path: when the first chunk's forward ends, the small second chunk must
                # get its SMs before the first chunk's (SM-filling) backward takes them all - measured the other way
                # round the second chunk simply queues behind everything (8.27 ms instead of 7.68 per update)
                self._chunk_stream = torch.cuda.Stream(device=self.device, priority=-4)
                self._ev_chunk_fork, self._ev_chunk_join = torch.cuda.Event(), torch.cuda.Event()
            self._ev_chunk_fork.record()
            with torch.cuda.stream(self._chunk_stream):
                self._chunk_stream.wait_event(self._ev_chunk_fork)
                passes(*chunks[1])
                self._ev_chunk_join.record()
            passes(*chunks[0])
            torch.cuda.current_stream().wait_event(self._ev_chunk_join)
        elif ac.use_chain:
            passes(0, (B + 127) // 128)
        else:
            ac.forward_teacher(B, save=True)
            _lib.check(self._lib.rl_ppo_loss(
                P(w["mean"]), P(w["value"]), None, P(w["Xac"]), ld("Xac"), ac.num_obs, P(w["Lrow"]), P(ac.std.data), B,
                A.clip_param, A.value_loss_coef, A.entropy_coef, int(A.use_clipped_value_loss), inv_gb, P(w["dmean"]),
                P(w["dvalue"]), P(w["dpred"]), P(ac.std_grad), P(self._stats), P(kl_slot), stream))
        if not ac.use_chain:
            self._dgrad(a[2], w["dmean"], 0, 16, w["dA3"], 0, ld("dA3"), B, aux=w["A3"], ld_aux=ld("A3"))
            self._dgrad(a[1], w["dA3"], 0, ld("dA3"), w["dA2"], 0, ld("dA2"), B, aux=w["A2"], ld_aux=ld("A2"))
            self._dgrad(a[0], w["dA2"], 0, ld("dA2"), w["dY1"], 0, ld("dY1"), B, aux=w["Y1"], ld_aux=ld("Y1"))
            self._dgrad(c[2], w["dvalue"], 0, 8, w["dC3"], 0, ld("dC3"), B, aux=w["C3"], ld_aux=ld("C3"))
            self._dgrad(c[1], w["dC3"], 0, ld("dC3"), w["dC2"], 0, ld("dC2"), B, aux=w["C2"], ld_aux=ld("C2"))
            self._dgrad(c[0], w["dC2"], 0, ld("dC2"), w["dY1"], H, ld("dY1"), B, aux=w["Y1"], aux_off=H, ld_aux=ld("Y1"))
            self._dgrad(ac.L_cat, w["dY1"], 0, ld("dY1"), w["dLat"], 0, ld("dLat"), B, col0=ac.num_obs, ncols=ac.latent_dim)
            self._dgrad(e[2], w["dLat"], 0, ld("dLat"), w["dH2"], 0, ld("dH2"), B, aux=w["H2"], ld_aux=ld("H2"))
            self._dgrad(e[1], w["dH2"], 0, ld("dH2"), w["dH1"], 0, ld("dH1"), B, aux=w["H1"], ld_aux=ld("H1"))
        # weight / bias gradients: split-K TN GEMMs over the batch rows
        self._wgrad(a[2], w["dmean"], 0, 16, w["A3"], 0, ld("A3"), B)
        self._wgrad(a[1], w["dA3"], 0, ld("dA3"), w["A2"], 0, ld("A2"), B)
        self._wgrad(a[0], w["dA2"], 0, ld("dA2"), w["Y1"], 0, ld("Y1"), B)
        self._wgrad(c[2], w["dvalue"], 0, 8, w["C3"], 0, ld("C3"), B)
        self._wgrad(c[1], w["dC3"], 0, ld("dC3"), w["C2"], 0, ld("C2"), B)
        self._wgrad(c[0], w["dC2"], 0, ld("dC2"), w["Y1"], H, ld("Y1"), B)
        self._wgrad(ac.L_cat, w["dY1"], 0, ld("dY1"), w["Xac"], 0, ld("Xac"), B)
        self._wgrad(e[2], w["dLat"], 0, ld("dLat"), w["H2"], 0, ld("H2"), B)
        self._wgrad(e[1], w["dH2"], 0, ld("dH2"), w["H1"], 0, ld("H1"), B)
        self._wgrad(e[0], w["dH1"], 0, ld("dH1"), w["Xp"], 0, ld("Xp"), B)
        self._wgrad_flush()
        # ---- data-parallel reduction: ONE call over [policy | adaptation | KL sum].  In the lagged
        # schedule the adaptation range holds minibatch i-1's gradient (joined here); serially it is still zero ----
        # several GPUs without the side communicator: the adaptation gradient rides in the ONE call
        early = lag and pending and allreduce is not None and self._ada_g is None
        if early:
            torch.cuda.current_stream().wait_event(self._ev_grads)
        self._reduce(allreduce, 0, norm_n=ac.n_main)
        if debug:      # parity tests read the raw gradient / statistics
            self.debug_grad = g_used[:ac.n_total].clone()
        # ---- clip + KL-adaptive lr + Adam, all on the device ----
        self._policy_step(B, world, allreduce)
        if lag:
            if early:
                self._adapt_step(g_used)
            # regression target of THIS minibatch (ppo.py:158: the latent of the already updated encoder); the next
            # call's teacher forward overwrites the latent slot, so it is copied out - after the side branch's loss
            # kernel has read the previous target
            ac.forward_encoder(B)
            if pending:
                torch.cuda.current_stream().wait_event(self._ev_loss)
            self._target(B)[:B, :ac.latent_dim].copy_(w["Xac"][:B, ac.num_obs:ac.num_obs + ac.latent_dim])
            torch.cuda.current_stream().wait_event(self._ev_join)          # every forked branch rejoins
            return
        # ---- adaptation module (ppo.py:156-170): target latent from the UPDATED encoder ----
        # (high_level_policy with USE_LATENT = False has no adaptation module to train: high_level_policy/ppo/ppo.py:157)
        for _ in range(A.num_adaptation_module_substeps if ac.use_latent else 0):
            ac.forward_encoder(B)
            self._adapt_grads(B, world, lagged=False)
            self._reduce(allreduce, ac.n_main)
            if debug:
                self.debug_grad[ac.n_main:] = g_used[ac.n_main:ac.n_total] if self._ada_g is None else self._ada_g
            self._adapt_step(g_used)
        if debug:
            self.debug_stats = self._loss_acc - acc0

    def _chunks(self, B):
        """[(tile_begin, tile_end)] x 2 when the batch's last wave of 128-row tiles would be poorly filled, else None
        (RL_PPO_CHUNKS=0: never)."""
        if os.environ.get("RL_PPO_CHUNKS", "1") == "0":
            return None
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        tiles = (B + 127) // 128
        full = tiles // sms * sms
        # (measured: 24000 rows 7.69 -> 7.28 ms per update; at 1536 tiles - 10.4 waves - the ragged wave is 1 / 11 of the
        # work and the extra launches cost as much as they save)
        if full == 0 or tiles == full or (tiles - full) > 0.7 * sms or tiles > 3 * sms:
            return None
        split = int(os.environ.get("RL_PPO_CHUNK_SPLIT", "0"))      # A/B knob: tiles in the first chunk (0: the full waves)
        if 0 < split < tiles:
            full = split
        return [(0, full), (full, tiles)]

    def _reduce(self, allreduce, start, norm_n=0):
        """Sum over ranks of the flat gradient buffer from element `start` (0: everything incl. the KL word behind
        it; n_main: the adaptation module's range).  No-op on one GPU."""
        ac = self.actor_critic
        if allreduce is None:
            return
        if allreduce == "peer":
            # ONE kernel over NVLink peer memory (csrc/peer_allreduce.cu): the sum, the squared norm of the policy part
            # and the zeroing of the accumulation buffer
            self._peer.all_reduce(self._g_red, norm_n=norm_n, start=start)
        elif self._ada_g is not None:
            # (the adaptation gradient lives in its own buffer: the flat buffer's adaptation range stays zero)
            if start == 0:
                allreduce(ac._grad_store[:ac.n_total + 4])
            else:
                allreduce(self._ada_g)
        else:
            allreduce(ac._grad_store[start:ac.n_total + 4])

    def _policy_step(self, B, world, allreduce):
        """ppo.py:116-124, 146-150: gradient norm -> clip coefficient, KL mean -> learning rate, Adam on the policy
        parameters; the finalize kernel also folds this minibatch's loss sums into the per-update accumulators."""
        ac, A = self.actor_critic, PPO_Args
        P, stream = _lib.ptr, _lib.current_stream()
        adaptive = int(A.desired_kl is not None and A.schedule == "adaptive")
        if allreduce == "peer":
            g_used = self._g_red
            _lib.check(self._lib.rl_grad_finalize_from_norm(
                P(self._peer.norm2), P(self._stats), P(self._ctrl), float(B * world), float(A.desired_kl or 0.0),
                float(A.max_grad_norm), adaptive, P(self._loss_acc), P(g_used[ac.n_total:ac.n_total + 1]), stream))
        else:
            g_used = ac.flat_grad
            kl_slot = ac._grad_store[ac.n_total:ac.n_total + 1] if allreduce is not None else None
            _lib.check(self._lib.rl_grad_finalize(
                P(ac.flat_grad[:ac.n_main]), ac.n_main, P(self._stats), P(self._ctrl), P(self._fin_ws), float(B * world),
                float(A.desired_kl or 0.0), float(A.max_grad_norm), adaptive, P(self._loss_acc), P(kl_slot), stream))
        if self._fused_adam and ac.use_chain:
            # Adam + the bf16 operands of every policy layer in one launch
            ac.adam_shadows(0, ac.n_main, g_used, P(self._ctrl), 0.0, 1, self._steps.data_ptr(), self._main_layers)
        else:
            _lib.check(self._lib.rl_adam(P(ac.flat), P(g_used), P(ac.flat_m), P(ac.flat_v), ac.n_main, P(self._ctrl), 0.0, 1,
                                         0.9, 0.999, 1e-8, 0, 1.0, self._steps.data_ptr(), stream))
            ac.refresh_shadows(self._main_layers)

    def _target(self, B):
        """bf16 [B, 24] copy of the adaptation regression target (lagged schedule only)."""
        t = getattr(self, "_tgt", None)
        if t is None or t.shape[0] < B:
            self._tgt = t = torch.zeros(B, 24, dtype=torch.bfloat16, device=self.device)
        return t

    def _adapt_grads(self, B, world, lagged, loss_event=None, forward=True):
        """ppo.py:157-164: adaptation forward, regression loss against the encoder latent, backward; the gradient lands
        in the adaptation range of the flat buffer, the squared error in the per-update accumulator.  lagged: the rows
        are those of the previous call (history already gathered, target in `_tgt`)."""
        ac = self.actor_critic
        w = ac._ws
        P = _lib.ptr
        stream = _lib.current_stream()
        ld = lambda k: w[k].shape[1]
        d = ac.L_ada
        inv_gb = 1.0 / (B * world)
        if forward:             # (lagged schedule: the previous call's side branch may have run it already)
            ac.forward_adaptation(B, save=True)
        if lagged:
            tgt = self._target(B)
            _lib.check(self._lib.rl_adapt_loss(P(w["pred"]), P(tgt), tgt.shape[1], 0, B, inv_gb, P(w["dpred"]), P(self._loss_acc), stream))
            if loss_event is not None:
                loss_event.record()
        else:
            _lib.check(self._lib.rl_adapt_loss(P(w["pred"]), P(w["Xac"]), ld("Xac"), ac.num_obs, B, inv_gb, P(w["dpred"]),
                                               P(self._loss_acc), stream))
        if ac.use_chain:
            ac._chain(("adaptation_backward",), chain.adaptation_backward_program).run(B)
        else:
            self._dgrad(d[2], w["dpred"], 0, 24, w["dD2"], 0, ld("dD2"), B, aux=w["D2"], ld_aux=ld("D2"))
            self._dgrad(d[1], w["dD2"], 0, ld("dD2"), w["dD1"], 0, ld("dD1"), B, aux=w["D1"], ld_aux=ld("D1"))
        self._wgrad(d[2], w["dpred"], 0, 24, w["D2"], 0, ld("D2"), B)
        self._wgrad(d[1], w["dD2"], 0, ld("dD2"), w["D1"], 0, ld("D1"), B)
        self._wgrad(d[0], w["dD1"], 0, ld("dD1"), w["Xh"], 0, ld("Xh"), B)
        self._wgrad_flush()

    def _adapt_step(self, g_used):
        """ppo.py:166-168: Adam on the adaptation module (its own fixed learning rate, no clipping)."""
        ac, A = self.actor_critic, PPO_Args
        n_ad = ac.n_total - ac.n_main
        off = ac.n_main * 4
        g_ptr = (g_used.data_ptr() + off) if self._ada_g is None else self._ada_g.data_ptr()
        if self._fused_adam and ac.use_chain:
            ac.adam_shadows(ac.n_main, n_ad, g_used, None, float(A.adaptation_module_learning_rate), 0,
                            self._steps.data_ptr() + 8, ac.L_ada, grad_ptr=g_ptr)
            return
        _lib.check(self._lib.rl_adam(ac.flat.data_ptr() + off, g_ptr, ac.flat_m.data_ptr() + off,
                                     ac.flat_v.data_ptr() + off, n_ad, None, float(A.adaptation_module_learning_rate), 0,
                                     0.9, 0.999, 1e-8, 0, 1.0, self._steps.data_ptr() + 8, _lib.current_stream()))
        ac.refresh_shadows(ac.L_ada)

    def sync_replicas(self):
        """Data parallelism keeps one replica of the learner per rank and only exchanges gradients, so the replicas
        must START identical: broadcast the parameters, Adam moments, learning-rate control and step counters from rank
        0 (collective; update() calls it once).  Without it, ranks that were not seeded identically diverge silently."""
        from ..sharding import broadcast_
        ac = self.actor_critic
        broadcast_(ac.flat, ac.flat_m, ac.flat_v, self._ctrl, self._steps)
        ac.refresh_shadows()
        self._replicas_synced = True

    def update(self):
        """ppo.py:94-178."""
        from ..sharding import all_reduce_sum_, world_size
        A, st, ac = PPO_Args, self.storage, self.actor_critic
        world = world_size()
        allreduce = all_reduce_sum_ if world > 1 else None
        if world > 1 and not getattr(self, "_replicas_synced", False):
            self.sync_replicas()
        # RL_PEER_ALLREDUCE=1: the fused NVLink peer kernel instead of NCCL (see DESIGN.md section 8 for the A/B)
        if world > 1 and os.environ.get("RL_PEER_ALLREDUCE", "0") == "1":
            if self._peer is None:
                self.enable_peer_allreduce()
            allreduce = "peer"
        elif world > 1 and os.environ.get("RL_PPO_SIDE_COMM", "0") == "1" and self._peer is None:
            # opt-in: measured SLOWER than the one collective per minibatch (2 GPUs, 4000 envs each: 8.71 vs 8.51 ms per
            # update - a second latency-bound NCCL launch per step costs more than the join it removes)
            self.enable_side_comm()
        batch = st.num_envs * st.num_transitions_per_env
        mb = batch // A.num_mini_batches
        indices = torch.randperm(A.num_mini_batches * mb, device=self.device)   # ONE permutation for all epochs (:103)
        self._loss_acc.zero_()
        # The ~25 launches of a minibatch step are captured ONCE in a CUDA graph that reads its row indices from a fixed
        # buffer; each of the 20 steps is then one index copy + one graph replay (with several GPUs the NCCL all-reduce
        # is captured too - every rank captures the same sequence; RL_PPO_GRAPH_MULTI=0 launches eagerly instead).
        multi_ok = world == 1 or os.environ.get("RL_PPO_GRAPH_MULTI", "1") != "0"
        use_graph = self.use_cuda_graph and multi_ok and not getattr(self, "debug_keep_grad", False)
        # Lagged adaptation schedule (minibatch_step docstring): two graphs - the first minibatch of an update has nothing
        # pending - and one eager flush after the last.  RL_PPO_OVERLAP=0 keeps the serial order.
        lag = (use_graph and self.overlap_adaptation and ac.use_chain and A.num_adaptation_module_substeps == 1 and
               ac.use_latent)
        ws_gen = getattr(ac, "_ws_gen", 0)
        # RL_PPO_ONE_GRAPH=1: ALL minibatch steps of the update in ONE graph (each step reads its rows from its own slice of a
        # [steps, mb] index buffer) - one replay per update instead of 20, no graph-launch gap between steps.  Same results
        # (test_lagged_schedule_matches_serial_update), measured no faster (4000 envs 7.37 / 7.38 vs 7.26 / 7.35 ms, 32768 envs
        # 37.9 vs 38.1, profiles/jobs/r2_job74.sh): the side branch already filled the gaps.  Opt-in.
        n_steps = A.num_learning_epochs * A.num_mini_batches
        one = use_graph and os.environ.get("RL_PPO_ONE_GRAPH", "0") == "1"
        shape_key = (one, n_steps, A.num_mini_batches)
        if use_graph and (self._graph is None or self._graph_B != mb or self._graph_lag != lag or self._graph_ws_gen != ws_gen or
                          getattr(self, "_graph_shape", None) != shape_key or
                          self._graph_reduce != (allreduce if isinstance(allreduce, str) else allreduce is not None)):
            ac.workspace(mb, backward=True)        # allocate outside the capture
            ac.prepare_update_chains()
            self._target(mb)
            self._idx_buf = torch.zeros(mb, dtype=torch.long, device=self.device)
            torch.cuda.synchronize()
            state = (ac.flat, ac.flat_m, ac.flat_v, ac._grad_store, self._ctrl, self._steps, self._loss_acc, self._stats) + \
                ((self._ada_g,) if self._ada_g is not None else ())
            snap = [t.clone() for t in state]
            # capture on a high-priority stream: kernel nodes keep their stream's priority, so the policy path's CTAs
            # are placed before those of the side branch (adaptation module), which only fills what is left
            kw = {}
            if lag and os.environ.get("RL_PPO_PRIO", "1") != "0":
                if self._hp is None:
                    self._hp = torch.cuda.Stream(device=self.device, priority=-1)
                kw["stream"] = self._hp
            graphs = []
            if one:
                self._idx_all = torch.zeros(n_steps, mb, dtype=torch.long, device=self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, **kw):
                    for k in range(n_steps):
                        self.minibatch_step(self._idx_all[k], world, allreduce, lag=lag, pending=lag and k > 0)
                graphs.append(g)
            else:
                for pending in ((False, True) if lag else (False,)):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, **kw):
                        self.minibatch_step(self._idx_buf, world, allreduce, lag=lag, pending=pending)
                    graphs.append(g)
            # capture does not execute, but keep the state bit-identical in any case
            for t, s0 in zip(state, snap):
                t.copy_(s0)
            self._graph, self._graph_rest, self._graph_B, self._graph_lag = graphs[0], graphs[-1], mb, lag
            self._graph_ws_gen = getattr(ac, "_ws_gen", 0)
            self._graph_shape = shape_key
            self._graph_reduce = allreduce if isinstance(allreduce, str) else allreduce is not None
        first = True
        if one:
            # the same permutation serves every epoch (:103): step e * num_mini_batches + i takes slice i
            self._idx_all.copy_(indices.view(A.num_mini_batches, mb).repeat(A.num_learning_epochs, 1))
            self._graph.replay()
        for _ in range(0 if one else A.num_learning_epochs):
            for i in range(A.num_mini_batches):
                if use_graph:
                    self._idx_buf.copy_(indices[i * mb:(i + 1) * mb])
                    (self._graph if first else self._graph_rest).replay()
                    first = False
                else:
                    self.minibatch_step(indices[i * mb:(i + 1) * mb], world, allreduce)
        if use_graph and lag:
            # the last minibatch's adaptation update
            self._adapt_grads(mb, world, lagged=True, forward=not getattr(self, "_ada_pre", False))
            self._reduce(allreduce, ac.n_main)
            self._adapt_step(self._g_red if allreduce == "peer" else ac.flat_grad)
        if world > 1:
            all_reduce_sum_(self._loss_acc)        # loss means span the env shards of every rank
        n_upd = A.num_learning_epochs * A.num_mini_batches
        acc = (self._loss_acc / (mb * world)).tolist()          # the only device->host read of the update
        self.learning_rate = float(self._ctrl[0])
        st.clear()
        lat = ac.latent_dim
        return acc[1] / n_upd, acc[0] / n_upd, acc[3] / lat / n_upd / A.num_adaptation_module_substeps
