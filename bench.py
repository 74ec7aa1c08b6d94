#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: fused env-steps/s; PPO samples/s rides along).

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA kernels)
    python bench.py --impl reference --gpus 1 --steps K ...   # reference arm: the reference's torch
                                                              # CPU algorithm (oracle port) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...         # one rank per GPU, env-sharded (weak scaling)

A "step" is one `LeggedRobot.step(actions)` of the fused pipeline (PD torques + terminations + all
reward terms + observations with noise/clipping + state updates) for every env of the workload, on
synthetic simulator state (the closed-source physics is outside the path).  Workload: Mini Cheetah
flat, 32768 envs per GPU (BASELINE.json configs[4] = configs[1] at the scale-out size; the 4000-env
configs[1] point and a 262144-env HBM-resident point are reported under "also").

Timing hygiene: W >= 3 warm-up steps; the K timed steps are captured in ONE CUDA graph and replayed
once between CUDA events on the launching stream; consecutive steps rotate over R independent env
replicas whose combined footprint exceeds 2x the 126 MB L2, so no step finds its inputs cache-hot;
multi-GPU time is the max over ranks; SM clocks / throttle reasons are sampled with NVML during the
timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# SURVEY.md 8(d) algorithmic bytes per env-step: Mini Cheetah flat (780 read + 645 written), Go1 (17 bodies: +48 B of
# contact rows), Mini Cheetah rough (229-column observations + measured heights; the 9.4 MB height table is L2 resident)
BYTES_PER_ENV_STEP = {"mc_flat": 1425, "go1": 1473, "mc_rough_full": 2921}
L2_BYTES = 126 * 1024 * 1024
H2D_PER_ENV = 48 + 52 + 96 + 156      # actions, root, dof_state, contact rows (Mini Cheetah)
D2H_PER_ENV = 168 + 72 + 4 + 1        # obs, privileged obs, reward, reset flag


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """NVML SM-clock / throttle-reason sampler running while the timed region executes."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.0005)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def bind_to_gpu_numa(local_rank):
    """Pins this process to the CPU cores NVML reports as local to its GPU, BEFORE any pinned host buffer is allocated, so
    that the staging buffers of the e2e section are first touched (and therefore placed) on the GPU's own NUMA node: with
    eight ranks on the default policy every rank's pinned pages can end up on one node and all host<->device traffic
    shares that node's memory controllers and the inter-socket link.  RL_BENCH_NUMA=0 switches it off.  Returns a
    description for the JSON line."""
    if os.environ.get("RL_BENCH_NUMA", "1") == "0":
        return "off"
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index(local_rank))
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        if not use or len(use) == len(allowed):
            return "no-op (%d GPU-local cores of %d allowed)" % (len(use), len(allowed))
        global _ORIG_AFFINITY
        _ORIG_AFFINITY = allowed
        os.sched_setaffinity(0, use)
        return "bound to %d GPU-local cores (%d..%d) of %d" % (len(use), use[0], use[-1], len(allowed))
    except Exception as e:      # no NVML / no permission: the default placement stands
        return "unavailable (%s)" % type(e).__name__


_ORIG_AFFINITY = None


def restore_affinity():
    """The reference / CPU-baseline legs (own processes, inheriting this one's mask) get every host core back."""
    if _ORIG_AFFINITY is not None:
        os.sched_setaffinity(0, _ORIG_AFFINITY)


def build_replicas(case, envs, n_rep, device, seed0=0):
    """n_rep independent env instances with synthetic simulator state."""
    import numpy as np
    import torch
    from cases import build_case
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    from rapid_locomotion_rl_b200.sim import synthetic_state
    reps = []
    for r in range(n_rep):
        cfg, robot, terrain = build_case(case, envs)
        env = LeggedRobot(cfg, sim_device=device, headless=True, terrain=terrain, seed=seed0 + r)
        p = env.params
        st = synthetic_state(seed0 * 1000 + r, envs, robot.num_bodies, 12, np.float32(p.default_dof_pos), p.feet_idx,
                             p.term_idx[:p.n_term_bodies])
        env.sim.root_states.copy_(torch.from_numpy(st["root_states"]))
        env.sim.dof_state.copy_(torch.from_numpy(st["dof_state"]).view(-1, 2))
        env.sim.contact_forces.copy_(torch.from_numpy(st["contact_forces"]).view(-1, 3))
        env._resample_commands(torch.arange(envs, device=device))       # commands from the GAC initial sample
        env.episode_length_buf.copy_(torch.randint(0, 1001, (envs,), device=device))   # init_at_random_ep_len
        actions = torch.randn(envs, 12, device=device)
        reps.append((env, actions, st))
    return reps


def time_env_steps(reps, steps, warmup):
    """W eager warm-up steps, then exactly `steps` steps captured in one CUDA graph, replayed once
    warm and once timed.  Returns elapsed milliseconds of the timed replay."""
    import torch
    for i in range(max(3, warmup)):
        env, a, _ = reps[i % len(reps)]
        env.step(a)
    for env, _, _ in reps:
        env.use_device_step_counter(True)
    # Behind a real simulator the state changes between two steps of an env.  A replica therefore alternates between TWO
    # synthetic simulator states (and action batches) from visit to visit: with a frozen state last_dof_vel == dof_vel after
    # the first step, the dof_acc quotient (last_dof_vel - dof_vel) / dt is exactly 0 in every lane and each fp32 division
    # of the kernel takes its slow path (ncu source page, profiles/r02_env_step.md) - a benchmark artefact, not the
    # workload.  The second state costs no launch: the kernel arguments of a visit simply point at it.
    alt = []
    for env, a, _ in reps:
        sim = env.sim
        g2 = torch.Generator(device=a.device); g2.manual_seed(12345 + len(alt))
        root2, dof2, con2 = sim.root_states.clone(), sim.dof_state.clone(), sim.contact_forces.clone()
        dof2.view(-1, 2)[:, 1].add_(torch.randn(dof2.view(-1, 2).shape[0], device=a.device, generator=g2) * 0.5)
        root2[:, 7:13].add_(torch.randn(root2.shape[0], 6, device=a.device, generator=g2) * 0.05)
        alt.append(((sim.root_states, sim.dof_state, sim.contact_forces, a),
                    (root2, dof2, con2, a + 0.1 * torch.randn(a.shape, device=a.device, generator=g2))))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    g._keepalive = alt
    with torch.cuda.graph(g):
        for i in range(steps):
            r = i % len(reps)
            env = reps[r][0]
            root, dof, con, a = alt[r][(i // len(reps)) % 2]
            env._bufs.root_states, env._bufs.dof_state, env._bufs.contact_forces = root.data_ptr(), dof.data_ptr(), con.data_ptr()
            env.step(a)
    for r, (env, _, _) in enumerate(reps):       # leave the envs bound to their own simulator tensors
        root, dof, con, _a = alt[r][0]
        env._bufs.root_states, env._bufs.dof_state, env._bufs.contact_forces = root.data_ptr(), dof.data_ptr(), con.data_ptr()
    g.replay()      # untimed replay: graph upload, clocks up
    torch.cuda.synchronize()
    return g


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    numa = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    pk, pk_kind = peaks()
    bpe = BYTES_PER_ENV_STEP["mc_flat"]
    if args.only_ppo:      # development aid: just the learner metric
        if rank == 0:
            print(json.dumps(ppo_bench(args.ppo_envs, 24, device, world, pk)))
        else:
            ppo_bench(args.ppo_envs, 24, device, world, pk)
        return

    def measure(envs, steps, warmup, sample_clocks, case="mc_flat"):
        n_rep = max(2, math.ceil(2.0 * L2_BYTES / (envs * BYTES_PER_ENV_STEP[case])))
        reps = build_replicas(case, envs, n_rep, device, seed0=rank)
        g = time_env_steps(reps, steps, warmup)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(physical_gpu_index(local)) if sample_clocks else None
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.__enter__()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        if sampler:
            sampler.__exit__()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, n_rep, reps, sampler

    ms, n_rep, reps, sampler = measure(args.envs, args.steps, args.warmup, True)
    value = args.envs * world * args.steps / (ms * 1e-3)
    per_launch_s = ms * 1e-3 / args.steps
    achieved = args.envs * bpe / per_launch_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "env_step_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(str(args.envs))

    # ---- end to end through the public API with HOST buffers --------------------------------------
    # Every step: pinned host -> H2D of that step's simulator rows and actions -> LeggedRobot.step ->
    # D2H of obs / priv / rew / reset -> the host waits for them.  Two env groups are double-buffered on
    # two streams (the usual vectorised-env split), so the D2H of one group's step overlaps the H2D of the
    # other's (PCIe is full duplex); "serial" is the same loop with one group and a sync after every step.
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    class HostSide:
        def __init__(self, rep):
            self.env, actions, st = rep
            self.env.use_device_step_counter(False)
            self.h_in = [pin(st["root_states"]), pin(st["dof_state"].reshape(-1, 2)), pin(st["contact_forces"].reshape(-1, 3)),
                         actions.cpu().pin_memory()]
            self.d_act = torch.empty_like(actions)
            e = self.env
            self.d_in = [e.sim.root_states, e.sim.dof_state, e.sim.contact_forces, self.d_act]
            self.h_out = [torch.empty(e.obs_buf.shape, pin_memory=True), torch.empty(e.privileged_obs_buf.shape, pin_memory=True),
                          torch.empty(e.rew_buf.shape, pin_memory=True), torch.empty(e.reset_buf.shape, dtype=torch.bool, pin_memory=True)]
            self.stream = torch.cuda.Stream()
            self.done = torch.cuda.Event()
            self.h2d_done = torch.cuda.Event()
            self.pending = False
            self.other = None

        def submit(self):
            with torch.cuda.stream(self.stream):
                # stagger the groups: this group's H2D starts when the other group's H2D has finished, i.e.
                # while the other group computes and copies its results back (H2D and D2H engines both busy)
                if self.other is not None and self.other.pending:
                    self.stream.wait_event(self.other.h2d_done)
                for d, h in zip(self.d_in, self.h_in):
                    d.copy_(h, non_blocking=True)
                self.h2d_done.record(self.stream)
                outs = self.env.step(self.d_act)[:4]
                for h, d in zip(self.h_out, outs):
                    h.copy_(d, non_blocking=True)
                self.done.record(self.stream)
            self.pending = True

        def wait(self):
            if self.pending:
                self.done.synchronize()      # this step's obs / rew / reset are now readable on the host
                self.pending = False

    def e2e_run(groups, k):
        for i in range(k):
            g = groups[i % len(groups)]
            g.wait()                         # the group's previous step has been read back: buffers are free
            g.submit()
        for g in groups:
            g.wait()

    torch.cuda.synchronize()
    groups = [HostSide(reps[0]), HostSide(reps[1])]
    k_e2e = min(args.steps, 200)

    def e2e_time(gs):
        e2e_run(gs, 4)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_run(gs, k_e2e)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt
    e2e_serial_s = e2e_time(groups[:1])
    groups[0].other, groups[1].other = groups[1], groups[0]
    e2e_s = e2e_time(groups)
    e2e_value = args.envs * world * k_e2e / e2e_s
    e2e_serial = args.envs * world * k_e2e / e2e_serial_s
    assert float(groups[0].h_out[0].abs().sum()) > 0 and float(groups[1].h_out[0].abs().sum()) > 0
    # ---- packed variant (LeggedRobot.pack_io): the simulator's rows + actions are views of ONE pinned block, the outputs of
    # another, so a step is one H2D copy, one launch, one D2H copy; three env groups in flight (PCIe moves one group's
    # inputs and another's outputs while the third is being submitted) ----
    e2e_packed = None
    if len(reps) >= 5:
        class PackedSide(HostSide):
            def __init__(self, rep):
                self.env, actions, st = rep
                e = self.env
                e.use_device_step_counter(False)
                self.d_in_block, self.d_out_block, lay_in, lay_out = e.pack_io()
                self.h_in_block = torch.empty(self.d_in_block.numel(), dtype=torch.uint8).pin_memory()
                self.h_out_block = torch.empty(self.d_out_block.numel(), dtype=torch.uint8).pin_memory()
                hv = e.host_views(self.h_in_block, lay_in)          # what a host-side simulator fills in place
                hv["root_states"].copy_(torch.from_numpy(st["root_states"]))
                hv["dof_state"].copy_(torch.from_numpy(st["dof_state"]).view(-1, 2))
                hv["contact_forces"].copy_(torch.from_numpy(st["contact_forces"]).view(-1, 3))
                hv["actions"].copy_(actions.cpu())
                self.h_out = [e.host_views(self.h_out_block, lay_out)["obs"]]
                self.stream = torch.cuda.Stream()
                self.done, self.h2d_done = torch.cuda.Event(), torch.cuda.Event()
                self.pending, self.other = False, None

            def submit(self):
                with torch.cuda.stream(self.stream):
                    if self.other is not None and self.other.pending:
                        self.stream.wait_event(self.other.h2d_done)
                    self.d_in_block.copy_(self.h_in_block, non_blocking=True)
                    self.h2d_done.record(self.stream)
                    self.env.step(self.env.packed_actions)
                    self.h_out_block.copy_(self.d_out_block, non_blocking=True)
                    self.done.record(self.stream)
                self.pending = True
        # (three groups: a fourth one measured the same 1.44e8 on a fast box, and touching a fourth replica here moves where
        # the allocator later places the replicas of the small-grid sizes - 4000 envs: 8.3e8 instead of 9.0e8 env-steps/s,
        # reproducibly; A/B of the two bench versions inside one gpurun call)
        pgroups = [PackedSide(reps[i]) for i in (2, 3, 4)]
        for i, gp in enumerate(pgroups):
            gp.other = pgroups[i - 1]
        e2e_packed = args.envs * world * k_e2e / e2e_time(pgroups)
        assert all(float(gp.h_out[0].abs().sum()) > 0 for gp in pgroups)
    # ---- zero-copy variant: the kernel reads the pinned host rows and writes the pinned host outputs itself
    # (LeggedRobot.bind_host_io / step_host: one launch per step, no staging copies); one group with a sync per
    # step, and two groups on two streams so that one group's PCIe reads overlap the other's writes ----
    e2e_zero_copy = e2e_zero_copy2 = None
    try:
        for gs in groups:
            e_ = gs.env
            gs.h_reset_u8 = torch.empty(e_.num_envs, dtype=torch.uint8, pin_memory=True)
            e_.bind_host_io(gs.h_in[0], gs.h_in[1], gs.h_in[2], gs.h_out[0], gs.h_out[1], gs.h_out[2], gs.h_reset_u8)

        def zc_submit(gs):
            with torch.cuda.stream(gs.stream):
                gs.env.step_host(gs.h_in[3])
                gs.done.record(gs.stream)
            gs.pending = True

        def zc_run(gl, k):
            for i in range(k):
                gs = gl[i % len(gl)]
                gs.wait()                    # this group's previous outputs are in host memory
                zc_submit(gs)
            for gs in gl:
                gs.wait()

        def zc_time(gl):
            zc_run(gl, 4)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            zc_run(gl, k_e2e)
            torch.cuda.synchronize()
            dtz = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dtz], device=device, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtz = float(t.item())
            return args.envs * world * k_e2e / dtz
        e2e_zero_copy = zc_time(groups[:1])
        e2e_zero_copy2 = zc_time(groups)
        assert float(groups[0].h_out[0].abs().sum()) > 0 and float(groups[1].h_out[0].abs().sum()) > 0
    except Exception as exc:
        e2e_zero_copy = "failed: %s" % str(exc)[:160]
    env = None
    del groups
    restore_affinity()
    # the headline e2e figure: the best of the ways through the public API (all of them: pinned HOST buffers in, results in
    # host memory, the host waiting for every step) - which one wins depends on how the box's PCIe path treats the copy
    # engines (boxes of this pool: packed copies 1.07e8 - 1.44e8, kernel-driven zero copy 1.0e8 everywhere)
    cands = {"separate_copies_two_groups": e2e_value, "packed_copies": e2e_packed or 0.0}
    for name, v in (("zero_copy_one_group", e2e_zero_copy), ("zero_copy_two_groups", e2e_zero_copy2)):
        if isinstance(v, float):
            cands[name] = v
    e2e_best = max(cands.items(), key=lambda kv: kv[1])

    # ---- other sizes (rank 0 only, N=1): the 4000-env configs[1] point and an HBM-resident one -----
    also = {}
    if world == 1 and not args.quick:
        del reps
        torch.cuda.empty_cache()
        for envs in (4000, 262144):
            k2 = min(args.steps, 1000 if envs == 4000 else 200)
            ms2, nr2, reps2, _ = measure(envs, k2, args.warmup, False)
            also[str(envs)] = {"value": envs * k2 / (ms2 * 1e-3), "ms_per_step": ms2 / k2, "steps": k2,
                               "replicas": nr2, "roofline_frac": envs * bpe / (ms2 * 1e-3 / k2) / 1e9 / pk["hbm_gbs"]}
            del reps2
            torch.cuda.empty_cache()

    if world == 1 and not args.quick:
        # BASELINE configs[3] (Go1: 17 bodies, 1473 B per env-step) and configs[2] (Mini Cheetah on the full 1800 x 2600
        # heightfield with measured heights: 2921 B per env-step), at the configs' 4000 envs and at the scale-out size
        for case in ("go1", "mc_rough_full"):
            b_case = BYTES_PER_ENV_STEP[case]
            for envs in (4000, 32768):
                k2 = min(args.steps, 300)
                try:
                    ms2, nr2, reps2, _ = measure(envs, k2, args.warmup, False, case=case)
                    also["%s_%d" % (case, envs)] = {
                        "value": envs * k2 / (ms2 * 1e-3), "ms_per_step": ms2 / k2, "steps": k2, "replicas": nr2,
                        "bytes_per_env_step": b_case, "roofline_frac": envs * b_case / (ms2 * 1e-3 / k2) / 1e9 / pk["hbm_gbs"]}
                    del reps2
                except Exception as exc:
                    also["%s_%d" % (case, envs)] = {"error": str(exc)[:200]}
                torch.cuda.empty_cache()

    if world == 1 and not args.quick:
        # the gymapi-shaped sequence - `decimation` x (torque kernel between physics sub-steps) +
        # one post-physics kernel per step (5 launches instead of the 1 fused launch)
        k3 = min(args.steps, 300)
        n_rep = max(2, math.ceil(2.0 * L2_BYTES / (args.envs * bpe)))
        reps3 = build_replicas("mc_flat", args.envs, n_rep, device, seed0=rank)
        for env3, a3, _ in reps3:
            env3.use_device_step_counter(True)

        def seq(i):
            env3, a3, _ = reps3[i % len(reps3)]
            for _ in range(env3.cfg.control.decimation):
                env3._compute_torques(a3)
            env3.post_physics_step(a3)
        for i in range(4):
            seq(i)
        torch.cuda.synchronize()
        g3 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g3):
            for i in range(k3):
                seq(i)
        g3.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g3.replay(); e1.record()
        torch.cuda.synchronize()
        ms3 = e0.elapsed_time(e1)
        also["decimation_sequence"] = {"value": args.envs * k3 / (ms3 * 1e-3), "ms_per_step": ms3 / k3, "steps": k3,
                                       "launches_per_step": 1 + reps3[0][0].cfg.control.decimation, "envs": args.envs,
                                       "note": "4 x rl_env_torques + rl_env_post_physics per step (the shape of the real gymapi loop)"}
        del reps3, g3
        torch.cuda.empty_cache()

    ppo = None
    if not args.quick:
        try:
            del reps
        except NameError:
            pass
        torch.cuda.empty_cache()
        ppo = ppo_bench(args.ppo_envs, 24, device, world, pk)
        # BASELINE configs[4]: the learner at the scale-out size (32768 envs per GPU x 24: 196608-row minibatches)
        try:
            torch.cuda.empty_cache()
            ppo["c5_32768_envs"] = ppo_bench(32768, 24, device, world, pk, iters=2)
        except Exception as exc:
            ppo["c5_32768_envs"] = {"error": str(exc)[:200]}
        torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            ppo["reference"] = ppo_reference_block(args.ppo_envs)

    runner = None
    if world == 1 and not args.quick:
        runner = runner_bench(args.ppo_envs, device)

    out = {
        "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "mini_cheetah_flat fused env step (PD torques + terminations + 12 reward terms + obs/noise/clip "
                               "+ state), %d envs per GPU" % args.envs,
                   "envs_per_gpu": args.envs, "parallelism": "env-sharded x%d, no data-path collective" % world,
                   "cache": "steps rotate over %d env replicas (%.0f MB > 2x L2) so inputs are never L2-hot" %
                            (n_rep, n_rep * args.envs * bpe / 1e6),
                   "launch": "K steps captured in one CUDA graph, device-side RNG step counter"},
        "clocks": sampler.summary(),
        "e2e": {"host_numa": numa, "value": e2e_best[1], "variant": e2e_best[0], "unit": "env-steps/s", "h2d_bytes_per_step": args.envs * H2D_PER_ENV,
                "d2h_bytes_per_step": args.envs * D2H_PER_ENV, "steps": k_e2e, "serial_value": e2e_serial, "zero_copy_value": e2e_zero_copy, "zero_copy_two_groups_value": e2e_zero_copy2,
                "separate_copies_two_groups_value": e2e_value, "packed_three_groups_value": e2e_packed,
                "note": "value = the best of the ways through the public API (`variant` names it), all with HOST buffers and the host waiting for "
                        "every step's results: (packed_three_groups_value) LeggedRobot.pack_io - the simulator rows + actions are "
                        "views of ONE pinned block, the outputs of another: one H2D copy, LeggedRobot.step, one D2H copy per step, "
                        "three env groups in flight; (separate_copies_two_groups_value) four H2D + four D2H copies per step.  "
                        "every step: pinned host buffers -> H2D -> LeggedRobot.step -> D2H of obs/priv/rew/reset -> host "
                        "wait; two env groups double-buffered on two streams, H2D of one staggered against the D2H of the other "
                        "(serial_value: one group, sync per step; zero_copy_value: LeggedRobot.bind_host_io / step_host - the kernel "
                        "reads the pinned rows and writes the pinned outputs in place, one launch per step, one group; "
                        "zero_copy_two_groups_value: the same on two streams)"},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_kind,
                     "kernel": "rows::env_step_rows_kernel<FUSE=true,MINB=7> (csrc/env_step_rows.cu)", "algorithmic_bytes_per_launch": args.envs * bpe},
        "also": also,
        "ppo": ppo,
        "runner": runner,
    }
    if rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_baseline(args.envs)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ppo_bench(n_envs, T, device, world, pk, iters=5):
    """BASELINE metric 2: PPO samples/s = N*T / (t_compute_returns + t_update), 5 epochs x 4 minibatches,
    ActorCritic 512-256-128 (BASELINE.json configs[0] shape), synthetic rollout produced by the policy itself."""
    import torch
    import torch.distributed as dist
    from rapid_locomotion_rl_b200.ppo import PPO, ActorCritic
    torch.manual_seed(0)
    ac = ActorCritic(42, 18, 630, 12, device=device)
    ppo = PPO(ac, device=device)
    ppo.init_storage(n_envs, T, [42], [18], [630], [12])
    obs = torch.randn(T + 1, n_envs, 42, device=device)
    priv = torch.rand(T + 1, n_envs, 18, device=device) * 2 - 1
    hist = torch.randn(n_envs, 630, device=device)
    bins = torch.zeros(n_envs, device=device)

    def rollout():
        for t in range(T):
            ppo.act(obs[t], priv[t], hist)
            ppo.process_env_step(torch.randn(n_envs, device=device) * 0.05, torch.rand(n_envs, device=device) < 0.01,
                                 {"env_bins": bins})
    times, t_gae = [], []
    local = int(os.environ.get("LOCAL_RANK", 0))
    sampler = ClockSampler(physical_gpu_index(local))
    warm = 2             # update() captures its CUDA graphs in the first call and uploads them at their first replay
    for it in range(iters + warm):
        rollout()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        if it == warm:
            sampler.__enter__()
        e0.record()
        ppo.compute_returns(obs[T], priv[T])
        e1.record()
        res = ppo.update()
        e2.record()
        torch.cuda.synchronize()
        if it >= warm:
            times.append(e0.elapsed_time(e2)); t_gae.append(e0.elapsed_time(e1))
    sampler.__exit__()
    # median of the timed iterations (each is one whole compute_returns + update; the mean is reported next to it)
    ms = sorted(times)[len(times) // 2]
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    samples = n_envs * T * world
    flops = 18.02e6 * n_envs * T       # 3.603 MFLOP per sample-visit x 5 epochs (per GPU)
    tf = flops / (ms * 1e-3) / 1e12
    return {"metric": "ppo_samples_per_s", "value": samples / (ms * 1e-3), "unit": "samples/s", "ms_per_iteration": ms,
            "ms_gae": sum(t_gae) / len(t_gae), "envs_per_gpu": n_envs, "steps_per_env": T, "epochs": 5, "minibatches": 4,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"] if "bf16_tflops_sustained" in pk else pk["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": tf / (pk.get("bf16_tflops_sustained") or pk["bf16_tflops"]), "traffic": None},
            "losses": list(res), "dtype": "bf16 operands, fp32 accumulate / master weights",
            "ms_per_iteration_all": [round(t, 3) for t in times], "clocks": sampler.summary()}


def runner_bench(n_envs, device, iters=10):
    """the whole training iteration through the drop-in Runner - 24 x [policy chain, Normal
    sample, fused env step, history push, storage writes] replayed from CUDA graphs, then GAE + PPO.update.
    The simulator is outside the path: its state tensors stay as they are between steps."""
    import torch
    from cases import build_case
    from rapid_locomotion_rl_b200.envs import HistoryWrapper, VelocityTrackingEasyEnv
    from rapid_locomotion_rl_b200.ppo import Runner
    cfg, robot, terrain = build_case("mc_flat", n_envs)
    env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device=device, headless=True, cfg=cfg, terrain=terrain, seed=0))
    torch.manual_seed(0)
    res = {}
    for graph in (True, False):
        r = Runner(env, device=device, graph_rollout=graph)
        # one learn() call: the first 7 iterations warm up (eager rollout, then one capture per history-ring
        # phase); every iteration ends with update()'s device->host read of the losses, so the per-iteration
        # host times are device-complete
        hist = r.learn(7 + iters)
        res["graph" if graph else "eager"] = sum(h["time_iter"] for h in hist[7:]) / iters
    return {"metric": "training env-steps/s (rollout + GAE + PPO.update, synthetic simulator state)", "envs": n_envs,
            "steps_per_env": 24, "value": n_envs * 24 / res["graph"], "ms_per_iteration": res["graph"] * 1e3,
            "ms_per_iteration_eager_rollout": res["eager"] * 1e3}


def reference_staged():
    """True when the unmodified reference is importable here: /root/reference (build container) or the copy staged
    by oracle/make_ref.py (the GPU box)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "_ref_shims"))
    import harness
    return harness.reference_available()


def reference_env_arm(envs, steps, warmup, robot="mini_cheetah"):
    """The UNMODIFIED reference on host cores: HistoryWrapper(VelocityTrackingEasyEnv).step (mini_gym/envs/
    mini_cheetah/velocity_tracking/velocity_tracking_easy_env.py:42-64, wrappers/history_wrapper.py:18-41, base/
    legged_robot.py:106-137) driven through the fake-simulator shims on the same synthetic state distribution as
    our arm.  Returns (env-steps/s, threads, seconds)."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "_ref_shims"))
    import harness
    from rapid_locomotion_rl_b200.sim import synthetic_state
    env, Cfg = harness.make_reference_env(robot, envs, device="cpu")
    e = env.env
    st = synthetic_state(0, envs, e.num_bodies, 12, e.default_dof_pos[0].numpy(), e.feet_indices.tolist(),
                         e.termination_contact_indices.tolist())
    T = torch.from_numpy
    e.all_root_states[:] = T(st["root_states"]); e.all_dof_state[:] = T(st["dof_state"].reshape(-1, 2))
    e.all_contact_forces[:] = T(st["contact_forces"].reshape(-1, 3))
    e.commands[:, :3] = torch.rand(envs, 3) * 2 - 1
    e.episode_length_buf[:] = torch.randint(0, 1001, (envs,))
    actions = torch.randn(envs, 12)
    for _ in range(warmup):
        env.step(actions)
    t0 = time.perf_counter()
    for _ in range(steps):
        env.step(actions)
    dt = time.perf_counter() - t0
    return envs * steps / dt, torch.get_num_threads(), dt


def cpu_env_arm(envs, steps, warmup, threads=None, device="cpu"):
    """Fallback when the reference is not staged, and the torch-eager-on-GPU leg: oracle/env_oracle.py (a torch
    restatement pinned bit-exactly to the reference).  Returns (env-steps/s, threads used, seconds)."""
    import numpy as np
    import torch
    import statekit
    from cases import build_case
    from oracle.env_oracle import OracleEnv
    from rapid_locomotion_rl_b200.sim import synthetic_state
    if threads:
        torch.set_num_threads(threads)
    cfg, robot, terrain = build_case("mc_flat", envs)
    o = OracleEnv(cfg, robot, terrain, device=device)
    st = synthetic_state(0, envs, robot.num_bodies, 12, o.default_dof_pos[0].cpu().numpy(), o.feet_indices.tolist(),
                         o.termination_contact_indices.tolist())
    statekit.apply_to_oracle(o, st)
    for d in (o.episode_sums, o.command_sums):
        for k in d:
            d[k] = d[k].to(device)
    o.commands[:, :3] = torch.rand(envs, 3, device=device) * 2 - 1
    actions = torch.randn(envs, 12, device=device)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    for _ in range(warmup):
        o.step(actions)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(actions)
    sync()
    dt = time.perf_counter() - t0
    return envs * steps / dt, torch.get_num_threads(), dt


def reference_ppo_arm(n_envs, T, device, epochs=5, iters=1):
    """The UNMODIFIED reference learner (mini_gym_learn/ppo/ppo.py:62-178, rollout_storage.py:76-139,
    actor_critic.py:23-173) on `device` ("cpu": host cores; "cuda:0": torch eager + cuBLAS, the same-box GPU number
    the fused learner replaces): rollout of synthetic observations through PPO.act / process_env_step, then the timed
    compute_returns + update.  Returns samples/s and what was run."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "_ref_shims"))
    import harness
    harness.install()
    import isaacgym  # noqa: F401  (fake)
    from mini_gym_learn.ppo import ActorCritic
    from mini_gym_learn.ppo.ppo import PPO, PPO_Args
    torch.manual_seed(0)
    PPO_Args.num_learning_epochs = epochs
    ac = ActorCritic(42, 18, 630, 12).to(device)
    ppo = PPO(ac, device=device)
    ppo.init_storage(n_envs, T, [42], [18], [630], [12])
    obs = torch.randn(T + 1, n_envs, 42, device=device)
    priv = torch.rand(T + 1, n_envs, 18, device=device) * 2 - 1
    hist = torch.randn(n_envs, 630, device=device)
    bins = torch.zeros(n_envs, device=device)
    sync = torch.cuda.synchronize if str(device) != "cpu" else (lambda: None)
    times = []
    for it in range(iters + (1 if str(device) != "cpu" else 0)):        # the GPU leg gets a warm-up iteration
        with torch.no_grad():
            for t in range(T):
                ppo.act(obs[t], priv[t], hist)
                ppo.process_env_step(torch.randn(n_envs, device=device) * 0.05, torch.rand(n_envs, device=device) < 0.01,
                                     {"env_bins": bins})
        sync()
        t0 = time.perf_counter()
        with torch.no_grad():
            ppo.compute_returns(obs[T], priv[T])
        res = ppo.update()
        sync()
        times.append(time.perf_counter() - t0)
    dt = times[-1] if str(device) != "cpu" else sum(times) / len(times)
    # samples/s of the FULL update (5 epochs): an arm run with fewer epochs is scaled by the epoch count, GAE included
    full = dt * (5.0 / epochs)
    return {"value": n_envs * T / full, "unit": "samples/s", "device": str(device), "ms_per_iteration": full * 1e3,
            "epochs_run": epochs, "threads": torch.get_num_threads(), "envs": n_envs, "steps_per_env": T,
            "losses": [float(x) for x in res],
            "what": "unmodified mini_gym_learn PPO.compute_returns + PPO.update (%d of 5 epochs x 4 minibatches timed%s)"
                    % (epochs, ", scaled to 5" if epochs != 5 else "")}


def _subprocess_json(argv, timeout=900):
    """Runs `python bench.py <argv>` and parses the JSON line it prints (the reference's Cfg / PPO_Args are
    process-global classes, and its torch thread settings should not leak into our arm)."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.abspath(__file__)] + argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=timeout)
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("no JSON from %s: %s" % (argv, r.stderr[-400:]))


def cpu_baseline(envs, target_s=15.0):
    """The reference arm inside our arm's line: a bounded number of steps of the SAME workload (same env count) through
    the unmodified reference on the host cores, plus the learner's reference numbers (host cores, and torch eager on the
    B200) - SURVEY.md 8(d).  Every leg runs in its own process."""
    import torch
    try:
        ref = _subprocess_json(["--impl", "reference", "--envs", str(envs), "--steps", "0", "--target-s", str(target_s)])
        out = dict(ref["cpu_baseline"])
    except Exception as exc:
        out = {"error": str(exc)[:300]}
    if torch.cuda.is_available():
        # the reference's torch-eager algorithm ON the B200 (the port issues the reference's ~300 small
        # ATen kernels per step), as the GPU number the fused kernel replaces
        try:
            vg, _, sg = cpu_env_arm(32768, 30, 5, device="cuda:0")
            out["torch_eager_gpu_port"] = {"value": vg, "unit": "env-steps/s", "sample": "32768 envs x 30 steps, oracle port on "
                                           "cuda:0 (torch eager fp32, %.2f s)" % sg}
        except Exception as exc:      # the baseline must never take the bench down
            out["torch_eager_gpu_port"] = {"error": str(exc)[:200]}
    return out


def ppo_reference_block(n_envs):
    """`ppo.reference`: the unmodified reference learner on identical synthetic rollouts - host cores (1 epoch timed,
    scaled to 5: an epoch is ~10 s of CPU at 4000 envs) and torch eager on cuda:0 (all 5 epochs)."""
    out = {}
    for key, argv in (("cpu", ["--impl", "reference-ppo", "--ppo-envs", str(n_envs), "--ref-device", "cpu", "--ref-epochs", "1"]),
                      ("torch_eager_gpu", ["--impl", "reference-ppo", "--ppo-envs", str(n_envs), "--ref-device", "cuda:0",
                                           "--ref-epochs", "5"])):
        try:
            out[key] = _subprocess_json(argv)
        except Exception as exc:
            out[key] = {"error": str(exc)[:300]}
    return out


def run_reference_ppo(args):
    if int(os.environ.get("RANK", 0)) != 0:
        return
    import torch
    if args.ref_device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    print(json.dumps(reference_ppo_arm(args.ppo_envs, 24, args.ref_device, epochs=args.ref_epochs)))


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    staged = reference_staged()
    arm = reference_env_arm if staged else cpu_env_arm
    kind = "reference" if staged else "port"
    envs = args.envs
    # one probe step sizes the run: the FULL workload (same env count as our arm) unless K steps of it would not
    # end within a few minutes - only then is each step a bounded sample of the workload
    _, _, probe = arm(envs, 1, 1)
    steps, warm = args.steps, max(3, args.warmup)
    if steps <= 0:                      # cpu_baseline leg of our arm: ~target_s seconds of CPU work
        steps, warm = max(5, int(args.target_s / max(probe, 1e-6))), 2
    sample = envs
    while (steps + warm) * probe * sample / envs > 240 and sample > 500:
        sample //= 2
    v, cores, secs = arm(sample, steps, warm)
    what = ("unmodified reference HistoryWrapper(VelocityTrackingEasyEnv).step through tests/_ref_shims" if staged else
            "oracle/env_oracle.py (pinned torch restatement; the reference is not staged here)")
    out = {
        "impl": "reference", "metric": "env_steps_per_s", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": secs / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "mini_cheetah_flat env step (PD torques + terminations + 12 reward terms + obs/noise/clip + state "
                               "+ observation history), %d envs per step on host cores" % sample,
                   "envs_per_gpu": envs, "same_config": sample == envs},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": kind,
                         "sample": "%d envs x %d steps, %s (torch CPU fp32, %.1f s)" % (sample, steps, what, secs)},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-ppo"])
    ap.add_argument("--target-s", type=float, default=15.0, help="reference arm with --steps 0: seconds of CPU work")
    ap.add_argument("--ref-device", default="cpu", help="--impl reference-ppo: device of the reference learner")
    ap.add_argument("--ref-epochs", type=int, default=1, help="--impl reference-ppo: epochs timed (scaled to 5)")
    ap.add_argument("--envs", type=int, default=32768, help="envs per GPU")
    ap.add_argument("--quick", action="store_true", help="skip the extra sizes and the PPO metric")
    ap.add_argument("--only-ppo", action="store_true", help="development aid: print only the PPO metric object")
    ap.add_argument("--ppo-envs", type=int, default=4000, help="envs per GPU for the PPO samples/s metric")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-ppo":
        run_reference_ppo(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
