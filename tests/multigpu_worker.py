"""torchrun worker for tests/test_multigpu.py: env-sharded PPO / GAE / GAC across ranks (NCCL) must equal
the single-process result on the concatenated envs.  Rank 0 prints "MULTIGPU OK"."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def fill(st, g):
    r = lambda t, s=1.0: t.copy_(torch.randn(t.shape, generator=g) * s)
    r(st["observations"]); st["privileged_observations"].copy_(torch.rand(st["privileged_observations"].shape, generator=g) * 2 - 1)
    r(st["observation_histories"]); r(st["actions"]); r(st["values"]); r(st["rewards"], 0.05)
    st["dones"].copy_((torch.rand(st["dones"].shape, generator=g) < 0.02).to(torch.uint8))
    st["actions_log_prob"].fill_(-17.0); r(st["mu"], 0.3); st["sigma"].fill_(1.0)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    dist.init_process_group("nccl", device_id=torch.device(dev))
    from rapid_locomotion_rl_b200.ppo import PPO, ActorCritic, PPO_Args
    from rapid_locomotion_rl_b200 import sharding
    T, N = 24, 96                     # envs per rank
    names = ("observations", "privileged_observations", "observation_histories", "actions", "values", "rewards", "dones",
             "actions_log_prob", "mu", "sigma")
    shapes = dict(observations=42, privileged_observations=18, observation_histories=630, actions=12, values=1, rewards=1,
                  dones=1, actions_log_prob=1, mu=12, sigma=12)
    g = torch.Generator().manual_seed(0)
    full = {k: torch.zeros(T, N * world, shapes[k], dtype=torch.uint8 if k == "dones" else torch.float32) for k in names}
    fill(full, g)
    last_values = torch.randn(N * world, 1, generator=g)
    perms = [torch.randperm(T * N, generator=g) for _ in range(world)]

    def make(n_envs, cols):
        torch.manual_seed(1)
        ac = ActorCritic(42, 18, 630, 12, device=dev)
        ppo = PPO(ac, device=dev)
        ppo.init_storage(n_envs, T, [42], [18], [630], [12])
        for k in names:
            getattr(ppo.storage, k).copy_(full[k][:, cols].to(dev))
        return ac, ppo

    # ---- sharded: this rank's envs ----
    s0, cnt = sharding.env_shard(N * world)
    assert (s0, cnt) == (rank * N, N)
    ac, ppo = make(N, slice(s0, s0 + cnt))
    ppo.storage.compute_returns(last_values[s0:s0 + cnt].to(dev), PPO_Args.gamma, PPO_Args.lam)
    mb = T * N // 4
    idx_local = perms[rank][:mb].to(dev)
    ppo.minibatch_step(idx_local, world, sharding.all_reduce_sum_)
    torch.cuda.synchronize()
    flat = ac.flat.clone()

    # ---- the same steps with the fused NVLink peer all-reduce instead of NCCL (csrc/peer_allreduce.cu) ----
    acp, ppop = make(N, slice(s0, s0 + cnt))
    ppop.storage.compute_returns(last_values[s0:s0 + cnt].to(dev), PPO_Args.gamma, PPO_Args.lam)
    ppop.enable_peer_allreduce()
    ppop.minibatch_step(idx_local, world, "peer")
    torch.cuda.synchronize()
    d1 = (acp.flat - flat).abs().max().item()
    assert d1 <= 2e-6, "peer all-reduce step differs from the NCCL step: %g" % d1
    for k in range(1, 4):                                   # more steps: call counters, zeroing, both segments
        idx_k = perms[rank][k * mb:(k + 1) * mb].to(dev)
        ppo.minibatch_step(idx_k, world, sharding.all_reduce_sum_)
        ppop.minibatch_step(idx_k, world, "peer")
    torch.cuda.synchronize()
    d4 = (acp.flat - ac.flat).abs().max().item()
    assert d4 <= 2e-4, "peer and NCCL paths drifted apart after 4 steps: %g" % d4
    peers_flat = [torch.zeros_like(acp.flat) for _ in range(world)]
    dist.all_gather(peers_flat, acp.flat)
    assert all(torch.equal(peers_flat[0], t) for t in peers_flat), "ranks diverged on the peer path"
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)

    # ---- the whole update() (CUDA-graph replay, adaptation module one call behind on the side branch): its gradient
    # summed over the side communicator (the default) must give what the single all-reduce per minibatch gives ----
    def run_update(side):
        os.environ["RL_PPO_SIDE_COMM"] = "1" if side else "0"
        acu, ppou = make(N, slice(s0, s0 + cnt))
        ppou.storage.compute_returns(last_values[s0:s0 + cnt].to(dev), PPO_Args.gamma, PPO_Args.lam)
        torch.manual_seed(5); torch.cuda.manual_seed(5)
        ppou.update()
        torch.cuda.synchronize()
        assert (ppou._ada_g is not None) == side
        out = acu.flat.clone()
        ppou.release_graphs()
        return out
    fu_side, fu_one = run_update(True), run_update(False)
    du = (fu_side - fu_one).abs().max().item()
    assert du <= 2e-5, "update() over the side communicator differs from the one-collective schedule: %g" % du
    both = [torch.zeros_like(fu_side) for _ in range(world)]
    dist.all_gather(both, fu_side)
    assert all(torch.equal(both[0], t) for t in both), "ranks diverged in update() (side communicator)"
    adv_local = ppo.storage.advantages.clone()
    adv_all = [torch.zeros_like(adv_local) for _ in range(world)]
    dist.all_gather(adv_all, adv_local)
    ok = True
    if rank == 0:
        for r in range(1, world):
            assert torch.equal(gathered[0], gathered[r]), "ranks diverged after one optimiser step"
        # ---- single process on the concatenated envs, same rows in the minibatch ----
        dist_world = world
        import rapid_locomotion_rl_b200.sharding as sh
        real = sh.world_size
        ac1, ppo1 = make(N * world, slice(0, N * world))
        try:
            sh._dist_backup = sh._dist
            sh._dist = lambda: None                       # single-process view
            ppo1.storage.compute_returns(last_values.to(dev), PPO_Args.gamma, PPO_Args.lam)
            adv_ref = ppo1.storage.advantages
            got = torch.cat(adv_all, dim=1)
            torch.testing.assert_close(got, adv_ref, rtol=1e-5, atol=1e-6)
            rows = []
            for r in range(dist_world):
                j = perms[r][:mb]
                t, n = j // N, j % N
                rows.append(t * (N * dist_world) + r * N + n)
            ppo1.minibatch_step(torch.cat(rows).to(dev), 1, None)
            torch.cuda.synchronize()
        finally:
            sh._dist = sh._dist_backup
        diff = (ac1.flat - flat).abs().max().item()
        assert diff <= 2.5e-3, "sharded step differs from the single-process step: %g" % diff     # |dw| <= lr per Adam step
        cos = torch.nn.functional.cosine_similarity(ac1.flat - ppo1_init(dev), flat - ppo1_init(dev), dim=0).item()
        assert cos > 0.98, cos
        print("MULTIGPU OK world=%d max|dw diff|=%.3g update cosine=%.4f; peer vs NCCL: %.3g after 1 step, %.3g after 4; "
              "update() side communicator vs one collective: %.3g" % (world, diff, cos, d1, d4, du))
    dist.barrier()
    dist.destroy_process_group()


def ppo1_init(dev):
    from rapid_locomotion_rl_b200.ppo import ActorCritic
    torch.manual_seed(1)
    return ActorCritic(42, 18, 630, 12, device=dev).flat.clone()


if __name__ == "__main__":
    main()
