"""World-size-2 `gloo` tests of the env-sharded path, on CPU.

Each rank holds half of the envs / minibatch rows, computes its local quantities with the ORACLE (the
checker), exchanges them through rapid_locomotion_rl_b200.sharding - the same helper the product's
PPO.update / RolloutStorage.compute_returns / LeggedRobot._resample_commands call - and the result must
equal the single-process oracle on the concatenated data:
  * PPO: sum over ranks of d(local loss / world) == gradient of the global-mean loss (ppo.py:131-144);
  * GAE: normalisation from the all-reduced (sum, sumsq, count) == rollout_storage.py:89-90 on all envs;
  * GAC: weights after the all-reduced saturating update are bit-identical on both ranks and equal
    curriculum.py:110-119 applied to the concatenated env list.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# ---- host restatements of what the kernels do with the all-reduced buffers (checkers, test-only) ----
def normalise_from_moments(x, moments):
    """(x - mean) / (std_unbiased + 1e-8) from float64 (sum, sumsq, count) - the host restatement of what
    rl_gae_normalize does with the all-reduced statistics (rollout_storage.py:90)."""
    s, ss, n = (float(v) for v in moments)
    mean = s / n
    var = max(0.0, (ss - n * mean * mean) / (n - 1.0))
    return ((x.double() - mean) / (var ** 0.5 + 1e-8)).to(x.dtype)


def saturating_bump_(weights, hit_count, own_flag, step=0.2):
    """w <- min(1, w + step) applied k = hit_count + (own_flag > 0) times (order independent, so the
    all-reduced integer counters give the same float64 weights on every rank): what rl_gac_update_sample does."""
    k = hit_count.to(torch.int64) + (own_flag > 0).to(torch.int64)
    for _ in range(int(k.max().item()) if k.numel() else 0):
        m = k > 0
        weights[m] = torch.clamp(weights[m] + step, 0.0, 1.0)
        k = k - m.to(torch.int64)
    return weights


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _small_params(seed=0):
    from oracle.ppo_oracle import PARAM_ORDER
    g = torch.Generator().manual_seed(seed)
    dims = {"env_factor_encoder": [18, 32, 16, 18], "adaptation_module": [630, 32, 16, 18],
            "actor_body": [60, 64, 32, 16, 12], "critic_body": [60, 64, 32, 16, 1]}
    p = {"std": torch.ones(12)}
    for name, d in dims.items():
        for i in range(len(d) - 1):
            p["%s.%d.weight" % (name, 2 * i)] = torch.randn(d[i + 1], d[i], generator=g) / d[i] ** 0.5
            p["%s.%d.bias" % (name, 2 * i)] = torch.randn(d[i + 1], generator=g) * 0.1
    assert set(p) == set(PARAM_ORDER)
    return p


def _minibatch(rows, seed=1):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return {"obs": r(rows, 42), "priv": torch.rand(rows, 18, generator=g) * 2 - 1, "hist": r(rows, 630),
            "actions": r(rows, 12), "values": r(rows, 1), "returns": r(rows, 1), "advantages": r(rows, 1),
            "old_logp": r(rows, 1) - 15.0, "old_mu": r(rows, 12) * 0.3, "old_sigma": torch.ones(rows, 12)}


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ppo_oracle as po
        from oracle.env_oracle import OracleCurriculum
        from rapid_locomotion_rl_b200 import sharding
        res = {}
        assert sharding.world_size() == world and sharding.rank() == rank

        # ---- env sharding covers every env exactly once -------------------------------------------------
        cover = torch.zeros(4001, dtype=torch.int64)
        s0, cnt = sharding.env_shard(4001)
        cover[s0:s0 + cnt] += 1
        sharding.all_reduce_sum_(cover)
        res["cover_ok"] = bool((cover == 1).all())

        # ---- PPO gradient + statistics ------------------------------------------------------------------
        rows = 96
        mb = _minibatch(rows)
        s0, cnt = sharding.env_shard(rows)
        local = {k: v[s0:s0 + cnt] for k, v in mb.items()}
        p = {k: v.clone().requires_grad_(True) for k, v in _small_params().items()}
        loss, surr, vloss, kl = po.minibatch_losses(p, local)
        (loss / world).backward()                       # the loss kernel's inv_gb = 1 / (B_local * world)
        flat = torch.cat([p[k].grad.reshape(-1) if p[k].grad is not None else torch.zeros(p[k].numel())
                          for k in po.PARAM_ORDER])
        stats = torch.tensor([surr.item() * cnt, vloss.item() * cnt, kl.item() * cnt, float(cnt)], dtype=torch.float64)
        sharding.all_reduce_sum_(flat, stats)
        res["grad"], res["stats"] = flat, stats

        # ---- GAE: global normalisation from all-reduced moments --------------------------------------------
        g = torch.Generator().manual_seed(5)
        T, N = 24, 50
        rew, val = torch.randn(T, N, 1, generator=g) * 0.05, torch.randn(T, N, 1, generator=g)
        dones = (torch.rand(T, N, 1, generator=g) < 0.05).to(torch.uint8)
        last = torch.randn(N, 1, generator=g)
        e0, ec = sharding.env_shard(N)
        sl = slice(e0, e0 + ec)
        ret = torch.zeros(T, ec, 1)
        adv = 0
        for t in reversed(range(T)):                     # rollout_storage.py:78-86 on the local envs
            nxt = last[sl] if t == T - 1 else val[t + 1, sl]
            alive = 1.0 - dones[t, sl].float()
            delta = rew[t, sl] + alive * 0.99 * nxt - val[t, sl]
            adv = delta + alive * 0.99 * 0.95 * adv
            ret[t] = adv + val[t, sl]
        raw = ret - val[:, sl]
        mom = torch.tensor([raw.double().sum(), (raw.double() ** 2).sum(), float(raw.numel())], dtype=torch.float64)
        sharding.all_reduce_sum_(mom)
        res["adv"] = normalise_from_moments(raw, mom)
        res["ret"] = ret

        # ---- GAC: all-reduced incidence counters -> identical saturating update ---------------------------
        cur = OracleCurriculum(100, x_vel=(-10, 10, 51), y_vel=(-0.6, 0.6, 2), yaw_vel=(-10, 10, 51))
        cur.set_to([-0.6, -0.6, -1.0], [0.6, 0.6, 1.0])
        rng = np.random.RandomState(3)
        n_env = 400
        bins = rng.choice(np.nonzero(cur.weights)[0], n_env)
        ok = rng.rand(n_env) < 0.3
        e0, ec = sharding.env_shard(n_env)
        hit = torch.zeros(cur.weights.size, dtype=torch.int32)
        own = torch.zeros(cur.weights.size, dtype=torch.int32)
        for e in range(e0, e0 + ec):
            if ok[e]:
                b = bins[e]
                own[b] = 1
                near = np.logical_and(cur.grid >= cur.grid[:, [b]] - 0.5, cur.grid <= cur.grid[:, [b]] + 0.5).all(axis=0)
                hit += torch.from_numpy(near.astype(np.int32))
        sharding.all_reduce_sum_(hit, own)
        w = saturating_bump_(torch.from_numpy(cur.weights.copy()), hit, own)
        res["gac_w"] = w
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def sharded():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    return {k: out[k] for k in range(2)}


def test_env_shards_partition(sharded):
    assert sharded[0]["cover_ok"] and sharded[1]["cover_ok"]
    from rapid_locomotion_rl_b200.sharding import env_shard
    assert [env_shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert env_shard(32768, 3, 8) == (3 * 4096, 4096)


def test_ppo_gradient_allreduce_matches_global_batch(sharded):
    from oracle import ppo_oracle as po
    mb = _minibatch(96)
    p = {k: v.clone().requires_grad_(True) for k, v in _small_params().items()}
    loss, surr, vloss, kl = po.minibatch_losses(p, mb)
    loss.backward()
    ref = torch.cat([p[k].grad.reshape(-1) if p[k].grad is not None else torch.zeros(p[k].numel()) for k in po.PARAM_ORDER])
    for r in range(2):
        torch.testing.assert_close(sharded[r]["grad"], ref, rtol=1e-4, atol=1e-6)
        st = sharded[r]["stats"]
        assert st[3] == 96
        np.testing.assert_allclose(st[:3].numpy() / 96, [surr.item(), vloss.item(), kl.item()], rtol=1e-5)
    assert torch.equal(sharded[0]["grad"], sharded[1]["grad"])       # ranks stay in lock step


def test_gae_global_normalisation(sharded):
    from oracle import ppo_oracle as po
    g = torch.Generator().manual_seed(5)
    T, N = 24, 50
    rew, val = torch.randn(T, N, 1, generator=g) * 0.05, torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < 0.05).to(torch.uint8)
    last = torch.randn(N, 1, generator=g)
    ret, adv = po.compute_returns(rew, val, dones, last, 0.99, 0.95)
    got_adv = torch.cat([sharded[0]["adv"], sharded[1]["adv"]], dim=1)
    got_ret = torch.cat([sharded[0]["ret"], sharded[1]["ret"]], dim=1)
    assert torch.equal(got_ret, ret)
    torch.testing.assert_close(got_adv, adv, rtol=1e-5, atol=1e-6)


def test_gac_weights_identical_across_ranks(sharded):
    from oracle.env_oracle import OracleCurriculum
    cur = OracleCurriculum(100, x_vel=(-10, 10, 51), y_vel=(-0.6, 0.6, 2), yaw_vel=(-10, 10, 51))
    cur.set_to([-0.6, -0.6, -1.0], [0.6, 0.6, 1.0])
    rng = np.random.RandomState(3)
    bins = rng.choice(np.nonzero(cur.weights)[0], 400)
    ok = rng.rand(400) < 0.3
    cur.update(bins, ok.astype(np.float64), ok.astype(np.float64), 0.5, 0.5)
    assert torch.equal(sharded[0]["gac_w"], sharded[1]["gac_w"])
    assert np.array_equal(sharded[0]["gac_w"].numpy(), cur.weights)      # float64, bit-exact
