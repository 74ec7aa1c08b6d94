"""Env-sharded data parallelism: one process per GPU, rank r owns the env rows
[r * N_local, (r + 1) * N_local).  The per-step path (K1-K8) needs no collective; the only exchange steps
are sum all-reduces of small buffers, all of which go through `all_reduce_sum_`:

  * PPO: the flat policy gradient + the 4 loss statistics, once per optimiser step.  Each rank's loss kernel
    already scales by 1 / (B_local * world), so the SUM over ranks is the global-mean gradient
    (mini_gym_learn/ppo/ppo.py:131,140,144 take means over the minibatch);
  * GAE: (sum, sum of squares, count) of the raw advantages between the scan and the normalisation
    (rollout_storage.py:89-90 normalises over ALL T x N samples, unbiased std);
  * GAC: the int32 incidence counters / own-bin flags between scatter and update, so every rank applies the
    identical saturating weight update (curriculum.py:115-119) and the float64 weights stay bit-identical.

The reference is single-GPU; with world == 1 every function here is the identity.
"""
import torch


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def world_size():
    d = _dist()
    return d.get_world_size() if d is not None else 1


def rank():
    d = _dist()
    return d.get_rank() if d is not None else 0


def all_reduce_sum_(*tensors):
    """In-place SUM all-reduce of each tensor over the default process group (NCCL on GPUs, gloo in the
    CPU tests); no-op for a single process."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return
    for t in tensors:
        d.all_reduce(t, op=d.ReduceOp.SUM)


def all_reduce_max_(t):
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.MAX)
    return t


def env_shard(num_envs_global, rank_=None, world=None):
    """(first env row, number of env rows) owned by `rank_`: contiguous equal blocks; the remainder rows
    go one each to the lowest ranks."""
    world = world_size() if world is None else world
    rank_ = rank() if rank_ is None else rank_
    base, rem = divmod(int(num_envs_global), world)
    start = rank_ * base + min(rank_, rem)
    return start, base + (1 if rank_ < rem else 0)
