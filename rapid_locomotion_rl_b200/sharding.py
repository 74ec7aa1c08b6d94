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


_side_group = None


def side_group():
    """A SECOND process group (its own NCCL communicator) for collectives issued from a side stream concurrently with
    those of the default group: the adaptation module's gradient all-reduce of PPO.update, which must not be ordered
    against the policy gradient's.  Collective: every rank must call it at the same point (first use)."""
    global _side_group
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return None
    if _side_group is None:
        _side_group = d.new_group(backend=d.get_backend())
    return _side_group


def side_all_reduce_sum_(t):
    """In-place SUM all-reduce over the side group (see side_group)."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return
    d.all_reduce(t, op=d.ReduceOp.SUM, group=side_group())


def broadcast_(*tensors, src=0):
    """In-place broadcast of each tensor from rank `src` (replica initialisation: PPO.sync_replicas); no-op for a
    single process."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return
    for t in tensors:
        d.broadcast(t, src=src)


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


class PeerComm:
    """Every rank's gradient buffer, staging buffer and flag block mapped into every rank (CUDA IPC over NVLink /
    NVSwitch peer access) for `rl_peer_allreduce` (csrc/peer_allreduce.cu): the gradient all-reduce of PPO.update
    as ONE kernel that also produces the gradient norm and zeroes the accumulation buffer, instead of an NCCL call
    followed by separate reduction / zeroing work.  Needs an initialised process group (handles are exchanged with
    all_gather_object) and all ranks on one node."""

    def __init__(self, n_floats, device):
        import ctypes as C
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor
        from . import _lib
        self._lib = _lib.lib()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > _lib.DEFINES["RL_PEER_MAX_RANKS"]:
            raise _lib.RlError("PeerComm supports at most %d ranks" % _lib.DEFINES["RL_PEER_MAX_RANKS"])
        self.n = (int(n_floats) + 3) // 4 * 4
        self.device = torch.device(device)
        # one allocation per rank: [gradient n | staging n | 64 flag words]
        self.block = torch.zeros(2 * self.n + 64, dtype=torch.float32, device=self.device)
        torch.cuda.synchronize(self.device)
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.device.index, reduce_tensor(self.block)))
        self.peers = []
        with torch.cuda.device(self.device):
            for p in range(self.world):
                if p == self.rank:
                    self.peers.append(self.block)
                    continue
                src_dev, (rebuild, args) = handles[p]
                _lib.check(self._lib.rl_enable_peer_access(int(src_dev)))
                # open the IPC handle with OUR device current (argument 6 of rebuild_cuda_tensor is the device the
                # storage is opened on): the mapping is then addressable by kernels of this device over NVLink
                args = list(args)
                args[6] = self.device.index
                self.peers.append(rebuild(*args))          # cudaIpcOpenMemHandle, peer access enabled lazily
        self.grad = self.block[:self.n]
        self.local_ws = torch.zeros(4, dtype=torch.float32, device=self.device)      # 16 B: norm accumulator, call counter
        self.norm2 = torch.zeros(1, dtype=torch.float64, device=self.device)
        c = _lib.RlPeerComm()
        for p, t in enumerate(self.peers):
            base = t.data_ptr()
            c.grad[p], c.stage[p], c.flags[p] = base, base + 4 * self.n, base + 8 * self.n
        c.local_ws, c.world, c.rank = self.local_ws.data_ptr(), self.world, self.rank
        self._c = c
        dist.barrier()          # nobody starts reducing before every rank has mapped every buffer

    def all_reduce(self, out, norm_n=0, start=0):
        """out[start:n] = sum over ranks of that range of their gradient buffers (start is rounded down to a multiple
        of 4); self.norm2 = squared norm of the first norm_n floats of the range; the own range is zeroed.  Uses the
        device-side call counter, so the launch can live in a CUDA graph."""
        import ctypes as C
        from . import _lib
        assert out.dtype == torch.float32 and out.numel() >= self.n and out.is_contiguous()
        start = int(start) // 4 * 4
        _lib.check(self._lib.rl_peer_allreduce(C.byref(self._c), start, self.n - start, int(norm_n), out.data_ptr() + 4 * start,
                                               self.norm2.data_ptr(), 0, _lib.current_stream()))
