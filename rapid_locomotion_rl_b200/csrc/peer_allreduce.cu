// Gradient all-reduce over NVLink peer memory, fused with what follows it in PPO.update
// (mini_gym_learn/ppo/ppo.py:146-150: loss.backward(); clip_grad_norm_; optimizer.step() - on several GPUs the
// gradients of the env shards are summed first).
//
// The NCCL path costs ~50 us per 2.4 MB all-reduce at 8 GPUs (latency, not bandwidth) and PPO.update needs 40 of
// them.  Here every rank's gradient buffer, a staging buffer and a few flags live in memory that every other
// rank has mapped (CUDA IPC, NVLink / NVSwitch peer access), and ONE kernel per all-reduce does
//   phase 0  publish "my gradient is complete" (release, system scope); wait for every peer's flag
//   phase 1  reduce-scatter by LOADS: rank r sums slice r of every rank's gradient (fixed rank order, so the
//            result is bit-identical everywhere), writes it to its staging buffer, accumulates the squared norm
//            of its slice (the input of clip_grad_norm_), publishes "slice r reduced" + its partial norm
//   phase 2  wait for every peer's slice; all-gather by LOADS into the local reduced gradient; sum the partial
//            norms; zero the own gradient buffer for the next accumulation (safe: every rank has finished
//            reading it once all slices are published)
// so the sum, the gradient-norm reduction and the zero_grad travel in one launch, each rank moves only
// 2 (W-1)/W of the buffer over NVLink, and nothing but two flag waits separates the phases.
// Flags are monotonically increasing call counters: no reset, no ABA.
#include <stdint.h>

#include "rl_common.cuh"

namespace rl {

constexpr int PEER_THREADS = 256;

struct PeerArgs {
  RlPeerComm c;
  long long offset, n;       // segment [offset, offset + n) of the gradient buffer (floats; offset, n multiples of 4)
  long long norm_n;          // the squared norm covers the first norm_n floats of the segment
  float* out;                // local reduced gradient [n]
  double* norm2_out;         // local: sum of squares of the reduced gradient
  unsigned int step;         // this call's ordinal (> 0, same on every rank, strictly increasing); 0: use the device
                             // counter at local_ws + 8 (CUDA-graph replay: kernel arguments are frozen in a graph)
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {       // peer data: never from a stale L1 line
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}

// flags layout (unsigned int) inside each rank's shared flag block:
//   [0] ready (gradient complete)   [1] reduced (slice + partial norm published)   [2..3] partial norm (double)
//   [8] local: block ticket of phase 1   [9] local: block ticket of phase 2
__device__ __forceinline__ void wait_all(const RlPeerComm& c, int which, unsigned int step) {
  if (threadIdx.x < c.world) {
    const unsigned int* f = reinterpret_cast<const unsigned int*>(c.flags[threadIdx.x]) + which;
    unsigned int spins = 0;
    while ((int)(ld_acquire_sys(f) - step) < 0) {
      if (++spins > (1u << 26)) {
        printf("peer_allreduce: rank %d waited too long for rank %d (flag %d, step %u)\n", c.rank, (int)threadIdx.x, which, step);
        __trap();
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(PEER_THREADS)
peer_allreduce_kernel(const __grid_constant__ PeerArgs a) {
  const RlPeerComm& c = a.c;
  const int W = c.world, r = c.rank;
  unsigned int* my_flags = reinterpret_cast<unsigned int*>(c.flags[r]);
  __shared__ double s_red[PEER_THREADS / 32];
  __shared__ bool s_last;
  unsigned int* step_dev = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(c.local_ws) + 8);
  const unsigned int step = a.step ? a.step : *reinterpret_cast<volatile unsigned int*>(step_dev) + 1u;

  // ---- phase 0: my gradient (written by the previous kernels of this stream) is complete -----------------
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(my_flags + 0, step);
  }
  wait_all(c, 0, step);

  // ---- phase 1: slice r of the sum ---------------------------------------------------------------------------
  const long long n4 = a.n / 4;
  const long long chunk4 = (n4 + W - 1) / W;
  const long long lo4 = (long long)r * chunk4, hi4 = min(n4, lo4 + chunk4);
  float* my_stage = reinterpret_cast<float*>(c.stage[r]);
  double acc = 0.0;
  for (long long i = lo4 + (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i < hi4; i += (long long)gridDim.x * PEER_THREADS) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < W; ++p) {
      const float4 v = ld_peer4(reinterpret_cast<const float*>(c.grad[p]) + a.offset + 4 * i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(my_stage + 4 * i) = s;
    const long long e = 4 * i;
    if (e + 3 < a.norm_n) acc += (double)s.x * s.x + (double)s.y * s.y + (double)s.z * s.z + (double)s.w * s.w;
    else {
      if (e < a.norm_n) acc += (double)s.x * s.x;
      if (e + 1 < a.norm_n) acc += (double)s.y * s.y;
      if (e + 2 < a.norm_n) acc += (double)s.z * s.z;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < PEER_THREADS / 32; ++w) t += s_red[w];
    atomicAdd(reinterpret_cast<double*>(c.local_ws), t);                 // this rank's partial norm
    __threadfence_system();                                               // my stage writes before the ticket
    const unsigned int done = atomicAdd(my_flags + 8, 1u);
    s_last = done == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    my_flags[8] = 0;
    const double part = *reinterpret_cast<volatile double*>(c.local_ws);
    *reinterpret_cast<volatile double*>(my_flags + 2) = part;
    *reinterpret_cast<double*>(c.local_ws) = 0.0;
    __threadfence_system();
    st_release_sys(my_flags + 1, step);                                   // slice r + partial norm are published
  }

  // ---- phase 2: gather every slice, total norm, zero my gradient ------------------------------------------------
  wait_all(c, 1, step);
  for (long long i = (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * PEER_THREADS) {
    const int p = (int)(i / chunk4);
    const float4 v = ld_peer4(reinterpret_cast<const float*>(c.stage[p]) + 4 * i);
    *reinterpret_cast<float4*>(a.out + 4 * i) = v;
    // every rank has published its slice, i.e. has finished reading my gradient: zero it for the next backward
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(c.grad[r]) + a.offset + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double t = 0.0;
    for (int p = 0; p < W; ++p) {                                          // rank order: identical on every rank
      const unsigned int* f = reinterpret_cast<const unsigned int*>(c.flags[p]) + 2;
      unsigned int lo, hi;
      asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(f) : "memory");
      t += __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    }
    *a.norm2_out = t;
  }
  // the last block to finish advances the device step counter (every block read it on entry)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(my_flags + 9, 1u) == gridDim.x - 1) {
      my_flags[9] = 0;
      *step_dev = step;
    }
  }
}

}  // namespace rl

extern "C" int rl_peer_allreduce(const RlPeerComm* comm, int64_t offset, int64_t n, int64_t norm_n, float* out, double* norm2_out,
                                 uint32_t step, void* stream) {
  RL_REQUIRE(comm && out && norm2_out, RL_ERR_BAD_ARG, "rl_peer_allreduce: null argument");
  RL_REQUIRE(comm->world >= 1 && comm->world <= RL_PEER_MAX_RANKS && comm->rank >= 0 && comm->rank < comm->world, RL_ERR_BAD_ARG,
             "rl_peer_allreduce: world=%d rank=%d", comm->world, comm->rank);
  RL_REQUIRE(n > 0 && (n % 4) == 0 && (offset % 4) == 0, RL_ERR_BAD_ARG, "rl_peer_allreduce: offset=%lld n=%lld step=%u",
             (long long)offset, (long long)n, step);
  for (int p = 0; p < comm->world; ++p)
    RL_REQUIRE(comm->grad[p] && comm->stage[p] && comm->flags[p], RL_ERR_BAD_ARG, "rl_peer_allreduce: rank %d buffers missing", p);
  RL_REQUIRE(comm->local_ws, RL_ERR_BAD_ARG, "rl_peer_allreduce: local workspace missing");
  rl::PeerArgs a;
  a.c = *comm; a.offset = offset; a.n = n; a.norm_n = norm_n; a.out = out; a.norm2_out = norm2_out; a.step = step;
  // all CTAs must be co-resident (they wait for each other's peers): one CTA per SM at most
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  long long want = (n / 4 + rl::PEER_THREADS - 1) / rl::PEER_THREADS;
  const int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
  rl::peer_allreduce_kernel<<<grid, rl::PEER_THREADS, 0, (cudaStream_t)stream>>>(a);
  return rl::check_launch("peer_allreduce_kernel");
}

// enables direct loads / stores from the current device to `peer_device` (no-op when already enabled)
extern "C" int rl_enable_peer_access(int32_t peer_device) {
  int cur = 0, can = 0;
  cudaGetDevice(&cur);
  if (cur == peer_device) return RL_OK;
  cudaDeviceCanAccessPeer(&can, cur, peer_device);
  RL_REQUIRE(can, RL_ERR_UNSUPPORTED, "rl_enable_peer_access: device %d cannot access device %d", cur, peer_device);
  cudaError_t err = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (err == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); err = cudaSuccess; }
  RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(err));
  return RL_OK;
}
