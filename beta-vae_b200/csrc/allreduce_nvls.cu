// Gradient all-reduce over NVLink 5 / NVSwitch with in-switch reduction (NVLS multicast), one kernel, in place.
//
// The data-parallel exchange of the training step (the reference has no multi-GPU code) is ONE sum of
// the flat fp32 gradient buffer (38.3 MB at the default config) across the ranks of a node.  NCCL's all-reduce of
// that size costs ~185 us on 8 x B200 (measured: profiles/r2_scale_n8_nccl.json) -- latency- rather than
// bandwidth-dominated (halving the bytes with bf16 saves 28 us).  With the buffer in symmetric memory bound to an
// NVSwitch multicast object, a rank reduces ITS 1/W slice with `multimem.ld_reduce` (the switch pulls the W copies,
// adds them in fp32 and returns one value) and broadcasts the result with `multimem.st` (one store lands in every
// rank's copy).  Per rank: ~1/W of the buffer read reduced, 1/W written multicast; every link carries ~size bytes each
// way; no staging, no second kernel.
//
// Synchronisation: block b of rank r only talks to block b of the other ranks.  Two barriers per launch over a
// symmetric flag buffer (flag[b][src] on every rank, monotonically increasing epoch values, st.release.sys /
// ld.acquire.sys): (1) every rank has ENTERED the kernel -- i.e. its backward pass, which precedes the kernel on its
// stream, has produced the gradients; (2) every rank's block b has finished its multicast stores.  The kernel retires
// when all of its blocks have passed (2), so everything the next kernel on this stream reads is complete.  The epoch
// lives in device memory (incremented by the last block to finish), so the launch is CUDA-graph replayable.  A rank
// that never arrives would hang its peers: the spins are bounded (~2 s) and report through `status`.
//
// The kernel needs no shared memory and 64 CTAs x 512 threads, so it can co-reside with the persistent tcgen05 kernels
// (one 192-thread CTA per SM, ~200 KB of shared memory each) -- which NCCL's kernels cannot; launching it per
// backward-ready bucket on a side stream is implemented (BVAE_NVLS_OVERLAP=1) but loses to one launch after backward,
// because a forked CUDA graph costs more than the hidden exchange gains.
// allreduce_p2p_kernel below is the same exchange over plain peer mappings; the engine picks per world size.
#include "common.cuh"

namespace bvae {

static constexpr int AR_THREADS = 512;
static constexpr int AR_MAX_WORLD = 16;

__device__ __forceinline__ void ar_st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ar_ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ar_mc_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void ar_mc_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

struct ArParams {
  float* mc;                      // multicast address of the first element of the range
  unsigned long long n4;          // float4 elements in the range
  uint32_t* const* flags;         // device array [world]: every rank's flag buffer [nblocks][world] (symmetric memory)
  uint32_t* epoch;                // this rank's launch counter (device memory)
  uint32_t* done;                 // this rank's finished-block counter
  int* status;                    // != 0: a barrier timed out
  int rank, world;
};

// all ranks' block `b`: arrive with `val`, wait for everyone
__device__ __forceinline__ void ar_barrier(const ArParams& p, uint32_t val) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    ar_st_release(p.flags[peer] + (size_t)blockIdx.x * p.world + p.rank, val);
    const uint32_t* mine = p.flags[p.rank] + (size_t)blockIdx.x * p.world + peer;
    const long long t0 = clock64();
    while ((int32_t)(ar_ld_acquire(mine) - val) < 0) {
      if (clock64() - t0 > 4000000000LL) {          // ~2 s at 2 GHz: a peer is missing
        atomicExch(p.status, 1);
        break;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(AR_THREADS) allreduce_nvls_kernel(const ArParams p) {
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(p.epoch);      // same value in every block of this launch
  ar_barrier(p, 2 * e + 1);
  // rank r owns float4s [r * slice, min((r + 1) * slice, n4)) of the range
  const unsigned long long slice = (p.n4 + p.world - 1) / p.world, lo = slice * p.rank;
  const unsigned long long per = lo >= p.n4 ? 0ull : (p.n4 - lo < slice ? p.n4 - lo : slice);
  float* base = p.mc + 4 * lo;
  const unsigned long long stride = (unsigned long long)gridDim.x * AR_THREADS;
  unsigned long long i = (unsigned long long)blockIdx.x * AR_THREADS + threadIdx.x;
  for (; i + 7 * stride < per; i += 8 * stride) {    // eight independent multicast reductions in flight per thread
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ar_mc_ld_reduce(base + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < 8; ++u) ar_mc_st(base + 4 * (i + u * stride), v[u]);
  }
  for (; i < per; i += stride) ar_mc_st(base + 4 * i, ar_mc_ld_reduce(base + 4 * i));
  ar_barrier(p, 2 * e + 2);
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.done, 1u) == gridDim.x - 1) {    // last block of this launch: every block has read `epoch`
      *p.done = 0;
      *p.epoch = e + 1;
    }
  }
}

// ---- the same exchange over plain peer mappings (NVLink loads / stores through the switch, no multicast object) ------
// Two-shot: rank r sums ITS 1/W slice by loading the W copies through the peer pointers of the symmetric buffer (W
// independent 16-byte loads in flight per element; summed in rank order 0..W-1, so the result does not depend on who
// computes it) and stores the sum into every rank's copy.  Per rank (W-1)/W of the buffer crosses the links each way --
// what the multicast form moves too -- but nothing waits on the switch's reduction engine: measured
// (profiles/r2_scale_p2p_*.json) against allreduce_nvls_kernel at 2 and 8 ranks.  Same flag barriers, same state.
struct ArP2pParams {
  float* const* bufs;             // device array [world]: every rank's buffer as mapped on this device (symmetric memory)
  unsigned long long off4, n4;    // float4 offset / length of the range
  uint32_t* const* flags;
  uint32_t* epoch;
  uint32_t* done;
  int* status;
  int rank, world;
};
__device__ __forceinline__ float4 ar_ld_peer(const float4* p) {
  float4 v;   // system-scope relaxed load: never served from a stale L1 line of an earlier step
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ar_st_peer(float4* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int W>
__global__ void __launch_bounds__(AR_THREADS) allreduce_p2p_kernel(const ArP2pParams p) {
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(p.epoch);
  ArParams bp;                    // the barrier only needs the flag / rank fields
  bp.flags = p.flags; bp.status = p.status; bp.rank = p.rank; bp.world = p.world;
  ar_barrier(bp, 2 * e + 1);
  const unsigned long long slice = (p.n4 + W - 1) / W, lo = slice * p.rank;
  const unsigned long long per = lo >= p.n4 ? 0ull : (p.n4 - lo < slice ? p.n4 - lo : slice);
  float4* base[W];
#pragma unroll
  for (int q = 0; q < W; ++q) base[q] = reinterpret_cast<float4*>(p.bufs[q]) + p.off4 + lo;
  const unsigned long long stride = (unsigned long long)gridDim.x * AR_THREADS;
  constexpr int U = W >= 8 ? 1 : (W >= 4 ? 2 : 4);      // U * W loads in flight per thread
  unsigned long long i = (unsigned long long)blockIdx.x * AR_THREADS + threadIdx.x;
  for (; i + (U - 1) * stride < per; i += U * stride) {
    float4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < W; ++q) v[u][q] = ar_ld_peer(base[q] + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 a = v[u][0];
#pragma unroll
      for (int q = 1; q < W; ++q) {
        a.x += v[u][q].x; a.y += v[u][q].y; a.z += v[u][q].z; a.w += v[u][q].w;
      }
#pragma unroll
      for (int q = 0; q < W; ++q) ar_st_peer(base[q] + i + u * stride, a);
    }
  }
  for (; i < per; i += stride) {
    float4 a = ar_ld_peer(base[0] + i);
#pragma unroll
    for (int q = 1; q < W; ++q) {
      const float4 t = ar_ld_peer(base[q] + i);
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
#pragma unroll
    for (int q = 0; q < W; ++q) ar_st_peer(base[q] + i, a);
  }
  ar_barrier(bp, 2 * e + 2);
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.done, 1u) == gridDim.x - 1) {
      *p.done = 0;
      *p.epoch = e + 1;
    }
  }
}

}  // namespace bvae

using namespace bvae;

/* Bytes of the per-rank state: [epoch, done, status, pad]. */
extern "C" size_t bvae_allreduce_nvls_state_bytes(void) { return 64; }
/* uint32 entries of the symmetric flag buffer every rank must provide (zero-initialised once). */
extern "C" size_t bvae_allreduce_nvls_flag_count(int nblocks, int world) { return (size_t)nblocks * (size_t)world; }

/* In-place sum over the ranks of a node of a symmetric fp32 buffer bound to an NVLS multicast object.
 *   mc_ptr      multicast address of the first float of the range (16-byte aligned; every rank passes its own mapping
 *               of the same object, same offset)
 *   n           floats in the range; a multiple of 4
 *   flag_ptrs   DEVICE array of `world` pointers: rank q's flag buffer as mapped on this device
 *   state       this rank's bvae_allreduce_nvls_state_bytes() of zero-initialised device memory
 * Every rank must launch it once per exchange, with the same nblocks (grid) on every rank. */
extern "C" int bvae_allreduce_nvls(void* mc_ptr, size_t n, void* flag_ptrs, void* state, int rank, int world, int nblocks,
                                   bvae_stream_t stream) {
  BVAE_REQUIRE(mc_ptr && flag_ptrs && state, BVAE_ERR_ARG, "allreduce_nvls: null pointer");
  BVAE_REQUIRE(world >= 2 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, BVAE_ERR_ARG, "allreduce_nvls: rank %d / %d",
               rank, world);
  BVAE_REQUIRE(n > 0 && n % 4 == 0 && (reinterpret_cast<uintptr_t>(mc_ptr) & 15) == 0, BVAE_ERR_ARG,
               "allreduce_nvls: range of %zu floats must be 16-byte aligned and a multiple of 4 floats", n);
  BVAE_REQUIRE(nblocks >= 1 && nblocks <= 148, BVAE_ERR_ARG, "allreduce_nvls: nblocks %d", nblocks);
  ArParams p;
  p.mc = reinterpret_cast<float*>(mc_ptr);
  p.n4 = n / 4;
  p.flags = reinterpret_cast<uint32_t* const*>(flag_ptrs);
  uint32_t* st32 = reinterpret_cast<uint32_t*>(state);
  p.epoch = st32;
  p.done = st32 + 1;
  p.status = reinterpret_cast<int*>(st32 + 2);
  p.rank = rank;
  p.world = world;
  allreduce_nvls_kernel<<<nblocks, AR_THREADS, 0, as_stream(stream)>>>(p);
  BVAE_LAUNCH_CHECK("allreduce_nvls");
  return BVAE_OK;
}

/* The same in-place sum through plain peer mappings of the symmetric buffer (no multicast object needed).
 *   buf_ptrs    DEVICE array of `world` pointers: rank q's buffer (element 0) as mapped on this device
 *   off, n      first float and number of floats of the range (multiples of 4); world in {2, 4, 8}
 * flag_ptrs / state / nblocks as for bvae_allreduce_nvls (the two entry points may share them). */
extern "C" int bvae_allreduce_p2p(void* buf_ptrs, size_t off, size_t n, void* flag_ptrs, void* state, int rank, int world,
                                  int nblocks, bvae_stream_t stream) {
  BVAE_REQUIRE(buf_ptrs && flag_ptrs && state, BVAE_ERR_ARG, "allreduce_p2p: null pointer");
  BVAE_REQUIRE((world == 2 || world == 4 || world == 8) && rank >= 0 && rank < world, BVAE_ERR_UNSUPPORTED,
               "allreduce_p2p: rank %d / %d (2, 4 or 8 ranks)", rank, world);
  BVAE_REQUIRE(n > 0 && n % 4 == 0 && off % 4 == 0, BVAE_ERR_ARG, "allreduce_p2p: range must be whole float4s");
  BVAE_REQUIRE(nblocks >= 1 && nblocks <= 148, BVAE_ERR_ARG, "allreduce_p2p: nblocks %d", nblocks);
  ArP2pParams p;
  p.bufs = reinterpret_cast<float* const*>(buf_ptrs);
  p.off4 = off / 4;
  p.n4 = n / 4;
  p.flags = reinterpret_cast<uint32_t* const*>(flag_ptrs);
  uint32_t* st32 = reinterpret_cast<uint32_t*>(state);
  p.epoch = st32;
  p.done = st32 + 1;
  p.status = reinterpret_cast<int*>(st32 + 2);
  p.rank = rank;
  p.world = world;
  cudaStream_t st = as_stream(stream);
  if (world == 2) allreduce_p2p_kernel<2><<<nblocks, AR_THREADS, 0, st>>>(p);
  else if (world == 4) allreduce_p2p_kernel<4><<<nblocks, AR_THREADS, 0, st>>>(p);
  else allreduce_p2p_kernel<8><<<nblocks, AR_THREADS, 0, st>>>(p);
  BVAE_LAUNCH_CHECK("allreduce_p2p");
  return BVAE_OK;
}
