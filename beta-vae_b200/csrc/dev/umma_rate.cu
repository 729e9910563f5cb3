// Development probe (not part of the library): sustained cycles per tcgen05.mma (kind::f16, bf16 -> fp32) on B200
// for the operand sources / shapes the convolution kernels can choose between.
//
//   mode ss      A and B K-major SWIZZLE_128B tiles in shared memory (fprop / dgrad form)
//   mode ss_mn   A and B MN-major SWIZZLE_128B tiles in shared memory (wgrad form)
//   mode ss_win  like ss, but A is read through halo windows: start shifted by 3 pixel rows (384 B), 8-row groups
//                1280 B apart (what conv_halo_tc_kernel issues)
//   mode ts      A in tensor memory (128 lanes x 8 columns per K = 16), B K-major in shared memory
//   cta_group 2  the same with a CTA pair: M = 256 (128 rows per CTA), each CTA holds N/2 rows of B
//
// Every CTA (one per SM, all 148 at once) issues NMMA back-to-back MMAs from one thread into one accumulator,
// commits to an mbarrier and waits; clock64() around issue + completion gives cycles per instruction in the SM clock
// domain (independent of DVFS).  Tensor-pipe floor per instruction: 128 x N x 16 MACs at 4096 MACs/clk/SM = N/2 cycles
// per CTA (M = 128 per CTA; M = 64 costs the same).  Printed: cycles/MMA (median and max over CTAs), floor, floor /
// measured (= the tensor-pipe duty cycle a kernel built from these instructions can reach), and the shared-memory
// bytes one instruction reads per CTA divided by the measured cycles.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_rate umma_rate.cu -lcuda
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cooperative_groups.h>

#include "../tc_ptx.cuh"

namespace bvae {
void set_error(const char* fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); fputc('\n', stderr); }
int cuda_fail(cudaError_t e, const char* what) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -3; }
unsigned long long g_launches = 0;
int num_sms() { return 148; }
}  // namespace bvae
using namespace bvae;

enum { MODE_SS = 0, MODE_SS_MN = 1, MODE_SS_WIN = 2, MODE_TS = 3 };

static constexpr int NMMA_BODY = 8;
static constexpr uint32_t A_BYTES = 65536, B_BYTES = 65536;

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 20); ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* dst_smem, uint32_t cols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t addr, uint32_t cols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
  else         asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  const uint32_t acc = 1;
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
  const uint32_t acc = 1;
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit_cg(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)1) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
}

// elect.sync: one lane of a converged warp.  The issuing warp runs its loops with all 32 lanes (warp index obtained
// through a shuffle, so the compiler knows it is warp-uniform) and predicates only the tcgen05 instructions on the
// elected lane: descriptors then live in uniform registers and consecutive UTCHMMAs are 2-3 SASS instructions apart.
// (With `if (threadIdx.x == 32)` around the loop, nvcc wraps every UTCHMMA in an ELECT / R2UR / BRA.U.ANY waterfall
// loop -- measured 74 cycles per instruction regardless of N, the bound the round-1 convolution kernels ran into.)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}

struct RateParams {
  int mode, n, iters;
  uint32_t idesc;
  long long* cycles;   // [grid]
  int* err;
};

template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(const RateParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_sm = smem;
  uint8_t* b_sm = smem + A_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));

  // operands: small bf16 values (pattern is irrelevant for timing, kept finite)
  for (uint32_t i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) {
    const uint32_t h = (i * 2654435761u) >> 28;                      // 0..15
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (h << 16) + h;  // bf16 pairs near 2^-7
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc_cg<CG>(tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  {
    // A operand slots for the TS form: columns 256..319 (8 slots x 8 columns), all 128 lanes
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int s = 0; s < 8; ++s) tmem_st8(tmem + lane_base + 256u + (uint32_t)s * 8u, 0x3C003C00u);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  tc_fence_after();

  const uint32_t warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp_u == 1 && rank == 0) {
    const bool leader = elect_one();
    uint64_t ad[NMMA_BODY], bd[NMMA_BODY];
    uint32_t at[NMMA_BODY];
    const uint32_t a0 = smem_u32(a_sm), b0 = smem_u32(b_sm);
#pragma unroll
    for (int s = 0; s < NMMA_BODY; ++s) {
      const uint32_t tile = (uint32_t)(s >> 2), k = (uint32_t)(s & 3);
      at[s] = tmem + 256u + (uint32_t)s * 8u;
      if (p.mode == MODE_SS_MN) {
        // [pixels][64 ch] blocks of 64 pixel rows (8 KB); M = 128 -> two blocks, N = n_cta -> n_cta / 64 blocks
        ad[s] = mnmajor_desc(a0 + tile * 16384u + k * 2048u, 8192u, 1024u);
        bd[s] = mnmajor_desc(b0 + tile * 32768u + k * 2048u, 8192u, 1024u);
      } else {
        if (p.mode == MODE_SS_WIN) {
          uint64_t d = (uint64_t)(((a0 + tile * 24576u + 384u + k * 32u) & 0x3FFFFu) >> 4);
          d |= 1ull << 16;
          d |= (uint64_t)(1280u >> 4) << 32;
          d |= 1ull << 46;
          d |= 2ull << 61;
          ad[s] = d;
        } else {
          ad[s] = kmajor_desc(a0 + tile * 16384u + k * 32u, 128);
        }
        bd[s] = kmajor_desc(b0 + tile * 32768u + k * 32u, 128);
      }
    }
    const long long t0 = clock64();
    if (p.mode == MODE_TS) {
      for (int it = 0; it < p.iters; ++it) {
        if (leader) {
#pragma unroll
          for (int s = 0; s < NMMA_BODY; ++s) mma_ts<CG>(tmem, at[s], bd[s], p.idesc);
        }
        __syncwarp();
      }
    } else {
      for (int it = 0; it < p.iters; ++it) {
        if (leader) {
#pragma unroll
          for (int s = 0; s < NMMA_BODY; ++s) mma_ss<CG>(tmem, ad[s], bd[s], p.idesc);
        }
        __syncwarp();
      }
    }
    if (leader) commit_cg<CG>(bar);
    __syncwarp();
    const bool ok = mbar_wait_bounded(bar, 0);
    const long long t1 = clock64();
    tc_fence_after();
    if (leader) {
      if (!ok) atomicExch(p.err, 1);
      p.cycles[blockIdx.x / CG] = t1 - t0;
    }
  }
  tc_fence_before();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc_cg<CG>(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// "conv replay": the exact tcgen05.mma sequence of conv_halo_tc_kernel<64, resident> for one K = 64 chunk -- 9 taps x 4
// k-steps, A = shifted windows of a 24 KB halo slot (8-row groups 1280 B apart), B = 9 resident [64][64] weight taps --
// with the kernel's features switched on one at a time (flags) to see which of them costs MMA rate:
//   1 real tap windows (else all taps read window 0)     2 nine distinct B taps (else tap 0)
//   4 tcgen05.commit to an mbarrier after every 36 MMAs   8 accumulator alternates between two TMEM column ranges
//  16 halo slot rotates over 4 slots                      32 first MMA of a tile overwrites (accumulate = 0)
// ---------------------------------------------------------------------------------------------------------------
struct ConvReplayParams {
  int flags, tiles, n;
  long long* cycles;
  int* err;
};

__device__ __forceinline__ void mma_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int FLAGS, int NN>
__global__ void __launch_bounds__(128, 1) conv_replay_kernel(const ConvReplayParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t B_TOTAL = 9u * (uint32_t)NN * 128u, A_SLOT = 24576u;
  uint8_t* b_sm = smem;
  uint8_t* a_sm = smem + B_TOTAL;
  constexpr uint32_t nslots = NN > 64 ? 3u : 4u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(a_sm + nslots * A_SLOT);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < (B_TOTAL + nslots * A_SLOT) / 4; i += blockDim.x) {
    const uint32_t h = (i * 2654435761u) >> 28;
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (h << 16) + h;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc_cg<1>(tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp_u == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t b_tap = (uint32_t)NN * 128u;
    constexpr bool f_win = FLAGS & 1, f_b = FLAGS & 2, f_commit = FLAGS & 4, f_d = FLAGS & 8, f_slot = FLAGS & 16,
               f_first = FLAGS & 32;
    const long long t0 = clock64();
    for (int t = 0; t < p.tiles; ++t) {
      const uint32_t slot = f_slot ? (uint32_t)t % nslots : 0u;
      uint64_t a0 = (uint64_t)(((smem_u32(a_sm) + slot * A_SLOT) & 0x3FFFFu) >> 4);
      a0 |= 1ull << 16;
      a0 |= (uint64_t)(1280u >> 4) << 32;
      a0 |= 1ull << 46;
      a0 |= 2ull << 61;
      const uint64_t b0 = kmajor_desc(smem_u32(b_sm), 128);
      const uint32_t d = tmem + (f_d ? (uint32_t)(t & 1) * (uint32_t)NN : 0u);
      if (leader) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const uint32_t win = f_win ? (uint32_t)((j / 3) * 10 + (j % 3)) * 128u : 0u;
          const uint32_t bo = f_b ? (uint32_t)j * b_tap : 0u;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_acc(d, a0 + (uint64_t)((win + k * 32) >> 4), b0 + (uint64_t)((bo + k * 32) >> 4), idesc,
                    (f_first && j == 0 && k == 0) ? 0u : 1u);
        }
        if (f_commit) commit_cg<1>(&bar[1]);     // nobody waits on it (phase flips are harmless)
      }
      __syncwarp();
    }
    if (leader) commit_cg<1>(&bar[0]);
    __syncwarp();
    const bool ok = mbar_wait_bounded(&bar[0], 0);
    const long long t1 = clock64();
    tc_fence_after();
    if (leader) {
      if (!ok) atomicExch(p.err, 1);
      p.cycles[blockIdx.x] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_cg<1>(tmem, 512);
}

static int run_conv_replay(int flags, int n, int tiles) {
  const int grid = 148;
  long long* d_cyc;
  int* d_err;
  cudaMalloc(&d_cyc, sizeof(long long) * grid);
  cudaMalloc(&d_err, sizeof(int));
  cudaMemset(d_err, 0, sizeof(int));
  ConvReplayParams p;
  p.flags = flags; p.tiles = tiles; p.n = n; p.cycles = d_cyc; p.err = d_err;
  const size_t smem = 9 * (size_t)n * 128 + (n > 64 ? 3 : 4) * 24576 + 1024 + 128;
  void (*kern)(const ConvReplayParams) =
      n == 64 ? (flags ? conv_replay_kernel<63, 64> : conv_replay_kernel<0, 64>)
              : (flags ? conv_replay_kernel<63, 128> : conv_replay_kernel<0, 128>);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<grid, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("conv_replay flags=%d N=%d : CUDA error %s\n", flags, n, cudaGetErrorString(e));
    return 1;
  }
  std::vector<long long> cyc(grid);
  int err = 0;
  cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_cyc);
  cudaFree(d_err);
  std::sort(cyc.begin(), cyc.end());
  const double nm = (double)tiles * 36.0;
  printf("conv_replay N=%3d flags=%2d [%s%s%s%s%s%s] : %6.1f cyc/MMA (max %6.1f)  floor %5.1f%s\n", n, flags,
         flags & 1 ? "win " : "", flags & 2 ? "9taps " : "", flags & 4 ? "commit " : "", flags & 8 ? "2acc " : "",
         flags & 16 ? "4slots " : "", flags & 32 ? "overwrite " : "", cyc[grid / 2] / nm, cyc[grid - 1] / nm, n / 2.0,
         err ? "  [TIMEOUT]" : "");
  fflush(stdout);
  return err;
}

static const char* mode_name(int m) {
  switch (m) {
    case MODE_SS: return "ss";
    case MODE_SS_MN: return "ss_mn";
    case MODE_SS_WIN: return "ss_win";
    default: return "ts";
  }
}

static int run_case(int cg, int mode, int m_total, int n, int iters, double* out_med) {
  const int sms = 148;
  const int grid = sms / cg * cg;
  long long* d_cyc;
  int* d_err;
  cudaMalloc(&d_cyc, sizeof(long long) * grid);
  cudaMalloc(&d_err, sizeof(int));
  cudaMemset(d_cyc, 0, sizeof(long long) * grid);
  cudaMemset(d_err, 0, sizeof(int));
  RateParams p;
  p.mode = mode; p.n = n; p.iters = iters; p.cycles = d_cyc; p.err = d_err;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m_total >> 4) << 24);
  if (mode == MODE_SS_MN) p.idesc |= (1u << 15) | (1u << 16);
  const size_t smem = A_BYTES + B_BYTES + 1024 + 64;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e;
  if (cg == 1) {
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e = cudaLaunchKernelEx(&cfg, rate_kernel<1>, p);
  } else {
    cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e = cudaLaunchKernelEx(&cfg, rate_kernel<2>, p);
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-7s cg%d M=%3d N=%3d : CUDA error %s\n", mode_name(mode), cg, m_total, n, cudaGetErrorString(e));
    return 1;
  }
  std::vector<long long> cyc(grid);
  int err = 0;
  cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_cyc);
  cudaFree(d_err);
  const int ncl = grid / cg;
  cyc.resize(ncl);
  std::sort(cyc.begin(), cyc.end());
  const double nm = (double)iters * NMMA_BODY;
  const double med = cyc[ncl / 2] / nm, mx = cyc[ncl - 1] / nm;
  const int m_cta = m_total / cg;
  const double floor_c = (m_cta < 128 ? 128 : m_cta) * (double)n / 256.0 / 1.0;   // cycles per instruction per CTA (pair works in parallel)
  const double a_bytes = (mode == MODE_TS) ? 0.0 : m_cta * 32.0;
  const double b_bytes = (double)(n / cg) * 32.0;
  const double tf = 2.0 * m_total * n * 16.0 / med * (148.0 / cg) * 1.965e9 / 1e12;
  printf("%-7s cg%d M=%3d N=%3d : %7.1f cyc/MMA (max %7.1f)  floor %6.1f  duty %5.1f%%  smem %5.0f B/MMA/CTA = %5.1f B/clk  "
         "-> %6.0f TFLOP/s @1965MHz%s\n",
         mode_name(mode), cg, m_total, n, med, mx, floor_c, 100.0 * floor_c / med, a_bytes + b_bytes, (a_bytes + b_bytes) / med, tf,
         err ? "  [TIMEOUT]" : "");
  if (out_med) *out_med = med;
  fflush(stdout);
  return err;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 512;
  const char* only = argc > 2 ? argv[2] : "all";
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("# %s, %d SMs; %d MMAs per CTA per case, all CTAs concurrently, K = 16 per instruction (bf16)\n", prop.name,
         prop.multiProcessorCount, iters * NMMA_BODY);
  const int ns[] = {16, 32, 64, 128, 256};
  int bad = 0;
  auto want = [&](const char* tag) { return !strcmp(only, "all") || !strcmp(only, tag); };
  if (want("cg1")) {
    for (int mode : {MODE_SS, MODE_SS_WIN, MODE_SS_MN, MODE_TS})
      for (int n : ns) {
        if (mode == MODE_SS_MN && n < 64) continue;
        bad |= run_case(1, mode, 128, n, iters, nullptr);
      }
    for (int mode : {MODE_SS, MODE_TS})
      for (int n : {64, 128, 256}) bad |= run_case(1, mode, 64, n, iters, nullptr);
  }
  if (want("cg2")) {
    for (int mode : {MODE_SS, MODE_SS_MN, MODE_TS})
      for (int n : {32, 64, 128, 256}) {
        if (mode == MODE_SS_MN && n < 128) continue;
        bad |= run_case(2, mode, 256, n, iters, nullptr);
      }
  }
  if (want("conv")) {
    for (int n : {64, 128})
      for (int flags : {0, 63}) bad |= run_conv_replay(flags, n, 56);
  }
  return bad;
}
