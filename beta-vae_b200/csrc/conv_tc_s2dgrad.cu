// Input gradient of the stride-2 3x3 convolutions (encoder ConvBlock, src/models/beta_vae.py:37,42: what
// convolution_backward computes for dx) on tcgen05, halo formulation.
//
// dx has four output parity classes (ph, pw) = (Y & 1, X & 1); class (ph, pw) of dx[2y+ph, 2x+pw] receives only the
// taps (r, s) with (ph + 1 - r) and (pw + 1 - s) even, read at dy[y + (ph+1-r)/2, x + (pw+1-s)/2]:
//     (0,0): 1 tap     (0,1): 2 taps     (1,0): 2 taps     (1,1): 4 taps        -- nine in all, offsets in {0, 1}^2.
// So one (17 x 9)-pixel x 64-channel halo of dy per (16 x 8 low-res tile, K chunk) -- one TMA box, out-of-image rows
// zero-filled -- feeds all four classes through shifted K-major windows (conv_tc_halo.cu explains the window trick),
// each class accumulating in its own TMEM columns, and the epilogue scatters class (ph, pw) to the odd / even positions
// of dx.  The tap-per-TMA-load kernel this replaces (conv_tc.cu) launched one single-tile CTA per (tile, class) --
// 1024 CTAs whose TMEM allocation / barrier prologue cost more than their one to four MMA groups (enc1: 28 us).
// CTAs are persistent; the weights of the CTA's classes stay resident in shared memory: all four classes (nine taps)
// when K <= 128; for K <= 256 two classes per CTA, paired {(0,0), (1,1)} and {(0,1), (1,0)} = 5 and 4 taps so that the
// two kinds of CTA carry nearly the same work (one class per CTA left the 1-tap CTAs idle behind the 4-tap ones).
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace bvae {

static constexpr int SD_THREADS = 192;
static constexpr int SD_TH = 16, SD_TW = 8;                    // low-res (dy) tile
static constexpr int SD_BN = 64;                               // dx channels per unit (UMMA N)
static constexpr uint32_t SD_ROW = 128;                        // one pixel = 64 bf16 channels
static constexpr uint32_t SD_TAP_BYTES = SD_BN * SD_ROW;       // one [64][64] weight tap
static constexpr uint32_t SD_BW = SD_TW + 1, SD_BH = SD_TH + 1;
static constexpr uint32_t SD_BYTES = SD_BH * SD_BW * SD_ROW;   // 19584
static constexpr uint32_t SD_SLOT = 20480;
static constexpr uint32_t SD_SBO = SD_BW * SD_ROW;             // 1152

struct SdParams {
  int kchunks;                    // K / 64 (K = channels of dy)
  int n_tiles, nfix;              // N / 64; units that differ only in (class group, n tile)
  int tiles_w, tiles_h, m_tiles;  // low-res tiling
  int units;
  int h, w, N;                    // low-res extent; channels of dx
  int stages;
  uint32_t a_off, bar_off;
  __nv_bfloat16* dst;
};

struct SdMaps {
  CUtensorMap dy;   // [B][h][w][K]
  CUtensorMap w;    // [9][N][K]: the dgrad ("transposed") pack
};

// class = ph * 2 + pw; local tap j of a class -> weight tap (r * 3 + s) and halo window offset (di, dj)
__host__ __device__ constexpr int sd_ntaps(int cls) { return ((cls >> 1) ? 2 : 1) * ((cls & 1) ? 2 : 1); }
__host__ __device__ constexpr int sd_slot0(int cls) { return cls == 0 ? 0 : (cls == 1 ? 1 : (cls == 2 ? 3 : 5)); }
__host__ __device__ constexpr int sd_r(int cls, int j) { return (cls >> 1) ? (((cls & 1) ? (j >> 1) : j) ? 2 : 0) : 1; }
__host__ __device__ constexpr int sd_s(int cls, int j) { return (cls & 1) ? (((cls >> 1) ? (j & 1) : j) ? 2 : 0) : 1; }
__host__ __device__ constexpr int sd_di(int cls, int j) { return sd_r(cls, j) == 0 ? 1 : 0; }
__host__ __device__ constexpr int sd_dj(int cls, int j) { return sd_s(cls, j) == 0 ? 1 : 0; }
__host__ __device__ constexpr uint32_t sd_win(int cls, int j) {
  return (uint32_t)(sd_di(cls, j) * (int)SD_BW + sd_dj(cls, j)) * SD_ROW;
}

__device__ __forceinline__ void sd_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t sd_kdesc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= 1ull << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// D fp32, A/B bf16 K-major, M = 128, N = 64
static constexpr uint32_t SD_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(SD_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// the MMAs of one class for one K chunk (unrolled, constant window / slot offsets)
template <int CLS>
__device__ __forceinline__ void sd_issue_class(uint32_t acc, uint64_t a_desc0, uint64_t b_desc0, int slot0, uint32_t acc_on) {
#pragma unroll
  for (int j = 0; j < sd_ntaps(CLS); ++j) {
    constexpr uint32_t dummy = 0;
    (void)dummy;
    const uint32_t win = sd_win(CLS, j);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(acc, a_desc0 + (uint64_t)((win + k * 32) >> 4),
                b_desc0 + (uint64_t)((((uint32_t)(slot0 + j) * SD_TAP_BYTES) >> 4) + k * 2), SD_IDESC, (j | k) != 0 ? 1u : acc_on);
  }
}

// class held in accumulator i of class group grp
__host__ __device__ constexpr int sd_class_of(int cg, int grp, int i) {
  return cg == 4 ? i : (grp == 0 ? (i == 0 ? 0 : 3) : (i == 0 ? 1 : 2));
}

// CG = classes per CTA: 4 (all; nine resident taps per K chunk) or 2 (the balanced pairs; five tap slots per chunk)
template <int CG>
__global__ void __launch_bounds__(SD_THREADS, 1)
    conv_s2_dgrad_halo_kernel(const __grid_constant__ SdMaps maps, const SdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_base = smem;
  uint8_t* a_ring = smem + p.a_off;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* a_empty = a_full + p.stages;
  uint64_t* acc_full = a_empty + p.stages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* b_res = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_res + 1);

  constexpr int NGRP = 4 / CG;
  constexpr int NSLOT = CG == 4 ? 9 : 5;            // weight slots per K chunk
  constexpr uint32_t TMEM_COLS = 2 * CG * SD_BN;    // 512 / 256
  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  const int fix = blockIdx.x % p.nfix;              // gridDim.x is a multiple of nfix: fixed for all units of the CTA
  const int grp = fix % NGRP, n0 = (fix / NGRP) * SD_BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    mbar_init(b_res, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // producer: all lanes run the loop, the elected lane issues the TMA loads (see elect_one() in tc_ptx.cuh)
    const bool leader = elect_one();
    if (leader) {
      tma_prefetch_desc(&maps.dy);
      tma_prefetch_desc(&maps.w);
      int ntap_cta = 0;
      for (int i = 0; i < CG; ++i) ntap_cta += sd_ntaps(sd_class_of(CG, grp, i));
      mbar_expect_tx(b_res, (uint32_t)(ntap_cta * p.kchunks) * SD_TAP_BYTES);
      for (int c = 0; c < p.kchunks; ++c)
        for (int i = 0, slot0 = 0; i < CG; ++i) {
          const int cls = sd_class_of(CG, grp, i);
          if (i > 0) slot0 += sd_ntaps(sd_class_of(CG, grp, i - 1));   // CG == 4: 0, 1, 3, 5 = sd_slot0
          for (int j = 0; j < sd_ntaps(cls); ++j)
            tma_load_3d(&maps.w, b_res, b_base + (size_t)(c * NSLOT + slot0 + j) * SD_TAP_BYTES, c * 64, n0,
                        sd_r(cls, j) * 3 + sd_s(cls, j));
        }
    }
    __syncwarp();
    uint32_t ia = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      int mt = u / p.nfix;
      const int w0 = (mt % p.tiles_w) * SD_TW;
      mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * SD_TH;
      const int b = mt / p.tiles_h;
      for (int c = 0; c < p.kchunks; ++c, ++ia) {
        const uint32_t sa = ia % p.stages;
        mbar_wait(&a_empty[sa], ((ia / p.stages) & 1u) ^ 1u);
        if (leader) {
          mbar_expect_tx(&a_full[sa], SD_BYTES);
          tma_load_4d(&maps.dy, &a_full[sa], a_ring + (size_t)sa * SD_SLOT, c * 64, w0, h0, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // MMA issuer: all lanes run the loop, the elected lane issues
    const bool leader = elect_one();
    mbar_wait(b_res, 0);
    tc_fence_after();
    uint32_t ia = 0, iu = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
      const uint32_t buf = iu & 1u;
      mbar_wait(&acc_empty[buf], ((iu >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * (uint32_t)(CG * SD_BN);
      for (int c = 0; c < p.kchunks; ++c, ++ia) {
        const uint32_t sa = ia % p.stages;
        mbar_wait(&a_full[sa], (ia / p.stages) & 1u);
        tc_fence_after();
        const uint64_t a_desc0 = sd_kdesc(smem_u32(a_ring + (size_t)sa * SD_SLOT), SD_SBO);
        const uint64_t b_desc0 = kmajor_desc(smem_u32(b_base + (size_t)c * NSLOT * SD_TAP_BYTES), SD_ROW);
        const uint32_t acc_on = c != 0 ? 1u : 0u;
        if (leader) {
          if constexpr (CG == 4) {
            sd_issue_class<0>(acc, a_desc0, b_desc0, sd_slot0(0), acc_on);
            sd_issue_class<1>(acc + SD_BN, a_desc0, b_desc0, sd_slot0(1), acc_on);
            sd_issue_class<2>(acc + 2 * SD_BN, a_desc0, b_desc0, sd_slot0(2), acc_on);
            sd_issue_class<3>(acc + 3 * SD_BN, a_desc0, b_desc0, sd_slot0(3), acc_on);
          } else if (grp == 0) {
            sd_issue_class<0>(acc, a_desc0, b_desc0, 0, acc_on);
            sd_issue_class<3>(acc + SD_BN, a_desc0, b_desc0, sd_ntaps(0), acc_on);
          } else {
            sd_issue_class<1>(acc, a_desc0, b_desc0, 0, acc_on);
            sd_issue_class<2>(acc + SD_BN, a_desc0, b_desc0, sd_ntaps(1), acc_on);
          }
          umma_commit(&a_empty[sa]);
        }
        __syncwarp();
      }
      if (leader) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int hl = r / SD_TW, wl = r % SD_TW;
    const int H = 2 * p.h, W = 2 * p.w;
    uint32_t iu = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
      int mt = u / p.nfix;
      const int w0 = (mt % p.tiles_w) * SD_TW;
      mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * SD_TH;
      const int b = mt / p.tiles_h;
      const uint32_t buf = iu & 1u;
      mbar_wait(&acc_full[buf], (iu >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int i = 0; i < CG; ++i) {
        const int cls = sd_class_of(CG, grp, i), ph = cls >> 1, pw = cls & 1;
        __nv_bfloat16* out =
            p.dst + (((size_t)b * H + (size_t)(2 * (h0 + hl) + ph)) * W + (size_t)(2 * (w0 + wl) + pw)) * p.N + n0;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)(CG * SD_BN) + (uint32_t)(i * SD_BN);
#pragma unroll 1
        for (int c = 0; c < SD_BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)c, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * e]), __uint_as_float(v[g * 8 + 2 * e + 1]));
              pk[e] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(out + c + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) sd_mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static constexpr size_t SD_SMEM_MAX = 227 * 1024;
static constexpr size_t SD_BAR_BYTES = 256;

// H, W = extent of dx (the conv's input); K = channels of dy (the conv's Cout), N = channels of dx (Cin)
bool conv_s2_dgrad_halo_supported(int H, int W, int N, int K) {
  if (getenv("BVAE_CONV_NO_HALO")) return false;
  if ((H | W) & 1) return false;
  return (H / 2) % SD_TH == 0 && (W / 2) % SD_TW == 0 && N % 64 == 0 && K % 64 == 0 && K <= 256;
}

int conv_s2_dgrad_halo(const void* dy, const void* wpt, void* dx, int B, int H, int W, int N, int K, cudaStream_t st) {
  SdMaps maps;
  SdParams p;
  memset(&p, 0, sizeof(p));
  const int h = H / 2, w = W / 2;
  const int cg = K <= 128 ? 4 : 2;
  const int nslot = cg == 4 ? 9 : 5;
  p.kchunks = K / 64;
  p.n_tiles = N / SD_BN;
  p.nfix = (4 / cg) * p.n_tiles;
  p.tiles_w = w / SD_TW;
  p.tiles_h = h / SD_TH;
  p.m_tiles = p.tiles_w * p.tiles_h * B;
  p.units = p.m_tiles * p.nfix;
  p.h = h; p.w = w; p.N = N;
  p.dst = reinterpret_cast<__nv_bfloat16*>(dx);
  const size_t wbytes = (size_t)nslot * p.kchunks * SD_TAP_BYTES;
  int sa = (int)((SD_SMEM_MAX - 1024 - SD_BAR_BYTES - wbytes) / SD_SLOT);
  BVAE_REQUIRE(sa >= 2, BVAE_ERR_UNSUPPORTED, "conv_s2_dgrad_halo: K=%d too large for resident weights", K);
  p.stages = sa > 4 ? 4 : sa;
  p.a_off = (uint32_t)wbytes;
  p.bar_off = p.a_off + (uint32_t)p.stages * SD_SLOT;
  const size_t smem = 1024 + p.bar_off + SD_BAR_BYTES;
  {
    const uint64_t dims[4] = {(uint64_t)K, (uint64_t)w, (uint64_t)h, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)K * 2, (uint64_t)w * K * 2, (uint64_t)h * w * K * 2};
    const uint32_t box[4] = {64u, SD_BW, SD_BH, 1u};
    int rc = make_map(&maps.dy, dy, 4, dims, strides, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, 9};
    const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)K * N * 2};
    const uint32_t box[3] = {64u, (uint32_t)SD_BN, 1u};
    int rc = make_map(&maps.w, wpt, 3, dims, strides, box, 128);
    if (rc) return rc;
  }
  const int sms = num_sms();
  int grid = p.units < sms ? p.units : sms;
  grid = grid / p.nfix * p.nfix;
  if (grid < p.nfix) grid = p.nfix;
  void (*kern)(const SdMaps, const SdParams) = cg == 4 ? conv_s2_dgrad_halo_kernel<4> : conv_s2_dgrad_halo_kernel<2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SD_SMEM_MAX);
  if (e != cudaSuccess) return cuda_fail(e, "conv_s2_dgrad_halo: cudaFuncSetAttribute");
  kern<<<grid, SD_THREADS, smem, st>>>(maps, p);
  BVAE_LAUNCH_CHECK("conv_s2_dgrad_halo");
  return BVAE_OK;
}

}  // namespace bvae
