// Stride-1 3x3 convolution (fprop and dgrad) on tcgen05: persistent implicit GEMM whose A operand
// is ONE shared-memory halo per (tile, 64-channel chunk), read nine times as shifted windows.
// (reference op: nn.Conv2d(3x3, stride 1, pad 1) at src/models/beta_vae.py:54 -- the decoder convs,
//  86 % of the model's FLOPs -- and convolution_backward's input gradient for it.)
//
// Why: the tap-per-TMA-load formulation (conv_tc.cu) re-reads every input pixel nine times from L2
// and re-fetches the weights for every tile.  At N = 64..128 output channels a 128-pixel tile then
// needs 110-190 B of shared-memory fill per tensor-core cycle while L2 delivers roughly 35-60 B
// per SM-cycle, so the tensor pipe idles.  Here
//   * a tile is 16 x 8 output pixels; its (18 x 10 pixel) x 64-channel input halo arrives with one
//     4-D TMA box (23 KB, out-of-image rows/columns zero-filled = the conv's padding) in the
//     128B-swizzled K-major layout.  Tap (di, dj) is the same buffer read through a descriptor whose
//     start address is shifted by ((di+1)*10 + (dj+1)) pixel rows and whose 8-row group stride is one
//     halo row (1280 B): the UMMA unit swizzles on absolute address bits, so windows that are not
//     aligned to the 1024-B swizzle atom are read correctly (dev/halo_desc_probe.cu is the probe
//     that established this on B200).  Input traffic drops 6.4x.
//   * CTAs are persistent (one per SM, static round-robin over work units) and keep the whole
//     [9][bn][K] weight slab resident in shared memory when it fits (K*bn <= 8192: the 64/128-channel
//     layers); otherwise weights stream through a second ring, shared by TM = 2 pixel tiles.
//   * accumulators are double-buffered in TMEM, so the epilogue of unit i (TMEM -> registers ->
//     +bias -> bf16 -> global) overlaps the MMAs of unit i+1.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace bvae {

static constexpr int HL_THREADS = 192;
static constexpr int HL_TH = 16, HL_TW = 8;                      // output tile
static constexpr int HL_HW = HL_TW + 2, HL_HH = HL_TH + 2;       // halo
static constexpr uint32_t HL_ROW = 128;                          // one pixel = 64 bf16 channels
static constexpr uint32_t HL_HALO_BYTES = HL_HH * HL_HW * HL_ROW;  // 23040
static constexpr uint32_t HL_HALO_SLOT = 24576;                  // padded to the 1024-B swizzle atom
static constexpr uint32_t HL_GROUP_STRIDE = HL_HW * HL_ROW;      // 1280: next tile row in the halo

struct HaloParams {
  int dgrad;               // 1: flipped taps (weight tap 8-j with halo window j)
  int kchunks;             // K / 64
  int bn, n_tiles;         // output channels per unit, N / bn
  int tm;                  // pixel tiles per unit (share every weight stage)
  int tiles_w, tiles_h, m_tiles;  // tiles per image row / column, total pixel tiles
  int m_groups, units;
  int B, H, W, N;
  int resident;            // weights resident in shared memory
  int sa, sb;              // ring depths
  int nbuf, tmem_cols;     // accumulator buffers in TMEM
  int dbg;                 // BVAE_HALO_DBG ablation bits (dev): 1 no global stores, 2 no epilogue work, 4 no MMAs, 8 no halo TMA
  uint32_t a_off, b_off, bar_off;
  uint32_t c_off;          // output staging: cst buffers of [128 px][64 ch] bf16 (SWIZZLE_128B), stored by TMA
  int cst;
  __nv_bfloat16* dst;
  const float* bias;
  double* stat_part;       // optional [B][tiles per sample * N/64 * 4][2]: per-warp (sum, sum of squares) of the stored output
};

struct HaloMaps {
  CUtensorMap x;
  CUtensorMap w;
  CUtensorMap y;           // dst [B][H][W][N], box 64 ch x 8 x 16 pixels
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major SW128 descriptor with an explicit 8-row-group stride (halo windows: 1280 B)
__device__ __forceinline__ uint64_t kmajor_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= 1ull << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

__device__ __forceinline__ float final_elem_loss(float r, float x, int type) {
  if (type == BVAE_LOSS_MSE) {
    const float d = r - x;
    return d * d;
  }
  if (type == BVAE_LOSS_L1) return fabsf(r - x);
  const float lp = fmaxf(logf(r), -100.f), l1p = fmaxf(logf(1.f - r), -100.f);  // F.binary_cross_entropy clamps at -100
  return -(x * lp + (1.f - x) * l1p);
}

// Window j = (r, s) of the halo starts ((r * 10) + s) pixel rows into the buffer.  fprop uses weight tap j with
// window j; dgrad (flipped kernel) uses weight tap 8 - j with window j -- the producer stores tap `wtap(j)` in
// weight slot j, so the MMA loop is identical for both and fully unrolled with constant offsets.
template <int BN, bool RESIDENT>
__global__ void __launch_bounds__(HL_THREADS, 1)
    conv_halo_tc_kernel(const __grid_constant__ HaloMaps maps, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem + p.a_off;
  uint8_t* b_base = smem + p.b_off;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* a_empty = a_full + p.sa;
  uint64_t* b_full = a_empty + p.sa;
  uint64_t* b_empty = b_full + p.sb;
  uint64_t* acc_full = b_empty + p.sb;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* b_res = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_res + 1);
  float* bias_sm = reinterpret_cast<float*>(smem + p.bar_off + 256);  // [N] (zeros when there is no bias)

  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  const uint32_t a_slot_bytes = (uint32_t)p.tm * HL_HALO_SLOT;
  constexpr uint32_t b_tap_bytes = (uint32_t)BN * HL_ROW;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.sa; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < p.sb; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    mbar_init(b_res, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.N; i += HL_THREADS) bias_sm[i] = p.bias ? p.bias[i] : 0.f;
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (all 32 lanes run the loop; the elected lane issues) =====
    const bool leader = elect_one();
    if (leader) {
      tma_prefetch_desc(&maps.x);
      tma_prefetch_desc(&maps.w);
      if (RESIDENT) {
        // every unit of this CTA has the same n-tile (gridDim.x is a multiple of n_tiles)
        const int n0 = (blockIdx.x % p.n_tiles) * BN;
        mbar_expect_tx(b_res, 9u * p.kchunks * b_tap_bytes);
        for (int c = 0; c < p.kchunks; ++c)
          for (int j = 0; j < 9; ++j)
            tma_load_3d(&maps.w, b_res, b_base + (size_t)(c * 9 + j) * b_tap_bytes, c * 64, n0, p.dgrad ? 8 - j : j);
      }
    }
    __syncwarp();
    uint32_t ia = 0, ib = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int nt = u % p.n_tiles, mg = u / p.n_tiles;
      const int n0 = nt * BN;
      for (int c = 0; c < p.kchunks; ++c, ++ia) {
        const uint32_t sa = ia % p.sa;
        mbar_wait(&a_empty[sa], ((ia / p.sa) & 1u) ^ 1u);
        int live = 0;
        for (int m = 0; m < p.tm; ++m) live += (mg * p.tm + m < p.m_tiles) ? 1 : 0;
        if (p.dbg & 8) live = 0;
        if (leader) {
          if (p.dbg & 8) mbar_arrive(&a_full[sa]);
          else mbar_expect_tx(&a_full[sa], (uint32_t)live * HL_HALO_BYTES);
        }
        for (int m = 0; m < live; ++m) {
          int mt = mg * p.tm + m;
          const int w0 = (mt % p.tiles_w) * HL_TW;
          mt /= p.tiles_w;
          const int h0 = (mt % p.tiles_h) * HL_TH;
          const int b = mt / p.tiles_h;
          if (leader)
            tma_load_4d(&maps.x, &a_full[sa], a_ring + (size_t)sa * a_slot_bytes + (size_t)m * HL_HALO_SLOT, c * 64, w0 - 1,
                        h0 - 1, b);
        }
        if (!RESIDENT) {
#pragma unroll 1
          for (int j = 0; j < 9; ++j, ++ib) {
            const uint32_t sb = ib % p.sb;
            mbar_wait(&b_empty[sb], ((ib / p.sb) & 1u) ^ 1u);
            if (leader) {
              mbar_expect_tx(&b_full[sb], b_tap_bytes);
              tma_load_3d(&maps.w, &b_full[sb], b_base + (size_t)sb * b_tap_bytes, c * 64, n0, p.dgrad ? 8 - j : j);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (all 32 lanes run the loop, the elected lane issues; the tap / k loops are unrolled with
    //       constant descriptor offsets, so consecutive tcgen05.mma are one or two uniform adds apart) =====
    const bool leader = elect_one();
    constexpr uint32_t idesc =
        (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (RESIDENT) {
      mbar_wait(b_res, 0);
      tc_fence_after();
    }
    uint32_t ia = 0, ib = 0, iu = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
      const int mg = u / p.n_tiles;
      const uint32_t buf = iu % p.nbuf;
      mbar_wait(&acc_empty[buf], ((iu / p.nbuf) & 1u) ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * (uint32_t)(p.tm * BN);
      const bool two = (mg * p.tm + 1 < p.m_tiles && p.tm == 2);
      for (int c = 0; c < p.kchunks; ++c, ++ia) {
        const uint32_t sa = ia % p.sa;
        mbar_wait(&a_full[sa], (ia / p.sa) & 1u);
        tc_fence_after();
        const uint64_t a_desc0 = kmajor_desc_sbo(smem_u32(a_ring + (size_t)sa * a_slot_bytes), HL_GROUP_STRIDE);
        const uint32_t acc_on = c != 0 ? 1u : 0u;
        if (RESIDENT) {
          const uint64_t b_desc_res = kmajor_desc(smem_u32(b_base + (size_t)c * 9 * b_tap_bytes), HL_ROW);
          if (leader && !(p.dbg & 4)) {
#pragma unroll
            for (int j = 0; j < 9; ++j) {
              const uint64_t b_desc = b_desc_res + (uint64_t)((j * b_tap_bytes) >> 4);
              const uint32_t win = (uint32_t)((j / 3) * HL_HW + (j % 3)) * HL_ROW;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(acc, a_desc0 + (uint64_t)((win + k * 32) >> 4), b_desc + (uint64_t)(k * 2), idesc,
                          (j | k) != 0 ? 1u : acc_on);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 9; ++j, ++ib) {
            const uint32_t sb = ib % p.sb;
            mbar_wait(&b_full[sb], (ib / p.sb) & 1u);
            tc_fence_after();
            const uint64_t b_desc = kmajor_desc(smem_u32(b_base + (size_t)sb * b_tap_bytes), HL_ROW);
            const uint32_t win = (uint32_t)((j / 3) * HL_HW + (j % 3)) * HL_ROW;
            if (leader) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(acc, a_desc0 + (uint64_t)((win + k * 32) >> 4), b_desc + (uint64_t)(k * 2), idesc,
                          (j | k) != 0 ? 1u : acc_on);
              if (two) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(acc + BN, a_desc0 + (uint64_t)((HL_HALO_SLOT + win + k * 32) >> 4), b_desc + (uint64_t)(k * 2),
                            idesc, (j | k) != 0 ? 1u : acc_on);
              }
              umma_commit(&b_empty[sb]);
            }
            __syncwarp();
          }
        }
        if (leader) umma_commit(&a_empty[sa]);
        __syncwarp();
      }
      if (leader) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    // TMEM -> registers -> (+bias, bf16) -> 128B-swizzled staging tile in shared memory -> ONE TMA store per 64-channel
    // block.  (Storing straight from registers -- 16 bytes per lane, every lane in another 128-byte line -- costs 32
    // LSU wavefronts per instruction: ~2000 cycles per tile, more than the tile's MMAs.)
    const int q = warp & 3;
    const int r = q * 32 + lane;            // pixel inside the 16x8 tile = row of the staging tile
    const bool store_leader = (warp == 2) && elect_one();
    uint8_t* stage0 = smem + p.c_off;
    uint32_t iu = 0, is = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
      const int nt = u % p.n_tiles, mg = u / p.n_tiles;
      const int n0 = nt * BN;
      const uint32_t buf = iu % p.nbuf;
      mbar_wait(&acc_full[buf], (iu / p.nbuf) & 1u);
      tc_fence_after();
      for (int m = 0; m < p.tm; ++m) {
        int mt = mg * p.tm + m;
        if (mt >= p.m_tiles) break;
        const int w0 = (mt % p.tiles_w) * HL_TW;
        mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * HL_TH;
        const int b = mt / p.tiles_h;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)(p.tm * BN) + (uint32_t)(m * BN);
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 64, ++is) {
          if (p.dbg & 2) break;
          uint8_t* stg = stage0 + (size_t)(is % (uint32_t)p.cst) * 16384u;
          // the TMA store that last read this staging buffer must be done with it
          if (store_leader) {
            if (p.cst == 2) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
          }
          named_bar_sync(1, 128);
          uint8_t* row = stg + (size_t)r * 128u;
          float st_s = 0.f, st_q = 0.f;     // GroupNorm statistics of what is stored (bf16-rounded), for the next pass
#pragma unroll
          for (int c = 0; c < 64; c += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)(cb + c), v);
            tmem_ld_wait();
            const float4* bq = reinterpret_cast<const float4*>(bias_sm + n0 + cb + c);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 b0 = bq[2 * g], b1 = bq[2 * g + 1];
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float f0 = __uint_as_float(v[g * 8 + 2 * e]) + bb[2 * e];
                const float f1 = __uint_as_float(v[g * 8 + 2 * e + 1]) + bb[2 * e + 1];
                __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
                pk[e] = *reinterpret_cast<uint32_t*>(&h2);
                const float r0 = __uint_as_float(pk[e] << 16), r1 = __uint_as_float(pk[e] & 0xffff0000u);
                st_s += r0 + r1;
                st_q = fmaf(r0, r0, fmaf(r1, r1, st_q));
              }
              const uint32_t chunk = (uint32_t)(c / 8 + g);     // 16-byte chunk of the 128-byte row, XOR-swizzled by the row
              *reinterpret_cast<uint4*>(row + ((chunk ^ (uint32_t)(r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
          if (p.stat_part) {
            st_s = warp_sum(st_s);
            st_q = warp_sum(st_q);
            if (lane == 0) {
              const int tps = p.tiles_w * p.tiles_h;                 // tiles per sample
              const int ts = (mg * p.tm + m) - b * tps;
              const size_t slot = ((size_t)b * tps + ts) * (size_t)(p.N / 64) + (size_t)((n0 + cb) / 64);
              p.stat_part[(slot * 4 + q) * 2] = (double)st_s;
              p.stat_part[(slot * 4 + q) * 2 + 1] = (double)st_q;
            }
          }
          fence_proxy_async();
          named_bar_sync(1, 128);
          if (store_leader && !(p.dbg & 1)) {
            tma_store_4d(&maps.y, stg, n0 + cb, w0, h0, b);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (store_leader) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static constexpr size_t HL_SMEM_MAX = 227 * 1024;

bool conv_halo_supported(int H, int W, int K, int N) {
  if (getenv("BVAE_CONV_NO_HALO")) return false;
  return H % HL_TH == 0 && W % HL_TW == 0 && K % 64 == 0 && N % 64 == 0 && (N <= 256 ? (N & (N - 1)) == 0 : N % 256 == 0);
}

int conv_halo_stat_parts(int H, int W, int N) { return (H / HL_TH) * (W / HL_TW) * (N / 64) * 4; }

int conv_halo_launch(const void* src, const void* wp, const float* bias, void* dst, int B, int H, int W, int K, int N,
                     int dgrad, cudaStream_t st, double* stat_part) {
  HaloMaps maps;
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.stat_part = stat_part;
  p.dgrad = dgrad ? 1 : 0;
  if (const char* e = getenv("BVAE_HALO_DBG")) p.dbg = atoi(e);
  p.kchunks = K / 64;
  p.tiles_w = W / HL_TW;
  p.tiles_h = H / HL_TH;
  p.m_tiles = p.tiles_w * p.tiles_h * B;
  p.B = B; p.H = H; p.W = W; p.N = N;
  p.dst = reinterpret_cast<__nv_bfloat16*>(dst);
  p.bias = bias;
  const int sms = num_sms();
  const size_t bar_bytes = 256 + (size_t)N * sizeof(float) + 64;  // mbarriers + bias[N]
  // resident weights: the whole [9][bn][K] slab plus at least two halo slots and one output staging tile must fit
  constexpr size_t STG = 16384;   // [128 px][64 ch] bf16
  int bn_res = 0;
  for (int bn = (N < 256 ? N : 256); bn >= 64; bn >>= 1) {
    if (N % bn) continue;
    if ((size_t)9 * K * bn * 2 + 2 * HL_HALO_SLOT + STG + bar_bytes + 1024 <= HL_SMEM_MAX) {
      bn_res = bn;
      break;
    }
  }
  size_t smem;
  if (bn_res) {
    p.resident = 1;
    p.bn = bn_res;
    p.tm = 1;
    p.sb = 0;
    const size_t wbytes = (size_t)9 * K * p.bn * 2;
    // halo slots first: the kernel is bound by bytes in flight per SM (Little's law: ~2.4 us load latency under load),
    // so every spare 24 KB goes to the input ring; the output staging is double-buffered only if room is left
    const size_t room = HL_SMEM_MAX - 1024 - bar_bytes - wbytes;
    int sa = (int)((room - STG) / HL_HALO_SLOT);
    p.sa = sa > 6 ? 6 : sa;
    if (const char* e = getenv("BVAE_HALO_SA")) p.sa = atoi(e) < p.sa ? atoi(e) : p.sa;
    p.cst = (room - (size_t)p.sa * HL_HALO_SLOT) >= 2 * STG ? 2 : 1;
    p.b_off = 0;
    p.a_off = (uint32_t)wbytes;  // multiple of 1024 (bn*128*9*kchunks)
    p.c_off = p.a_off + (uint32_t)p.sa * HL_HALO_SLOT;
    p.bar_off = p.c_off + (uint32_t)(p.cst * STG);
    smem = 1024 + p.bar_off + bar_bytes;
  } else {
    p.resident = 0;
    p.bn = N >= 128 ? 128 : 64;
    const int units1 = p.m_tiles * (N / p.bn);
    p.tm = (units1 / 2 >= (sms * 4) / 5) ? 2 : 1;
    p.sa = p.tm == 2 ? 2 : 3;
    p.cst = 2;
    const size_t a_bytes = (size_t)p.sa * p.tm * HL_HALO_SLOT;
    int sb = (int)((HL_SMEM_MAX - 1024 - bar_bytes - a_bytes - p.cst * STG) / ((size_t)p.bn * HL_ROW));
    p.sb = sb > 8 ? 8 : sb;
    p.a_off = 0;
    p.b_off = (uint32_t)a_bytes;
    p.c_off = p.b_off + (uint32_t)p.sb * p.bn * HL_ROW;
    p.bar_off = p.c_off + (uint32_t)(p.cst * STG);
    smem = 1024 + p.bar_off + bar_bytes;
  }
  p.n_tiles = N / p.bn;
  p.m_groups = (p.m_tiles + p.tm - 1) / p.tm;
  p.units = p.m_groups * p.n_tiles;
  p.nbuf = (2 * p.tm * p.bn <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < p.nbuf * p.tm * p.bn) cols <<= 1;
  p.tmem_cols = cols;
  BVAE_REQUIRE(smem <= HL_SMEM_MAX, BVAE_ERR_UNSUPPORTED, "conv_halo: shared memory plan %zu too large", smem);

  {
    const uint64_t dims[4] = {(uint64_t)K, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)K * 2, (uint64_t)W * K * 2, (uint64_t)H * W * K * 2};
    const uint32_t box[4] = {64u, (uint32_t)HL_HW, (uint32_t)HL_HH, 1u};
    int rc = make_map(&maps.x, src, 4, dims, strides, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, 9};
    const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)K * N * 2};
    const uint32_t box[3] = {64u, (uint32_t)p.bn, 1u};
    int rc = make_map(&maps.w, wp, 3, dims, strides, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)N, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)N * 2, (uint64_t)W * N * 2, (uint64_t)H * W * N * 2};
    const uint32_t box[4] = {64u, (uint32_t)HL_TW, (uint32_t)HL_TH, 1u};
    int rc = make_map(&maps.y, dst, 4, dims, strides, box, 128);
    if (rc) return rc;
  }
  int grid = p.units < sms ? p.units : sms;
  if (p.resident) {
    // every CTA must keep one n-tile for all its units: grid a multiple of n_tiles
    grid = grid / p.n_tiles * p.n_tiles;
    if (grid < p.n_tiles) grid = p.n_tiles;
  }
  void (*kern)(const HaloMaps, const HaloParams) = nullptr;
  if (p.resident) kern = p.bn == 64 ? conv_halo_tc_kernel<64, true> : (p.bn == 128 ? conv_halo_tc_kernel<128, true> : conv_halo_tc_kernel<256, true>);
  else            kern = p.bn == 64 ? conv_halo_tc_kernel<64, false> : (p.bn == 128 ? conv_halo_tc_kernel<128, false> : conv_halo_tc_kernel<256, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_MAX);
  if (e != cudaSuccess) return cuda_fail(e, "conv_halo: cudaFuncSetAttribute");
  kern<<<grid, HL_THREADS, smem, st>>>(maps, p);
  BVAE_LAUNCH_CHECK("conv_halo_tc");
  return BVAE_OK;
}

// ------------------------------------------------------------------------------------------
// final_conv (C -> Co <= 3) + sigmoid + reconstruction loss  (src/models/beta_vae.py:163-164,185-192)
//
// With one to three output channels the tap-by-tap implicit GEMM would issue nine N = 16 MMAs per K step whose cost is
// all A-operand fetch.  Here the channel contraction is done ONCE per halo pixel for all nine taps,
//     G[hp][t * Co + co] = sum_c halo[hp][c] * w[co][c][t]          (N = 9 * Co padded to 16 / 32),
// as two M = 128 MMAs per K step over the 180 halo rows read as a dense K-major tile (rows 0..127 and 52..179; the
// second start is not swizzle-atom aligned, which the UMMA address swizzle handles -- see the header of this file), and
// the 3x3 spatial sum is a nine-term stencil over G in the epilogue:
//     logit[p][co] = bias[co] + sum_t G[p + off_t][t * Co + co].
// 4.5x fewer MMAs than tap-by-tap; the kernel is bound by the single read of h.
// Epilogue warps: TMEM -> G in shared memory (double buffered) -> named barrier -> stencil + sigmoid + loss.
// ------------------------------------------------------------------------------------------
static constexpr int FG_HALO_PX = HL_HH * HL_HW;   // 180
static constexpr int FG_ROW2 = FG_HALO_PX - 128;   // 52: first halo row of the second MMA

struct FinalGsParams {
  int kchunks, tiles_w, tiles_h, units;
  int B, H, W, Co, nt;       // nt = 9 * Co
  int sa;
  uint32_t a_off, g_off, bar_off;
  const float* bias;
  const float* target;       // [B][Co][H][W] fp32 or null
  float* recon;              // [B][Co][H][W] fp32
  double* lpart;             // [gridDim.x * 4] or null
  int loss_type;
};

// 320 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 and 6..9 two epilogue groups that take alternate
// tiles (group = accumulator / G buffer): the stencil + sigmoid + loss epilogue of one tile (~1500 cycles) is longer
// than its eight MMAs, so with one group the kernel was epilogue-bound (28 us for a 67 MB read).
static constexpr int FG_THREADS = 320;

template <int N>
__global__ void __launch_bounds__(FG_THREADS, 1)
    finalconv_gs_kernel(const __grid_constant__ HaloMaps maps, const FinalGsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_base = smem;                                  // [kchunks][N][64] bf16, 128B-swizzled rows
  uint8_t* a_ring = smem + p.a_off;
  float* g_sm = reinterpret_cast<float*>(smem + p.g_off);  // [2][180][nt]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* a_empty = a_full + p.sa;
  uint64_t* acc_full = a_empty + p.sa;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* b_res = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_res + 1);
  constexpr uint32_t TMEM_COLS = 4 * N;                    // two accumulators per buffer, two buffers
  constexpr uint32_t B_CHUNK = (uint32_t)N * HL_ROW;
  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.sa; ++s) {
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
      tma_prefetch_desc(&maps.x);
      tma_prefetch_desc(&maps.w);
      mbar_expect_tx(b_res, (uint32_t)p.kchunks * B_CHUNK);
      for (int c = 0; c < p.kchunks; ++c) tma_load_3d(&maps.w, b_res, b_base + (size_t)c * B_CHUNK, c * 64, 0, 0);
    }
    __syncwarp();
    uint32_t ia = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      int mt = u;
      const int w0 = (mt % p.tiles_w) * HL_TW;
      mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * HL_TH;
      const int b = mt / p.tiles_h;
      for (int c = 0; c < p.kchunks; ++c, ++ia) {
        const uint32_t sa = ia % p.sa;
        mbar_wait(&a_empty[sa], ((ia / p.sa) & 1u) ^ 1u);
        if (leader) {
          mbar_expect_tx(&a_full[sa], HL_HALO_BYTES);
          tma_load_4d(&maps.x, &a_full[sa], a_ring + (size_t)sa * HL_HALO_SLOT, c * 64, w0 - 1, h0 - 1, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc =
        (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    mbar_wait(b_res, 0);
    tc_fence_after();
    uint32_t ia = 0, iu = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
      const uint32_t buf = iu & 1u;
      mbar_wait(&acc_empty[buf], ((iu >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * (uint32_t)(2 * N);
      for (int c = 0; c < p.kchunks; ++c, ++ia) {
        const uint32_t sa = ia % p.sa;
        mbar_wait(&a_full[sa], (ia / p.sa) & 1u);
        tc_fence_after();
        const uint64_t a_desc0 = kmajor_desc(smem_u32(a_ring + (size_t)sa * HL_HALO_SLOT), HL_ROW);   // dense rows
        const uint64_t b_desc0 = kmajor_desc(smem_u32(b_base + (size_t)c * B_CHUNK), HL_ROW);
        const uint32_t acc_on = c != 0 ? 1u : 0u;
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(acc, a_desc0 + (uint64_t)((k * 32) >> 4), b_desc0 + (uint64_t)(k * 2), idesc, k != 0 ? 1u : acc_on);
            umma_bf16(acc + N, a_desc0 + (uint64_t)((FG_ROW2 * HL_ROW + k * 32) >> 4), b_desc0 + (uint64_t)(k * 2), idesc,
                      k != 0 ? 1u : acc_on);
          }
          umma_commit(&a_empty[sa]);
        }
        __syncwarp();
      }
      if (leader) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int group = (warp - 2) >> 2;      // epilogue group 0 / 1 = accumulator buffer = G buffer
    const int r = q * 32 + lane;            // TMEM lane = halo row (first MMA) / halo row - 52 (second); also the output pixel
    const int hl = r / HL_TW, wl = r % HL_TW;
    const int nt = p.nt;
    uint32_t iu = 0;
    float lsum = 0.f;
    float bias_r[3];
#pragma unroll
    for (int co = 0; co < 3; ++co) bias_r[co] = (p.bias && co < p.Co) ? p.bias[co] : 0.f;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
      if ((int)(iu & 1u) != group) continue;
      int mt = u;
      const int w0 = (mt % p.tiles_w) * HL_TW;
      mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * HL_TH;
      const int b = mt / p.tiles_h;
      const uint32_t buf = iu & 1u;
      float* g = g_sm + (size_t)buf * FG_HALO_PX * nt;
      // the target pixel is fetched before waiting on the accumulator so its DRAM latency is off the per-tile chain
      const size_t o0 = ((size_t)b * p.Co * p.H + (size_t)(h0 + hl)) * p.W + (size_t)(w0 + wl);
      const size_t plane = (size_t)p.H * p.W;
      float tgt[3] = {0.f, 0.f, 0.f};
      if (p.target) {
#pragma unroll
        for (int co = 0; co < 3; ++co)
          if (co < p.Co) tgt[co] = __ldg(p.target + o0 + (size_t)co * plane);
      }
      mbar_wait(&acc_full[buf], (iu >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)(2 * N);
      {
        uint32_t v[N], v2[N];
        if constexpr (N == 16) {
          tmem_ld16(taddr, v);
          tmem_ld16(taddr + N, v2);
        } else {
          tmem_ld32(taddr, v);
          tmem_ld32(taddr + N, v2);
        }
        tmem_ld_wait();
        tc_fence_before();
        float* g1 = g + (size_t)r * nt;
#pragma unroll
        for (int n = 0; n < N; ++n)
          if (n < nt) g1[n] = __uint_as_float(v[n]);
        if (r >= 128 - FG_ROW2) {            // halo rows 128..179 come from the second accumulator
          float* g2 = g + (size_t)(r + FG_ROW2) * nt;
#pragma unroll
          for (int n = 0; n < N; ++n)
            if (n < nt) g2[n] = __uint_as_float(v2[n]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);          // accumulators are in registers / shared memory now
      named_bar_sync(1 + group, 128);                       // G of this tile complete (this epilogue group only)
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        if (co >= p.Co) break;
        float logit = bias_r[co];
#pragma unroll
        for (int t = 0; t < 9; ++t) logit += g[(size_t)((hl + t / 3) * HL_HW + wl + t % 3) * nt + t * p.Co + co];
        const float rr = 1.f / (1.f + expf(-logit));
        p.recon[o0 + (size_t)co * plane] = rr;
        if (p.target) lsum += final_elem_loss(rr, tgt[co], p.loss_type);
      }
      named_bar_sync(1 + group, 128);                       // everyone has read G before the group's next tile overwrites it
    }
    if (p.lpart) {
      const double s = warp_sum((double)lsum);
      if (lane == 0) p.lpart[blockIdx.x * 8 + (warp - 2)] = s;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

bool conv_halo_final_supported(int H, int W, int C, int Co) {
  if (getenv("BVAE_CONV_NO_HALO")) return false;
  return H % HL_TH == 0 && W % HL_TW == 0 && C % 64 == 0 && C <= 256 && Co >= 1 && Co <= 3;
}

int conv_halo_final_rows(int Co) { return Co == 1 ? 16 : 32; }

// wg = [rows][C] bf16 with row t * Co + co = w[co][:][t] (rows >= 9 * Co zero), rows = conv_halo_final_rows(Co)
int conv_halo_final_fwd(const void* h, const void* wg, const float* bias, const float* target, float* recon,
                        double* lpart, int* n_lpart, int loss_type, int B, int H, int W, int C, int Co,
                        cudaStream_t st) {
  HaloMaps maps;
  FinalGsParams p;
  memset(&p, 0, sizeof(p));
  const int N = conv_halo_final_rows(Co);
  p.kchunks = C / 64;
  p.tiles_w = W / HL_TW;
  p.tiles_h = H / HL_TH;
  p.units = p.tiles_w * p.tiles_h * B;
  p.B = B; p.H = H; p.W = W; p.Co = Co; p.nt = 9 * Co;
  p.bias = bias;
  p.target = target; p.recon = recon; p.lpart = lpart; p.loss_type = loss_type;
  const size_t wbytes = align_up((size_t)p.kchunks * N * HL_ROW, 1024);
  const size_t gbytes = align_up((size_t)2 * FG_HALO_PX * p.nt * sizeof(float), 1024);
  const size_t bar_bytes = 512;
  int sa = (int)((HL_SMEM_MAX - 1024 - bar_bytes - wbytes - gbytes) / HL_HALO_SLOT);
  p.sa = sa > 6 ? 6 : sa;
  p.a_off = (uint32_t)wbytes;
  p.g_off = p.a_off + (uint32_t)p.sa * HL_HALO_SLOT;
  p.bar_off = p.g_off + (uint32_t)gbytes;
  const size_t smem = 1024 + p.bar_off + bar_bytes;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {64u, (uint32_t)HL_HW, (uint32_t)HL_HH, 1u};
    int rc = make_map(&maps.x, h, 4, dims, strides, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)N, 1};
    const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)C * N * 2};
    const uint32_t box[3] = {64u, (uint32_t)N, 1u};
    int rc = make_map(&maps.w, wg, 3, dims, strides, box, 128);
    if (rc) return rc;
  }
  const int sms = num_sms();
  const int grid = p.units < sms ? p.units : sms;
  auto kern = N == 16 ? finalconv_gs_kernel<16> : finalconv_gs_kernel<32>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_MAX);
  if (e != cudaSuccess) return cuda_fail(e, "conv_halo_final: cudaFuncSetAttribute");
  kern<<<grid, FG_THREADS, smem, st>>>(maps, p);
  BVAE_LAUNCH_CHECK("finalconv_gs");
  *n_lpart = grid * 8;
  return BVAE_OK;
}

}  // namespace bvae
