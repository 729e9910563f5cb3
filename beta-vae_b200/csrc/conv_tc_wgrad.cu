// Weight gradient of the 3x3 convolutions on the 5th-generation tensor cores.
//   dW[t][ci][co] = sum over output pixels p of  x[p + off_t][ci] * dy[p][co]
// (what convolution_backward computes for nn.Conv2d at src/models/beta_vae.py:37,42,54).
//
// GEMM view: the reduction (K) runs over PIXELS, so both operands are "MN-major": a TMA box of
// NHWC data -- 64 pixels x 64 channels, one 128-byte row per pixel -- is exactly the 128B-swizzled
// MN-major layout tcgen05.mma reads.  No transposes, no im2col.
//   A (M = 128 rows)  : two 64-channel x-boxes stacked along M.  A "unit" is (tap, 64-channel chunk
//                       of Cin); a 128-row block holds units 2*rb and 2*rb+1, so with Cin = 64 one
//                       MMA covers two taps, with Cin >= 128 two channel chunks of one tap.
//   B (N = BN <= 256) : the dy box(es) of the same 64 pixels, BN/64 boxes side by side.
//   D                 : up to 512/BN accumulator slots live in TMEM at once, each [128 x BN] fp32;
//                       one dy tile is reused by all slots of the CTA.
// Stride-2 convolutions read x through the four parity tensor maps (see conv_tc.cu).
// Split-K: every CTA owns a contiguous range of 64-pixel tiles; partial results go to the fp32
// workspace [split][tap][ci][co] that the (deterministic) reduce kernel of conv_simt.cu sums.
// Warp roles as in conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue.
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace bvae {

static constexpr int WG_THREADS = 192;
static constexpr int WG_PIX = 64;                    // pixels per pipeline stage
static constexpr uint32_t WG_BOX = WG_PIX * 64 * 2;  // bytes of one 64-pixel x 64-channel box (8 KB)

struct WgParams {
  TcTap taps[9];
  int cin_chunks, units;  // Cin/64, 9*Cin/64
  int slots;              // accumulator slots per CTA
  int bn;                 // co tile
  int tw, th, tn, tiles_w, tiles_h, tiles_total, tiles_per_split;
  int B, Cin, Cout;
  int stages, tmem_cols;
  float* part;
};

struct WgMaps {
  CUtensorMap x[4];
  CUtensorMap dy;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
    conv_wgrad_tc_kernel(const __grid_constant__ WgMaps maps, const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nb = p.bn / 64;                                      // dy boxes per stage
  const uint32_t stage_bytes = (uint32_t)(nb + 2 * p.slots) * WG_BOX;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* accum_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  const int rb0 = blockIdx.x * p.slots;  // first 128-row block of this CTA
  const int n0 = blockIdx.y * p.bn;
  const int split = blockIdx.z;
  const int t_beg = split * p.tiles_per_split;
  const int t_end = min(p.tiles_total, t_beg + p.tiles_per_split);
  const int iters = t_end - t_beg;
  // valid units of this CTA (the last row block may be half empty, the last group short)
  int nslots = 0;
  for (int s = 0; s < p.slots; ++s)
    if (2 * (rb0 + s) < p.units) nslots = s + 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // producer: all lanes run the loop, the elected lane issues (see elect_one() in tc_ptx.cuh)
    const bool leader = elect_one();
    int nunits = 0;
    for (int s = 0; s < nslots; ++s)
      for (int h = 0; h < 2; ++h)
        if (2 * (rb0 + s) + h < p.units) ++nunits;
    const uint32_t tx_bytes = (uint32_t)(nb + nunits) * WG_BOX;
    for (int it = 0; it < iters; ++it) {
      const int st = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(&empty_bar[st], ph ^ 1u);
      int tile = t_beg + it;
      const int w0 = (tile % p.tiles_w) * p.tw;
      tile /= p.tiles_w;
      const int h0 = (tile % p.tiles_h) * p.th;
      const int b0 = (tile / p.tiles_h) * p.tn;
      uint8_t* base = smem + (size_t)st * stage_bytes;
      uint8_t* xb = base + (size_t)nb * WG_BOX;
      if (leader) {
        mbar_expect_tx(&full_bar[st], tx_bytes);
        for (int j = 0; j < nb; ++j) tma_load_4d(&maps.dy, &full_bar[st], base + (size_t)j * WG_BOX, n0 + 64 * j, w0, h0, b0);
        for (int s = 0; s < nslots; ++s)
          for (int h = 0; h < 2; ++h) {
            const int u = 2 * (rb0 + s) + h;
            if (u >= p.units) continue;
            const TcTap tap = p.taps[u / p.cin_chunks];
            tma_load_4d(&maps.x[tap.map], &full_bar[st], xb + (size_t)(2 * s + h) * WG_BOX, (u % p.cin_chunks) * 64,
                        w0 + tap.dj, h0 + tap.di, b0);
          }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    // D fp32, A/B bf16, both MN-major (bits 15, 16), N = bn, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int it = 0; it < iters; ++it) {
      const int st = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(&full_bar[st], ph);
      tc_fence_after();
      const uint32_t dy_addr = smem_u32(smem + (size_t)st * stage_bytes);
      const uint32_t x_addr = dy_addr + (uint32_t)nb * WG_BOX;
      if (leader) {
        for (int s = 0; s < nslots; ++s) {
#pragma unroll
          for (int k = 0; k < WG_PIX / 16; ++k) {
            // 16 pixels = 16 rows of 128 bytes = two swizzle atoms: advance 2048 bytes per K step
            const uint64_t adesc = mnmajor_desc(x_addr + (uint32_t)(2 * s) * WG_BOX + k * 2048u, WG_BOX, 1024u);
            const uint64_t bdesc = mnmajor_desc(dy_addr + k * 2048u, WG_BOX, 1024u);
            umma_bf16(tmem_base + (uint32_t)(s * p.bn), adesc, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[st]);
      }
      __syncwarp();
    }
    if (leader && iters > 0) umma_commit(accum_bar);
    __syncwarp();
  } else if (iters > 0) {
    const int q = warp & 3;
    const int m = q * 32 + lane;  // accumulator row: unit half (m / 64), channel within chunk (m % 64)
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    for (int s = 0; s < nslots; ++s) {
      const int u = 2 * (rb0 + s) + (m >> 6);
      const bool valid = u < p.units;
      const int tap = valid ? p.taps[u / p.cin_chunks].widx : 0;
      const int ci = valid ? (u % p.cin_chunks) * 64 + (m & 63) : 0;
      float* out = p.part + (((size_t)split * 9 + tap) * p.Cin + ci) * p.Cout + n0;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * p.bn);
      for (int c = 0; c < p.bn; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(out + c + g * 4) = make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------
// wgrad, halo formulation (same idea as conv_tc_halo.cu).  Per (th x 8)-pixel output tile the input arrives
// as ONE TMA box per source plane and the nine taps are shifted MN-major windows of those buffers
// (start = window offset, 8-pixel atoms one buffer row apart):
//   stride 1: one plane, box (th+2) x 10 pixels x 64 channels, rows 1280 B apart;
//   stride 2: the four parity planes x[2i+ph][2j+pw] (plain strided tensor maps), each a (th+1) x 9 box with rows
//             1152 B apart; tap (r, s) reads plane ((r != 1), (s != 1)) at row / column offset (r != 0), (s != 0).
// A CTA owns one (64 input channels) x (64 output channels) block of dW for ALL nine taps -- five 128-row
// accumulators in TMEM, two taps each, paired in increasing window address so the second tap is reached through the
// descriptor's leading-dimension offset -- and walks its share of the pixel tiles (split-K over tiles).  The spare
// half of the fifth accumulator multiplies an all-ones A block with the dy tile, i.e. it is the bias gradient: no
// separate column-sum pass over dy.  Per stride-1 tile: 23 KB of x + 16 KB of dy for 40 MMAs (128x64x16).
// ------------------------------------------------------------------------------------------
static constexpr uint32_t WH_ONES_BYTES = 2560;

struct WgHaloParams {
  int cin_chunks, co_blocks;
  int tiles_w, tiles_h, m_tiles, tiles_per_split;
  int Cin, Cout;
  int stages, nplanes, th, ksteps;
  uint32_t slot_bytes;      // one plane buffer (padded to 1024)
  uint32_t box_bytes;       // bytes TMA delivers per plane
  uint32_t dy_bytes;        // th * 8 * 128
  uint32_t stage_bytes;     // nplanes * slot_bytes + dy_bytes
  uint32_t sbo;             // bytes between 8-pixel atoms of a window = one buffer row
  uint32_t win[10];         // byte offset of accumulator-slot position q's window inside a stage (q = 2*rb + half)
  uint32_t ones_off, bar_off;
  float* part;   // [split][block][rb 5][column group 16][row 128][4]: the accumulators as they sit in TMEM
};
struct WgHaloMaps {
  CUtensorMap x[4], dy;
};

template <int KSTEPS>
__global__ void __launch_bounds__(WG_THREADS, 1)
    conv_wgrad_halo_kernel(const __grid_constant__ WgHaloMaps maps, const WgHaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* accum_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  const int cc = blockIdx.x % p.cin_chunks, cb = blockIdx.x / p.cin_chunks;
  const int split = blockIdx.y;
  const int t_beg = split * p.tiles_per_split;
  const int t_end = min(p.m_tiles, t_beg + p.tiles_per_split);
  const int iters = t_end - t_beg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  for (uint32_t i = threadIdx.x; i < WH_ONES_BYTES / 4; i += WG_THREADS)
    reinterpret_cast<uint32_t*>(smem + p.ones_off)[i] = 0x3F803F80u;  // bf16 1.0 pairs
  fence_proxy_async();
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // producer: all lanes run the loop, the elected lane issues the TMA loads
    const bool leader = elect_one();
    if (leader) {
      for (int q = 0; q < p.nplanes; ++q) tma_prefetch_desc(&maps.x[q]);
      tma_prefetch_desc(&maps.dy);
    }
    int st = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty_bar[st], ph ^ 1u);
      int tile = t_beg + it;
      const int w0 = (tile % p.tiles_w) * 8;
      tile /= p.tiles_w;
      const int h0 = (tile % p.tiles_h) * p.th;
      const int b = tile / p.tiles_h;
      uint8_t* base = smem + (size_t)st * p.stage_bytes;
      if (leader) {
        mbar_expect_tx(&full_bar[st], (uint32_t)p.nplanes * p.box_bytes + p.dy_bytes);
        for (int q = 0; q < p.nplanes; ++q)
          tma_load_4d(&maps.x[q], &full_bar[st], base + (size_t)q * p.slot_bytes, cc * 64, w0 - 1, h0 - 1, b);
        tma_load_4d(&maps.dy, &full_bar[st], base + (size_t)p.nplanes * p.slot_bytes, cb * 64, w0, h0, b);
      }
      __syncwarp();
      if (++st == p.stages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    // D fp32, A/B bf16, both MN-major, N = 64, M = 128
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint64_t dconst = (1ull << 46) | (2ull << 61);
    const uint64_t a_sbo = (uint64_t)(p.sbo >> 4) << 32;
    constexpr uint64_t b_fixed = dconst | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32);
    // Everything that does not depend on the stage is folded into five 64-bit descriptor bases up front; the
    // tile loop then issues its 5 x KSTEPS MMAs with one uniform add each.
    const uint32_t ones16 = smem_u32(smem + p.ones_off) >> 4;
    const uint32_t ks16 = (2u * p.sbo) >> 4;            // 16 pixels = two tile rows per K step
    const uint32_t win8_16 = p.win[8] >> 4;
    uint64_t abase[5];
#pragma unroll
    for (int rb = 0; rb < 5; ++rb) {
      const uint64_t lbo = rb < 4 ? (uint64_t)((p.win[2 * rb + 1] - p.win[2 * rb]) >> 4) << 16 : 0ull;
      abase[rb] = dconst | a_sbo | lbo | (uint64_t)(p.win[2 * rb] >> 4);
    }
    int st = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full_bar[st], ph);
      tc_fence_after();
      const uint32_t a0 = smem_u32(smem + (size_t)st * p.stage_bytes);
      const uint32_t a16 = a0 >> 4;
      const uint64_t b_desc0 = b_fixed | (uint64_t)((a0 + (uint32_t)p.nplanes * p.slot_bytes) >> 4);
      const uint32_t acc_on = it != 0 ? 1u : 0u;
      if (leader) {
#pragma unroll
        for (int rb = 0; rb < 5; ++rb) {
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) {
            const uint32_t addr16 = a16 + (uint32_t)k * ks16;
            uint64_t adesc = abase[rb] + (uint64_t)addr16;
            // fifth accumulator: the second 64-row half is the all-ones block (bias gradient)
            if (rb == 4) adesc |= (uint64_t)((ones16 - (win8_16 + addr16)) & 0x3FFFu) << 16;
            umma_bf16(tmem_base + (uint32_t)(rb * 64), adesc, b_desc0 + (uint64_t)(k * 128), idesc, k != 0 ? 1u : acc_on);
          }
        }
        umma_commit(&empty_bar[st]);
      }
      __syncwarp();
      if (++st == p.stages) { st = 0; ph ^= 1u; }
    }
    if (leader) umma_commit(accum_bar);
    __syncwarp();
  } else {
    // Partials leave the CTA in the accumulators' own shape -- [rb 5][column group 16][row 128][4 floats], 160 KB per
    // CTA -- so that a warp's store instruction covers 32 consecutive rows x 16 bytes = one contiguous 512-byte run.
    // (In the reference layout [tap][ci][co] every lane wrote into another 256-byte row: 32 sectors per instruction,
    // ~5 us of un-overlapped stores at the end of every launch.)  conv_wgrad_halo_reduce_kernel undoes the permutation.
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;  // accumulator row: slot half (m / 64), input channel (m % 64)
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    float4* tile = reinterpret_cast<float4*>(p.part) + ((size_t)split * gridDim.x + blockIdx.x) * (5 * 16 * 128);
#pragma unroll 1
    for (int rb = 0; rb < 5; ++rb) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(rb * 64);
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g)
          tile[((size_t)rb * 16 + (size_t)(c / 4 + g)) * 128 + m] =
              make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                          __uint_as_float(v[4 * g + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// dw[co][ci][tap] (reference layout) and dbias from the tile-native partials of conv_wgrad_halo_kernel:
// part[split][blk = cb * cin_chunks + cc][rb][g][m][4].  A thread owns one float4 and a share of the splits (G groups,
// combined in fixed order through shared memory: deterministic), as in conv_wgrad_reduce_kernel (conv_simt.cu).
struct WgTapTable {
  int t[10];
};
__global__ void __launch_bounds__(256)
    conv_wgrad_halo_reduce_kernel(const float4* __restrict__ part, float* __restrict__ dw, float* __restrict__ dbias, int Cin,
                                  int Cout, int splits, int nblk, int cin_chunks, const WgTapTable taps, int G) {
  __shared__ float4 sm[256];
  const long long total4 = (long long)nblk * (5 * 16 * 128);
  const int E = 256 / G;
  const int e = threadIdx.x % E, g = threadIdx.x / E;
  const long long q = (long long)blockIdx.x * E + e;
  const bool live = q < total4;
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const float4* src = part + q;
    int s = g;
    for (; s + 3 * G < splits; s += 4 * G) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = src[(size_t)(s + u * G) * total4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
      }
    }
    for (; s < splits; s += G) {
      const float4 v = src[(size_t)s * total4];
      acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
    }
  }
  float4 r = make_float4((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                         (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
  if (G > 1) {
    sm[threadIdx.x] = r;
    __syncthreads();
    if (g == 0) {
      for (int k = 1; k < G; ++k) {
        const float4 o = sm[k * E + e];
        r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
      }
    }
  }
  if (g == 0 && live) {
    const int m = (int)(q % 128);
    const int gi = (int)((q / 128) % 16);
    const int rb = (int)((q / 2048) % 5);
    const int blk = (int)(q / 10240);
    const int cc = blk % cin_chunks, cb = blk / cin_chunks;
    const int pos = 2 * rb + (m >> 6);
    const int co = cb * 64 + 4 * gi;
    if (pos < 9) {
      const int ci = cc * 64 + (m & 63);
      float* o = dw + ((size_t)co * Cin + ci) * 9 + taps.t[pos];
      const size_t cs = (size_t)Cin * 9;
      o[0] = r.x; o[cs] = r.y; o[2 * cs] = r.z; o[3 * cs] = r.w;
    } else if (m == 64 && cc == 0 && dbias) {     // the all-ones row: sum over pixels of dy
      dbias[co] = r.x; dbias[co + 1] = r.y; dbias[co + 2] = r.z; dbias[co + 3] = r.w;
    }
  }
}

static bool wgrad_halo_ok(int H, int W, int Cin, int Cout, int stride) {
  if (getenv("BVAE_CONV_NO_HALO")) return false;
  if (Cin % 64 || Cout % 64) return false;
  if (stride == 2 && ((H | W) & 1)) return false;
  const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
  return (stride == 1 || stride == 2) && OH % 8 == 0 && OW % 8 == 0;
}
static int wgrad_halo_th(int OH) { return OH % 16 == 0 ? 16 : 8; }
// stride 2: accumulator position -> tap, sorted by window address inside the four parity planes
static const int kWgS2TapOrder[9] = {4, 3, 5, 1, 7, 0, 2, 6, 8};

bool conv_tc_wgrad_tile_partials(int H, int W, int Cin, int Cout, int stride) { return wgrad_halo_ok(H, W, Cin, Cout, stride); }

int conv_tc_wgrad_tile_reduce(const float* part, float* dw, float* dbias, const WgradTcPlan& pl, int Cin, int Cout, int stride,
                              cudaStream_t st) {
  WgTapTable taps;
  for (int pos = 0; pos < 9; ++pos) taps.t[pos] = stride == 1 ? pos : kWgS2TapOrder[pos];
  taps.t[9] = -1;
  const int nblk = (Cin / 64) * (Cout / 64);
  const long long elems = (long long)nblk * (5 * 16 * 128);
  int G = 1;
  // enough threads to cover the latency of the split walk while every group still has >= 4 splits to sum
  while (G < 16 && elems * G < 64LL * 1024 * 2 && pl.splits / (2 * G) >= 4) G *= 2;
  const int E = 256 / G;
  const long long blocks = (elems + E - 1) / E;
  conv_wgrad_halo_reduce_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(part), dw, dbias, Cin, Cout,
                                                           pl.splits, nblk, Cin / 64, taps, G);
  BVAE_LAUNCH_CHECK("conv_wgrad_halo_reduce");
  return BVAE_OK;
}

static int conv_wgrad_halo(const void* x, const void* dy, float* part, float* bpart, const WgradTcPlan& pl, int B, int H,
                           int W, int Cin, int Cout, int stride, cudaStream_t st) {
  WgHaloMaps maps;
  WgHaloParams p;
  memset(&p, 0, sizeof(p));
  const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
  p.th = wgrad_halo_th(OH);
  p.ksteps = p.th / 2;
  p.cin_chunks = Cin / 64;
  p.co_blocks = Cout / 64;
  p.tiles_w = OW / 8;
  p.tiles_h = OH / p.th;
  p.m_tiles = pl.tiles_total;
  p.tiles_per_split = pl.tiles_per_split;
  p.Cin = Cin; p.Cout = Cout;
  p.part = part;
  (void)bpart;   // the bias sums ride in the fifth accumulator
  p.dy_bytes = (uint32_t)p.th * 8 * 128;
  uint32_t bw, bh;   // plane box: pixels per row, rows
  if (stride == 1) {
    p.nplanes = 1; bw = 10; bh = (uint32_t)p.th + 2;
    for (int j = 0; j < 9; ++j) {
      p.win[j] = (uint32_t)((j / 3) * 10 + (j % 3)) * 128u;
    }
  } else {
    p.nplanes = 4; bw = 9; bh = (uint32_t)p.th + 1;
  }
  p.sbo = bw * 128;
  p.box_bytes = bw * bh * 128;
  p.slot_bytes = (uint32_t)align_up(p.box_bytes, 1024);
  p.stage_bytes = (uint32_t)p.nplanes * p.slot_bytes + p.dy_bytes;
  if (stride == 2) {
    // taps sorted by window address: plane slot q = ph*2+pw with ph = (r != 1), pw = (s != 1); row/col offset (r != 0), (s != 0)
    for (int pos = 0; pos < 9; ++pos) {
      const int t = kWgS2TapOrder[pos], r = t / 3, s = t % 3;
      const int q = (r != 1 ? 2 : 0) + (s != 1 ? 1 : 0);
      p.win[pos] = (uint32_t)q * p.slot_bytes + (uint32_t)((r != 0 ? 1 : 0) * (int)bw + (s != 0 ? 1 : 0)) * 128u;
      if (pos > 0) BVAE_REQUIRE(p.win[pos] > p.win[pos - 1], BVAE_ERR_UNSUPPORTED, "conv_wgrad_halo: window order");
    }
  }
  p.win[9] = 0;
  const size_t budget = 226 * 1024 - 1024 - 4096;
  int stages = (int)(budget / p.stage_bytes);
  p.stages = stages > 5 ? 5 : stages;
  BVAE_REQUIRE(p.stages >= 2, BVAE_ERR_UNSUPPORTED, "conv_wgrad_halo: stage of %u bytes does not fit twice", p.stage_bytes);
  p.ones_off = (uint32_t)p.stages * p.stage_bytes;
  p.bar_off = p.ones_off + 3072;
  const uint32_t xbox[4] = {64u, bw, bh, 1u};
  if (stride == 1) {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    int rc = make_map(&maps.x[0], x, 4, dims, strides, xbox, 128);
    if (rc) return rc;
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)(W / 2), (uint64_t)(H / 2), (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)2 * Cin * 2, (uint64_t)2 * W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const char* base = reinterpret_cast<const char*>(x) + ((size_t)ph * W + pw) * Cin * 2;
        int rc = make_map(&maps.x[ph * 2 + pw], base, 4, dims, strides, xbox, 128);
        if (rc) return rc;
      }
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)OW, (uint64_t)OH, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)OW * Cout * 2, (uint64_t)OH * OW * Cout * 2};
    const uint32_t box[4] = {64u, 8u, (uint32_t)p.th, 1u};
    int rc = make_map(&maps.dy, dy, 4, dims, strides, box, 128);
    if (rc) return rc;
  }
  const size_t smem = 1024 + p.bar_off + (2 * p.stages + 1) * sizeof(uint64_t) + 16;
  auto kern = p.ksteps == 8 ? conv_wgrad_halo_kernel<8> : conv_wgrad_halo_kernel<4>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  if (e != cudaSuccess) return cuda_fail(e, "conv_wgrad_halo: cudaFuncSetAttribute");
  dim3 grid(p.cin_chunks * p.co_blocks, pl.splits);
  kern<<<grid, WG_THREADS, smem, st>>>(maps, p);
  BVAE_LAUNCH_CHECK("conv_wgrad_halo");
  return BVAE_OK;
}

// Column sums of dy (bias gradient), fp32 partials [chunk][C]; memory bound, one pass over dy.
__global__ void __launch_bounds__(256)
    colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ bpart, long long npix, int C, int cv,
                       int rows, int pix_per_chunk) {
  extern __shared__ float sm[];
  const int col = threadIdx.x % cv, row = threadIdx.x / cv;
  const long long p0 = (long long)blockIdx.x * pix_per_chunk, p1 = min(npix, p0 + pix_per_chunk);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (long long p = p0 + row; row < rows && p < p1; p += rows) {
    Vec16<__nv_bfloat16> v;
    v.load(dy + (size_t)p * C + (size_t)col * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += v.v[i];
  }
  if (row < rows) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[row * C + col * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0;
    for (int r = 0; r < rows; ++r) a += (double)sm[r * C + c];
    bpart[(size_t)blockIdx.x * C + c] = (float)a;
  }
}

bool conv_tc_wgrad_supported(int H, int W, int Cin, int Cout, int stride) {
  if (wgrad_halo_ok(H, W, Cin, Cout, stride)) return true;
  if (Cin % 64 || Cout % 64) return false;
  if (Cout > 256 && Cout % 256) return false;
  if (Cout < 256 && (Cout & (Cout - 1))) return false;  // 64, 128, 256
  if (stride == 2 && ((H | W) & 1)) return false;
  const int OW = (W - 1) / stride + 1, OH = (H - 1) / stride + 1;
  int tw, th, tn;
  return tile_geom(OW, OH, &tw, &th, &tn, WG_PIX);
}

WgradTcPlan conv_tc_wgrad_plan(int B, int H, int W, int Cin, int Cout, int stride) {
  WgradTcPlan pl;
  if (wgrad_halo_ok(H, W, Cin, Cout, stride)) {
    pl.bn = 64; pl.slots = 5; pl.groups = (Cin / 64) * (Cout / 64);
    {
      const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
      pl.tiles_total = B * (OH / wgrad_halo_th(OH)) * (OW / 8);
    }
    int splits = num_sms() / pl.groups;
    if (splits < 1) splits = 1;
    if (splits > pl.tiles_total) splits = pl.tiles_total;
    pl.tiles_per_split = (pl.tiles_total + splits - 1) / splits;
    pl.splits = (pl.tiles_total + pl.tiles_per_split - 1) / pl.tiles_per_split;
    pl.bias_chunks = pl.splits;
    pl.bias_pix_per_chunk = 0;
    // tile-native partials: [split][block][5][16][128][4] floats (the bias row lives in the fifth accumulator)
    pl.part_bytes = align_up((size_t)pl.splits * pl.groups * (5 * 16 * 128 * 4) * sizeof(float), 256);
    pl.bpart_bytes = 256;
    return pl;
  }
  const int OW = (W - 1) / stride + 1, OH = (H - 1) / stride + 1;
  int tw, th, tn;
  tile_geom(OW, OH, &tw, &th, &tn, WG_PIX);
  pl.bn = Cout < 256 ? Cout : 256;
  const int units = 9 * (Cin / 64), rbs = (units + 1) / 2;
  int slots = 512 / pl.bn;
  const int nb = pl.bn / 64;
  const int budget = (96 * 1024 - nb * (int)WG_BOX) / (2 * (int)WG_BOX);  // keep a stage under ~96 KB
  if (slots > budget) slots = budget;
  if (slots > rbs) slots = rbs;
  if (slots < 1) slots = 1;
  // balance the row blocks over the groups (e.g. 9 blocks -> 5 + 4 rather than 8 + 1)
  const int groups = (rbs + slots - 1) / slots;
  slots = (rbs + groups - 1) / groups;
  pl.slots = slots;
  pl.groups = groups;
  pl.tiles_total = (OW / tw) * (OH / th) * ((B + tn - 1) / tn);
  const int ctas_fixed = groups * (Cout / pl.bn);
  int splits = (num_sms() + ctas_fixed - 1) / ctas_fixed;
  if (splits > pl.tiles_total) splits = pl.tiles_total;
  if (splits < 1) splits = 1;
  pl.tiles_per_split = (pl.tiles_total + splits - 1) / splits;
  pl.splits = (pl.tiles_total + pl.tiles_per_split - 1) / pl.tiles_per_split;
  const long long npix = (long long)B * OH * OW;
  {
    // enough chunks for ~4 CTAs per SM; a chunk is at least 4 passes of the CTA's pixel rows
    const int cv = Cout / 8, rows = 256 / cv > 0 ? 256 / cv : 1;
    long long per = (npix + 4LL * num_sms() - 1) / (4LL * num_sms());
    if (per < 4LL * rows) per = 4LL * rows;
    pl.bias_pix_per_chunk = (int)per;
  }
  pl.bias_chunks = (int)((npix + pl.bias_pix_per_chunk - 1) / pl.bias_pix_per_chunk);
  pl.part_bytes = align_up((size_t)pl.splits * 9 * Cin * Cout * sizeof(float), 256);
  pl.bpart_bytes = align_up((size_t)pl.bias_chunks * Cout * sizeof(float), 256);
  return pl;
}

int conv_tc_wgrad(const void* x, const void* dy, float* part, float* bpart, const WgradTcPlan& pl, int B, int H,
                  int W, int Cin, int Cout, int stride, cudaStream_t st) {
  if (wgrad_halo_ok(H, W, Cin, Cout, stride)) return conv_wgrad_halo(x, dy, part, bpart, pl, B, H, W, Cin, Cout, stride, st);
  WgMaps maps;
  WgParams p;
  memset(&p, 0, sizeof(p));
  const int OW = (W - 1) / stride + 1, OH = (H - 1) / stride + 1;
  BVAE_REQUIRE(tile_geom(OW, OH, &p.tw, &p.th, &p.tn, WG_PIX), BVAE_ERR_UNSUPPORTED, "conv_tc_wgrad: tile geometry");
  p.tiles_w = OW / p.tw;
  p.tiles_h = OH / p.th;
  p.tiles_total = pl.tiles_total;
  p.tiles_per_split = pl.tiles_per_split;
  p.cin_chunks = Cin / 64;
  p.units = 9 * p.cin_chunks;
  p.slots = pl.slots;
  p.bn = pl.bn;
  p.B = B; p.Cin = Cin; p.Cout = Cout;
  p.part = part;
  const uint32_t box[4] = {64u, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)OW, (uint64_t)OH, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)OW * Cout * 2, (uint64_t)OH * OW * Cout * 2};
    rc = make_map(&maps.dy, dy, 4, dims, strides, box, 128);
    if (rc) return rc;
  }
  if (stride == 1) {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    rc = make_map(&maps.x[0], x, 4, dims, strides, box, 128);
    if (rc) return rc;
    for (int t = 0; t < 9; ++t) p.taps[t] = TcTap{0, (int8_t)(t / 3 - 1), (int8_t)(t % 3 - 1), (int8_t)t};
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)(W / 2), (uint64_t)(H / 2), (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)2 * Cin * 2, (uint64_t)2 * W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const char* base = reinterpret_cast<const char*>(x) + ((size_t)ph * W + pw) * Cin * 2;
        rc = make_map(&maps.x[ph * 2 + pw], base, 4, dims, strides, box, 128);
        if (rc) return rc;
      }
    for (int t = 0; t < 9; ++t) {
      const int r = t / 3, s = t % 3;
      const int ph = (r == 1) ? 0 : 1, di = (r == 0) ? -1 : 0;
      const int pw = (s == 1) ? 0 : 1, dj = (s == 0) ? -1 : 0;
      p.taps[t] = TcTap{(int8_t)(ph * 2 + pw), (int8_t)di, (int8_t)dj, (int8_t)t};
    }
  }
  const uint32_t stage_bytes = (uint32_t)(p.bn / 64 + 2 * p.slots) * WG_BOX;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > 4) stages = 4;
  if (stages < 2) stages = 2;
  p.stages = stages;
  int cols = 32;
  while (cols < p.slots * p.bn) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "conv_tc_wgrad: cudaFuncSetAttribute");
    configured = smem;
  }
  dim3 grid(pl.groups, Cout / p.bn, pl.splits);
  conv_wgrad_tc_kernel<<<grid, WG_THREADS, smem, st>>>(maps, p);
  BVAE_LAUNCH_CHECK("conv_wgrad_tc");
  if (bpart) {
    const int cv = Cout / 8;
    const int rows = 256 / cv > 0 ? 256 / cv : 1;
    const int threads = (cv * rows + 31) / 32 * 32;
    const long long npix = (long long)B * OH * OW;
    colsum_bf16_kernel<<<pl.bias_chunks, threads, (size_t)rows * Cout * sizeof(float), st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), bpart, npix, Cout, cv, rows, pl.bias_pix_per_chunk);
    BVAE_LAUNCH_CHECK("colsum_bf16");
  }
  return BVAE_OK;
}

}  // namespace bvae
