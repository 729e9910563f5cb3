// Backward of final_conv (C = 64 -> 1 channel, 3x3, pad 1) on tcgen05 (reference: convolution_backward for
// nn.Conv2d at src/models/beta_vae.py:127,163).  Both products contract against the 9-tap im2col of dlogit,
// D[p][t] = dlogit[p - off_t] (bf16, 16 columns, taps 9..15 zero; built once by a small kernel):
//   dh[p][c] = sum_t D[p][t] * w[c][t]        M = 128 pixels, N = 64, K = 16  -> one MMA per tile
//   dw[c][t] = sum_p h[p][c] * D[p][t]        M = 64 channels (+64 unused rows), N = 16, K = pixels -> 8 MMAs per tile
// The [128 pixels][16] D tile (32-byte rows, SWIZZLE_32B, one TMA box) is the K-major A operand of the first product and
// the MN-major B operand of the second (dev/sw32_mn_probe.cu established the MN-major 32-byte-row descriptor);
// the h tile [128 pixels][64] (SWIZZLE_128B) is the MN-major A operand of the second.  dw accumulates in TMEM over all
// tiles of the persistent CTA and leaves as fp32 partials; the kernel is bound by streaming h in and dh out.
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace bvae {

static constexpr int FB_THREADS = 192;
static constexpr uint32_t FB_H_BYTES = 16384, FB_D_BYTES = 4096, FB_STAGE = FB_H_BYTES + FB_D_BYTES;

struct FbParams {
  int tiles_w, tiles_h, m_tiles;
  int H, W;
  int stages;
  uint32_t wt_off, bar_off;
  __nv_bfloat16* dh;   // [B][H][W][64] or null: first product (D . Wt^T) and its epilogue
  float* wpart;        // [cta][9][64] or null: second product (h^T . D), accumulated over the CTA's tiles
  float* cpart;        // [cta][64] or null: column 9 of the second product (D[:, 9] = 1 -> sum over pixels of h)
};
struct FbMaps {
  CUtensorMap h, d, wt;
};

__device__ __forceinline__ void fb_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(FB_THREADS, 1)
    finalconv_bwd_tc_kernel(const __grid_constant__ FbMaps maps, const FbParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* acc_full = empty_bar + p.stages;   // [2] dh accumulators
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* dw_full = acc_empty + 2;
  uint64_t* wt_bar = dw_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wt_bar + 1);
  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    mbar_init(dw_full, 1);
    mbar_init(wt_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_dw = tmem_base + 128;   // columns 0..127: two dh accumulators; 128..143: dw

  if (warp == 0) {
    // producer: all lanes run the loop, the elected lane issues the TMA loads (see elect_one() in tc_ptx.cuh)
    const bool leader = elect_one();
    if (leader) {
      if (p.wpart) tma_prefetch_desc(&maps.h);
      tma_prefetch_desc(&maps.d);
      if (p.dh) {
        mbar_expect_tx(wt_bar, 2048);
        tma_load_3d(&maps.wt, wt_bar, smem + p.wt_off, 0, 0, 0);
      }
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      mbar_wait(&empty_bar[st], ph ^ 1u);
      int t = tile;
      const int w0 = (t % p.tiles_w) * 8;
      t /= p.tiles_w;
      const int h0 = (t % p.tiles_h) * 16;
      const int b = t / p.tiles_h;
      uint8_t* base = smem + (size_t)st * FB_STAGE;
      if (leader) {
        mbar_expect_tx(&full_bar[st], p.wpart ? FB_STAGE : FB_D_BYTES);
        if (p.wpart) tma_load_4d(&maps.h, &full_bar[st], base, 0, w0, h0, b);
        tma_load_4d(&maps.d, &full_bar[st], base + FB_H_BYTES, 0, w0, h0, b);
      }
      __syncwarp();
      if (++st == p.stages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    // dh: A = D K-major (32-byte rows), B = Wt K-major (32-byte rows), N = 64
    constexpr uint32_t idesc_dh = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // dw: A = h MN-major SW128, B = D MN-major SW32, N = 16
    constexpr uint32_t idesc_dw = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                  ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint64_t dconst = 1ull << 46;
    constexpr uint64_t a_dw_fixed = dconst | (2ull << 61) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32);
    constexpr uint64_t b_dw_fixed = dconst | (6ull << 61) | ((uint64_t)(16 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32);
    if (p.dh) {
      mbar_wait(wt_bar, 0);
      tc_fence_after();
    }
    const uint64_t wt_desc = kmajor_desc(smem_u32(smem + p.wt_off), 32);
    int st = 0;
    uint32_t ph = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      mbar_wait(&full_bar[st], ph);
      if (p.dh) mbar_wait(&acc_empty[buf], ((it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t h_addr = smem_u32(smem + (size_t)st * FB_STAGE), d_addr = h_addr + FB_H_BYTES;
      if (leader) {
        if (p.dh) {
          umma_bf16(tmem_base + buf * 64u, kmajor_desc(d_addr, 32), wt_desc, idesc_dh, 0u);
          umma_commit(&acc_full[buf]);
        }
        if (p.wpart) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(acc_dw, a_dw_fixed | (uint64_t)((h_addr + k * 2048u) >> 4),
                      b_dw_fixed | (uint64_t)((d_addr + k * 512u) >> 4), idesc_dw, (it | (uint32_t)k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[st]);
      }
      __syncwarp();
      if (++st == p.stages) { st = 0; ph ^= 1u; }
    }
    if (leader) umma_commit(dw_full);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // pixel inside the 16x8 tile
    const int hl = r >> 3, wl = r & 7;
    uint32_t it = 0;
    for (int tile = blockIdx.x; p.dh && tile < p.m_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      mbar_wait(&acc_full[buf], (it >> 1) & 1u);
      tc_fence_after();
      {
        int t = tile;
        const int w0 = (t % p.tiles_w) * 8;
        t /= p.tiles_w;
        const int h0 = (t % p.tiles_h) * 16;
        const int b = t / p.tiles_h;
        __nv_bfloat16* out = p.dh + (((size_t)b * p.H + (size_t)(h0 + hl)) * p.W + (size_t)(w0 + wl)) * 64;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 64u;
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
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
      if (lane == 0) fb_mbar_arrive(&acc_empty[buf]);
    }
    // weight-gradient partials: accumulator rows 0..63 = channels, columns 0..8 = taps
    if (p.wpart) {
      mbar_wait(dw_full, 0);
      tc_fence_after();
      if (r < 64) {
        uint32_t v[16];
        tmem_ld16(acc_dw + ((uint32_t)(q * 32) << 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 9; ++t) p.wpart[((size_t)blockIdx.x * 9 + t) * 64 + r] = __uint_as_float(v[t]);
        if (p.cpart) p.cpart[(size_t)blockIdx.x * 64 + r] = __uint_as_float(v[9]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// D[p][t] = dlogit[p - off_t] for t < 9 (zero outside the image), 0 for t >= 9; bpart[cta] = sum of dlogit (bias grad)
__global__ void __launch_bounds__(1024)
    final_im2col16_kernel(const float* __restrict__ dlogit, __nv_bfloat16* __restrict__ dcol, float* __restrict__ bpart,
                          int B, int H, int W) {
  __shared__ float red[32];
  const int npix = B * H * W;
  float bsum = 0.f;
  for (int pidx = blockIdx.x * blockDim.x + threadIdx.x; pidx < npix; pidx += gridDim.x * blockDim.x) {
    const int x = pidx % W, yb = pidx / W, y = yb % H, b = yb / H;
    const float* dl = dlogit + (size_t)b * H * W;
    float d[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);
      d[t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? dl[(size_t)yy * W + xx] : 0.f;
    }
    bsum += d[4];
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(d[2 * i], d[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h2);
    }
    __nv_bfloat162 last = __floats2bfloat162_rn(d[8], 0.f);
    w[4] = *reinterpret_cast<uint32_t*>(&last);
    w[5] = w[6] = w[7] = 0u;
    uint4* dst = reinterpret_cast<uint4*>(dcol + (size_t)pidx * 16);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
  const float s = block_sum(bsum, red);
  if (threadIdx.x == 0) bpart[blockIdx.x] = s;
}

// wt[c][t] = bf16(w[0][c][t]) for t < 9, zero for t >= 9   (w: [1][64][3][3] fp32)
__global__ void final_pack_wt_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 64 * 16) {
    const int c = i >> 4, t = i & 15;
    wt[i] = __float2bfloat16_rn(t < 9 ? w[c * 9 + t] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------
// First encoder layer (Cin = 1, 3x3, stride 2, 64 output channels; reference: ConvBlock.conv at
// src/models/beta_vae.py:37 with in_ch = 1): the same two products against the 16-column im2col of the IMAGE,
//   xcol[p][t] = x[2*oh + r - 1][2*ow + s - 1] (t = r*3+s < 9),  xcol[p][9] = 1,  rest 0:
//   fprop: y[p][co] = sum_t xcol[p][t] * wt[co][t]        with wt[co][9] = bias[co]  (the ones column adds the bias)
//   wgrad: dw[co][t] = sum_p dy[p][co] * xcol[p][t],      column 9 of the same product = db[co]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    first_im2col16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xcol, int B, int H, int W, int Ho, int Wo) {
  const int npix = B * Ho * Wo;
  for (int pidx = blockIdx.x * blockDim.x + threadIdx.x; pidx < npix; pidx += gridDim.x * blockDim.x) {
    const int ow = pidx % Wo, yb = pidx / Wo, oh = yb % Ho, b = yb / Ho;
    const float* img = x + (size_t)b * H * W;
    float d[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int ih = 2 * oh + t / 3 - 1, iw = 2 * ow + t % 3 - 1;
      d[t] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? img[(size_t)ih * W + iw] : 0.f;
    }
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(d[2 * i], d[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h2);
    }
    __nv_bfloat162 last = __floats2bfloat162_rn(d[8], 1.0f);   // column 9 = 1: bias / bias-gradient column
    w[4] = *reinterpret_cast<uint32_t*>(&last);
    w[5] = w[6] = w[7] = 0u;
    uint4* dst = reinterpret_cast<uint4*>(xcol + (size_t)pidx * 16);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// wt[co][t] = bf16(w[co][0][t]) for t < 9, wt[co][9] = bias[co] (or 0), rest zero
__global__ void first_pack_wt_kernel(const float* __restrict__ w, const float* __restrict__ bias, __nv_bfloat16* __restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 64 * 16) {
    const int c = i >> 4, t = i & 15;
    wt[i] = __float2bfloat16_rn(t < 9 ? w[c * 9 + t] : (t == 9 && bias ? bias[c] : 0.f));
  }
}

// dw[co][t] = sum_cta wpart[cta][t][co];  db[co] = sum_cta cpart[cta][co]
__global__ void __launch_bounds__(256)
    first_wgrad_reduce_kernel(const float* __restrict__ wpart, const float* __restrict__ cpart, float* __restrict__ dw,
                              float* __restrict__ db, int nctas) {
  __shared__ float sm[8][32];
  const int e = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + e;     // 0..575: (t, co);  576..639: db
  float a = 0.f;
  if (i < 576) {
    for (int k = grp; k < nctas; k += 8) a += wpart[(size_t)k * 576 + i];
  } else if (i < 640 && db) {
    for (int k = grp; k < nctas; k += 8) a += cpart[(size_t)k * 64 + (i - 576)];
  }
  sm[grp][e] = a;
  __syncthreads();
  if (grp == 0) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) v += sm[g][e];
    if (i < 576) {
      const int t = i / 64, co = i % 64;
      dw[co * 9 + t] = v;
    } else if (i < 640 && db) {
      db[i - 576] = v;
    }
  }
}

static int fb_launch(const void* h64, const void* dcol, const void* wt, __nv_bfloat16* out, float* wpart, float* cpart,
                     int* ncta, int B, int H, int W, cudaStream_t st) {
  FbMaps maps;
  FbParams p;
  memset(&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  const int sms = num_sms();
  p.tiles_w = W / 8;
  p.tiles_h = H / 16;
  p.m_tiles = p.tiles_w * p.tiles_h * B;
  p.H = H; p.W = W;
  p.stages = 6;
  p.wt_off = (uint32_t)p.stages * FB_STAGE;
  p.bar_off = p.wt_off + 2048;
  p.dh = out;
  p.wpart = wpart;
  p.cpart = cpart;
  if (wpart) {
    const uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
    const uint32_t box[4] = {64u, 8u, 16u, 1u};
    int rc = make_map(&maps.h, h64, 4, dims, strides, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {16, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {32, (uint64_t)W * 32, (uint64_t)H * W * 32};
    const uint32_t box[4] = {16u, 8u, 16u, 1u};
    int rc = make_map(&maps.d, dcol, 4, dims, strides, box, 32);
    if (rc) return rc;
  }
  if (out) {
    const uint64_t dims[3] = {16, 64, 1};
    const uint64_t strides[2] = {32, 2048};
    const uint32_t box[3] = {16u, 64u, 1u};
    int rc = make_map(&maps.wt, wt, 3, dims, strides, box, 32);
    if (rc) return rc;
  }
  const size_t smem = 1024 + p.bar_off + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(finalconv_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "finalconv_bwd_tc: cudaFuncSetAttribute");
    configured = true;
  }
  const int grid = p.m_tiles < sms ? p.m_tiles : sms;
  finalconv_bwd_tc_kernel<<<grid, FB_THREADS, smem, st>>>(maps, p);
  BVAE_LAUNCH_CHECK("im2col16_tc");
  if (ncta) *ncta = grid;
  return BVAE_OK;
}

bool finalconv_bwd_tc_supported(int H, int W, int C, int Co) {
  if (getenv("BVAE_CONV_NO_HALO")) return false;
  return C == 64 && Co == 1 && H % 16 == 0 && W % 8 == 0;
}

size_t finalconv_bwd_tc_dcol_bytes(int B, int H, int W) { return align_up((size_t)B * H * W * 16 * 2, 256) + 4096; }

// dlogit [B][1][H][W] fp32 -> dh, and partials wpart [ncta][9][64] / bpart [ncta]; returns the CTA count in *ncta
int finalconv_bwd_tc(const void* h, const float* w, const float* dlogit, void* dh, float* wpart, float* bpart, void* dcol_ws,
                     int* ncta, int B, int H, int W, cudaStream_t st) {
  BVAE_REQUIRE((long long)B * H * W < (1LL << 31) - (1LL << 24), BVAE_ERR_UNSUPPORTED, "finalconv_bwd: too many pixels");
  __nv_bfloat16* dcol = reinterpret_cast<__nv_bfloat16*>(dcol_ws);
  __nv_bfloat16* wt = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(dcol_ws) + align_up((size_t)B * H * W * 32, 256));
  const int sms = num_sms();
  const int tiles = (W / 8) * (H / 16) * B;
  const int grid = tiles < sms ? tiles : sms;
  // same CTA count as the main kernel (one bias partial per CTA); 1024 threads each so that a thread walks ~4 pixels
  final_im2col16_kernel<<<grid, 1024, 0, st>>>(dlogit, dcol, bpart, B, H, W);
  BVAE_LAUNCH_CHECK("final_im2col16");
  final_pack_wt_kernel<<<4, 256, 0, st>>>(w, wt);
  BVAE_LAUNCH_CHECK("final_pack_wt");
  return fb_launch(h, dcol, wt, reinterpret_cast<__nv_bfloat16*>(dh), wpart, nullptr, ncta, B, H, W, st);
}

}  // namespace bvae

using namespace bvae;

// ---- C ABI: first encoder layer on tcgen05 ----
extern "C" int bvae_conv_first_supported(int H, int W, int Cin, int Cout, int stride) {
  if (getenv("BVAE_CONV_NO_HALO") || getenv("BVAE_CONV_SIMT")) {
    const char* e = getenv("BVAE_CONV_SIMT");
    if (getenv("BVAE_CONV_NO_HALO") || (e && e[0] == '1')) return 0;
  }
  if (Cin != 1 || Cout != 64 || stride != 2 || ((H | W) & 1)) return 0;
  return ((H / 2) % 16 == 0 && (W / 2) % 8 == 0) ? 1 : 0;
}

extern "C" int bvae_conv_first_im2col(const float* x, void* xcol, int B, int H, int W, bvae_stream_t stream) {
  BVAE_REQUIRE(x && xcol && B > 0 && H > 0 && W > 0, BVAE_ERR_ARG, "conv_first_im2col: bad argument");
  BVAE_REQUIRE(bvae_conv_first_supported(H, W, 1, 64, 2), BVAE_ERR_UNSUPPORTED, "conv_first_im2col: shape %dx%d", H, W);
  const int Ho = H / 2, Wo = W / 2;
  BVAE_REQUIRE((long long)B * Ho * Wo < (1LL << 31) - (1LL << 24), BVAE_ERR_UNSUPPORTED, "conv_first_im2col: too many pixels");
  long long blocks = ((long long)B * Ho * Wo + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  first_im2col16_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(xcol), B, H, W, Ho, Wo);
  BVAE_LAUNCH_CHECK("first_im2col16");
  return BVAE_OK;
}

extern "C" size_t bvae_conv_first_workspace(void) { return 4096 + (size_t)148 * 4 * (576 + 64) * sizeof(float); }

extern "C" int bvae_conv_first_fprop(const void* xcol, const float* w, const float* bias, void* y, void* ws, size_t ws_bytes,
                                     int B, int Ho, int Wo, bvae_stream_t stream) {
  BVAE_REQUIRE(xcol && w && y && ws && B > 0, BVAE_ERR_ARG, "conv_first_fprop: bad argument");
  BVAE_REQUIRE(Ho % 16 == 0 && Wo % 8 == 0, BVAE_ERR_UNSUPPORTED, "conv_first_fprop: output %dx%d", Ho, Wo);
  BVAE_REQUIRE(ws_bytes >= bvae_conv_first_workspace(), BVAE_ERR_WORKSPACE, "conv_first_fprop: workspace too small");
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* wt = reinterpret_cast<__nv_bfloat16*>(ws);
  first_pack_wt_kernel<<<4, 256, 0, st>>>(w, bias, wt);
  BVAE_LAUNCH_CHECK("first_pack_wt");
  return fb_launch(nullptr, xcol, wt, reinterpret_cast<__nv_bfloat16*>(y), nullptr, nullptr, nullptr, B, Ho, Wo, st);
}

extern "C" int bvae_conv_first_wgrad(const void* xcol, const void* dy, float* dw, float* db, void* ws, size_t ws_bytes, int B,
                                     int Ho, int Wo, bvae_stream_t stream) {
  BVAE_REQUIRE(xcol && dy && dw && ws && B > 0, BVAE_ERR_ARG, "conv_first_wgrad: bad argument");
  BVAE_REQUIRE(Ho % 16 == 0 && Wo % 8 == 0, BVAE_ERR_UNSUPPORTED, "conv_first_wgrad: output %dx%d", Ho, Wo);
  BVAE_REQUIRE(ws_bytes >= bvae_conv_first_workspace(), BVAE_ERR_WORKSPACE, "conv_first_wgrad: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* wpart = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 4096);
  float* cpart = wpart + (size_t)148 * 4 * 576;
  int ncta = 0;
  BVAE_REQUIRE(num_sms() <= 148 * 4, BVAE_ERR_UNSUPPORTED, "conv_first_wgrad: more SMs than the workspace plan");
  const int rc = fb_launch(dy, xcol, nullptr, nullptr, wpart, cpart, &ncta, B, Ho, Wo, st);
  if (rc != BVAE_OK) return rc;
  first_wgrad_reduce_kernel<<<20, 256, 0, st>>>(wpart, cpart, dw, db, ncta);
  BVAE_LAUNCH_CHECK("first_wgrad_reduce");
  return BVAE_OK;
}
