// 3x3 convolution (pad 1, stride 1|2) on the CUDA cores: fprop, dgrad, wgrad.
//
// This is the exact-fp32 path (training.mixed_precision: false, and the reference's fp32
// validation pass) -- tensor cores cannot meet the 1e-5 fp32 parity bar (TF32 has a 10-bit
// mantissa) -- and it also serves bf16 storage for shapes the tcgen05 kernels do not cover.
// Replaces cuDNN/oneDNN conv fprop/dgrad/wgrad behind nn.Conv2d
// (reference: src/models/beta_vae.py:37,42 encoder, :54 decoder).
//
// All four of {fprop s1, fprop s2, dgrad s1, dgrad s2} are one "tap GEMM":
//   out[b, oi, oj, n] = bias[n] + sum_t sum_k src[b, oi*ss + di_t, oj*ss + dj_t, k] * wp[widx_t][k][n]
// written to dst[b, oi*ds + po, oj*ds + pw, n]; out-of-range source pixels read as zero.
//   fprop  : taps (r,s), di = r-1, dj = s-1, ss = stride, ds = 1
//   dgrad 1: src = dy, taps (r,s), di = 1-r, dj = 1-s, weights transposed (k = Cout, n = Cin)
//   dgrad 2: four output parity classes (po,pw); class rows take r with (po+1-r) even,
//            di = (po+1-r)/2, ss = 1, ds = 2   (a transposed conv as 4 dense sub-convs)
// Tile: 128 output pixels x 64 channels per CTA, 8x4 outputs per thread, BK = 16 channels of
// one tap per step, double-buffered shared memory with register prefetch.
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.cuh"

namespace bvae {

struct TapGemm {
  const void* src;
  const void* wp;
  const float* bias;
  void* dst;
  int B, IH, IW, K;     // source [B][IH][IW][K]
  int OH, OW;           // output grid iterated by the GEMM rows
  int DH, DW, N;        // destination [B][DH][DW][N]
  int ss, ds, po, pw;
  int ntaps;
  int di[9], dj[9], widx[9];
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, float* o);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* o) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* o) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  o[0] = __uint_as_float(t.x << 16);
  o[1] = __uint_as_float(t.x & 0xffff0000u);
  o[2] = __uint_as_float(t.y << 16);
  o[3] = __uint_as_float(t.y & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float* v);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

static constexpr int BM = 128, BN = 64, BK = 16, APAD = 4, BPAD = 4;

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) conv_tap_gemm_kernel(const TapGemm p) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN + BPAD];
  const TI* __restrict__ src = reinterpret_cast<const TI*>(p.src);
  const TI* __restrict__ wp = reinterpret_cast<const TI*>(p.wp);
  TO* __restrict__ dst = reinterpret_cast<TO*>(p.dst);

  const int tid = threadIdx.x;
  const long long M = (long long)p.B * p.OH * p.OW;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A loader: rows (tid/4) and (tid/4 + 64), channels (tid%4)*4 .. +3 of the BK chunk
  const int a_kq = (tid & 3) * 4;
  int a_b[2], a_i[2], a_j[2];
  bool a_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long m = m0 + (tid >> 2) + h * 64;
    a_ok[h] = m < M;
    const long long mm = a_ok[h] ? m : 0;
    a_j[h] = (int)(mm % p.OW) * p.ss;
    a_i[h] = (int)((mm / p.OW) % p.OH) * p.ss;
    a_b[h] = (int)(mm / ((long long)p.OW * p.OH));
  }
  // B loader: k row tid/16, columns (tid%16)*4 .. +3
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;

  const int kchunks = (p.K + BK - 1) / BK;
  const int iters = p.ntaps * kchunks;

  float ra[2][4], rb[4];
  auto fetch = [&](int it) {
    const int t = it / kchunks, k0 = (it % kchunks) * BK;
    const int di = p.di[t], dj = p.dj[t];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ih = a_i[h] + di, iw = a_j[h] + dj, k = k0 + a_kq;
      if (a_ok[h] && ih >= 0 && ih < p.IH && iw >= 0 && iw < p.IW && k < p.K) {
        load4<TI>(src + (((size_t)a_b[h] * p.IH + ih) * p.IW + iw) * p.K + k, ra[h]);
      } else {
        ra[h][0] = ra[h][1] = ra[h][2] = ra[h][3] = 0.f;
      }
    }
    const int k = k0 + b_k, n = n0 + b_n;
    if (k < p.K && n < p.N) {
      load4<TI>(wp + ((size_t)p.widx[t] * p.K + k) * p.N + n, rb);
    } else {
      rb[0] = rb[1] = rb[2] = rb[3] = 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 4; ++i) As[buf][a_kq + i][(tid >> 2) + h * 64] = ra[h][i];
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
  };

  const int ty = tid >> 4, tx = tid & 15;  // thread tile: rows ty*8..+7, cols tx*4..+3
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  fetch(0);
  stash(0);
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < iters) fetch(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (it + 1 < iters) stash(buf ^ 1);
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  if (n >= p.N) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = p.bias[n + j];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + ty * 8 + i;
    if (m >= M) continue;
    const int oj = (int)(m % p.OW), oi = (int)((m / p.OW) % p.OH), b = (int)(m / ((long long)p.OW * p.OH));
    const int dh = oi * p.ds + p.po, dw = oj * p.ds + p.pw;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    store4<TO>(dst + (((size_t)b * p.DH + dh) * p.DW + dw) * p.N + n, v);
  }
}

template <typename TI, typename TO>
static int launch_tap_gemm(const TapGemm& p, cudaStream_t st) {
  const long long M = (long long)p.B * p.OH * p.OW;
  if (M == 0) return BVAE_OK;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)cdiv(p.N, BN));
  conv_tap_gemm_kernel<TI, TO><<<grid, 256, 0, st>>>(p);
  BVAE_LAUNCH_CHECK("conv_tap_gemm");
  return BVAE_OK;
}

static int dispatch_tap_gemm(const TapGemm& p, int in_dt, int out_dt, cudaStream_t st) {
  if (in_dt == BVAE_F32 && out_dt == BVAE_F32) return launch_tap_gemm<float, float>(p, st);
  if (in_dt == BVAE_BF16 && out_dt == BVAE_BF16) return launch_tap_gemm<__nv_bfloat16, __nv_bfloat16>(p, st);
  BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "conv3x3: dtype pair (%d,%d)", in_dt, out_dt);
}

// ------------------------------------------------------------------------------------------
// small-Cin forward (first encoder layer: Cin = 1 grayscale / 3 RGB).  Direct, memory bound.
// x is fp32 or act dtype [B][H][W][Cin]; weights come in the layout-0 pack [9][Cin][Cout].
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO, int CIN>
__global__ void __launch_bounds__(256)
    conv_smallcin_fprop_kernel(const TI* __restrict__ x, const TO* __restrict__ wp, const float* __restrict__ bias,
                               TO* __restrict__ y, int B, int H, int W, int Ho, int Wo, int Cout, int stride) {
  constexpr int V = Vec16<TO>::N;
  extern __shared__ float wsm[];  // [9*CIN][Cout] + bias[Cout]
  for (int i = threadIdx.x; i < 9 * CIN * Cout; i += blockDim.x) wsm[i] = Elem<TO>::to_f(wp[i]);
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) wsm[9 * CIN * Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int cv = Cout / V;
  // 32-bit index arithmetic (host checks B*Ho*Wo*cv < 2^31): 64-bit div/mod per output dominated this loop
  const unsigned total = (unsigned)B * Ho * Wo * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % (unsigned)cv);
    unsigned r = i / (unsigned)cv;
    const int ow = (int)(r % (unsigned)Wo);
    r /= (unsigned)Wo;
    const int oh = (int)(r % (unsigned)Ho);
    const int b = (int)(r / (unsigned)Ho);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = wsm[9 * CIN * Cout + c * V + k];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int ih = oh * stride + t / 3 - 1, iw = ow * stride + t % 3 - 1;
      if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
      const TI* px = x + (((size_t)b * H + ih) * W + iw) * CIN;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float xv = Elem<TI>::to_f(px[ci]);
        const float* wr = wsm + (t * CIN + ci) * Cout + c * V;
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = fmaf(xv, wr[k], acc[k]);
      }
    }
    Vec16<TO> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = acc[k];
    o.store(y + (((size_t)b * Ho + oh) * Wo + ow) * Cout + (size_t)c * V);
  }
}

// ------------------------------------------------------------------------------------------
// wgrad, generic:  dWp_t[ci][co] = sum_pix src[pix + tap_t][ci] * dy[pix][co]
// GEMM with M = Cin, N = Cout, K = output pixels; 64x64 tile per (tap, split) CTA;
// deterministic two-stage split-K (partials in workspace, then a reduce that also
// transposes into the reference layout [Cout][Cin][3][3]).
// ------------------------------------------------------------------------------------------
static constexpr int WM = 64, WN = 64, WK = 16;

template <typename TX, typename TD>
__global__ void __launch_bounds__(256)
    conv_wgrad_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, float* __restrict__ part,
                      float* __restrict__ bpart, int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int stride,
                      int pix_per_split) {
  __shared__ __align__(16) float As[2][WK][WM + 4];
  __shared__ __align__(16) float Bs[2][WK][WN + 4];
  const int tid = threadIdx.x;
  const int tiles_n = (Cout + WN - 1) / WN;
  const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
  const int t = blockIdx.y, split = blockIdx.z;
  const int di = t / 3 - 1, dj = t % 3 - 1;
  const long long Mpix = (long long)B * Ho * Wo;
  const long long pbeg = (long long)split * pix_per_split;
  const long long pend = min(Mpix, pbeg + pix_per_split);
  const int iters = (int)((pend - pbeg + WK - 1) / WK);

  const int lk = tid >> 4, lc = (tid & 15) * 4;  // pixel row within chunk, channel quad
  float ra[4], rb[4];
  auto fetch = [&](int it) {
    const long long m = pbeg + (long long)it * WK + lk;
    ra[0] = ra[1] = ra[2] = ra[3] = 0.f;
    rb[0] = rb[1] = rb[2] = rb[3] = 0.f;
    if (m < pend) {
      const int ow = (int)(m % Wo), oh = (int)((m / Wo) % Ho), b = (int)(m / ((long long)Wo * Ho));
      const int ih = oh * stride + di, iw = ow * stride + dj;
      const int ci = tm * WM + lc, co = tn * WN + lc;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W && ci < Cin)
        load4<TX>(x + (((size_t)b * H + ih) * W + iw) * Cin + ci, ra);
      if (co < Cout) load4<TD>(dy + (size_t)m * Cout + co, rb);
    }
  };
  auto stash = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lk][lc]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    *reinterpret_cast<float4*>(&Bs[buf][lk][lc]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
  };
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = (bpart != nullptr) && t == 0 && tm == 0 && ty == 0;

  if (iters > 0) {
    fetch(0);
    stash(0);
  }
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < iters) fetch(it + 1);
#pragma unroll
    for (int k = 0; k < WK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bsum[j] += bv[j];
      }
    }
    if (it + 1 < iters) stash(buf ^ 1);
    __syncthreads();
  }
  // partial layout: [split][t][Cin][Cout]
  float* o = part + ((size_t)split * 9 + t) * Cin * Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = tm * WM + ty * 4 + i;
    if (ci >= Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = tn * WN + tx * 4 + j;
      if (co < Cout) o[(size_t)ci * Cout + co] = acc[i][j];
    }
  }
  if (do_bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = tn * WN + tx * 4 + j;
      if (co < Cout) bpart[(size_t)split * Cout + co] = bsum[j];
    }
  }
}

// few splits: one thread per element
__global__ void __launch_bounds__(256)
    conv_wgrad_reduce_simple_kernel(const float* __restrict__ part, const float* __restrict__ bpart,
                                    float* __restrict__ dw, float* __restrict__ dbias, int Cin, int Cout, int splits,
                                    int bsplits) {
  const long long total = 9LL * Cin * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total + Cout;
       i += (long long)gridDim.x * blockDim.x) {
    if (i < total) {
      const int co = (int)(i % Cout);
      const int ci = (int)((i / Cout) % Cin);
      const int t = (int)(i / ((long long)Cout * Cin));
      float acc = 0.f;
#pragma unroll 4
      for (int s = 0; s < splits; ++s) acc += part[(size_t)s * total + i];
      dw[((size_t)co * Cin + ci) * 9 + t] = acc;
    } else if (dbias) {
      const int co = (int)(i - total);
      float acc = 0.f;
      for (int s = 0; s < bsplits; ++s) acc += bpart[(size_t)s * Cout + co];
      dbias[co] = acc;
    }
  }
}

// dw[co][ci][t] = sum_split part[split][t][ci][co];  dbias[co] = sum_split bpart[split][co]
// A thread owns one float4 (four consecutive co) of the output and a share of the splits: thread (e, g) of a 256-thread
// CTA sums splits g, g + G, ... of element e with four independent 16-byte loads in flight, the G group partials are
// combined in fixed order through shared memory (deterministic), and group 0 scatters the four values into the
// reference layout.  G (1..16, power of two) is chosen on the host so that elements x G fills the machine whether the
// layer has 9K elements and 147 splits (dec3) or 295K elements and 4 splits (dec0).  The round-1 kernel (1024 threads,
// always 16 groups, two barriers per 256 elements) spent its time in barriers: 16 us for 21 MB of L2-resident partials.
// Requires Cout % 4 == 0.
__global__ void __launch_bounds__(256)
    conv_wgrad_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bpart, float* __restrict__ dw,
                             float* __restrict__ dbias, int Cin, int Cout, int splits, int bsplits, int G) {
  __shared__ float4 sm[256];
  const long long total = 9LL * Cin * Cout;
  const int E = 256 / G;                         // elements (float4) per CTA
  const int e = threadIdx.x % E, g = threadIdx.x / E;
  const long long q = (long long)blockIdx.x * E + e;      // float4 index
  const long long i = 4 * q;
  const bool live = i < total;
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const float* src = part + i;
    int s = g;
    for (; s + 3 * G < splits; s += 4 * G) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(src + (size_t)(s + u * G) * total);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
      }
    }
    for (; s < splits; s += G) {
      const float4 v = *reinterpret_cast<const float4*>(src + (size_t)s * total);
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
    const int co = (int)(i % Cout);
    const int ci = (int)((i / Cout) % Cin);
    const int t = (int)(i / ((long long)Cout * Cin));
    float* o = dw + ((size_t)co * Cin + ci) * 9 + t;
    const size_t cs = (size_t)Cin * 9;
    o[0] = r.x; o[cs] = r.y; o[2 * cs] = r.z; o[3 * cs] = r.w;
  }
  if (dbias && blockIdx.x == 0) {
    for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
      float a = 0.f;
      for (int s = 0; s < bsplits; ++s) a += bpart[(size_t)s * Cout + co];
      dbias[co] = a;
    }
  }
}

static int launch_wgrad_reduce(const float* part, const float* bpart, float* dw, float* dbias, int Cin, int Cout,
                               int splits, int bsplits, cudaStream_t st) {
  const long long total = 9LL * Cin * Cout + Cout;
  if (Cout % 4 == 0) {
    const long long elems = (total - Cout) / 4;
    int G = 1;
    // enough threads (>= ~8 warps per scheduler-slot across the chip) while every group still has >= 4 splits to sum
    while (G < 16 && elems * G < 64LL * 1024 * 2 && splits / (2 * G) >= 4) G *= 2;
    const int E = 256 / G;
    const long long blocks = (elems + E - 1) / E;
    conv_wgrad_reduce_kernel<<<(int)blocks, 256, 0, st>>>(part, bpart, dw, dbias, Cin, Cout, splits, bsplits, G);
  } else {
    long long blocks = (total + 255) / 256;
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    conv_wgrad_reduce_simple_kernel<<<(int)blocks, 256, 0, st>>>(part, bpart, dw, dbias, Cin, Cout, splits, bsplits);
  }
  BVAE_LAUNCH_CHECK("conv_wgrad_reduce");
  return BVAE_OK;
}

// ------------------------------------------------------------------------------------------
// small-Cin wgrad (first encoder layer).  Memory bound on dy; one CTA = a run of output pixels
// for one input channel; thread = one 16-byte vector of output channels x 9 taps.
// ------------------------------------------------------------------------------------------
template <typename TX, typename TD>
__global__ void __launch_bounds__(256)
    conv_smallcin_wgrad_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, double* __restrict__ part,
                               double* __restrict__ bpart, int B, int H, int W, int Cin, int Ho, int Wo, int Cout,
                               int stride, int cv, int rows, int pix_per_chunk) {
  constexpr int V = Vec16<TD>::N;
  extern __shared__ float sm[];  // [rows][Cout]
  const int chunk = blockIdx.x, ci = blockIdx.y;
  const int col = threadIdx.x % cv, row = threadIdx.x / cv;
  const long long Mpix = (long long)B * Ho * Wo;
  const long long p0 = (long long)chunk * pix_per_chunk, p1 = min(Mpix, p0 + pix_per_chunk);
  float acc[9][V], bs[V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < V; ++k) acc[t][k] = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) bs[k] = 0.f;
  for (long long m = p0 + row; m < p1; m += rows) {
    Vec16<TD> d;
    d.load(dy + (size_t)m * Cout + (size_t)col * V);
    const unsigned mu = (unsigned)m, rw = mu / (unsigned)Wo;   // 32-bit div/mod (host checks Mpix < 2^31)
    const int ow = (int)(mu - rw * (unsigned)Wo), oh = (int)(rw % (unsigned)Ho), b = (int)(rw / (unsigned)Ho);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int ih = oh * stride + t / 3 - 1, iw = ow * stride + t % 3 - 1;
      float xv = 0.f;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) xv = Elem<TX>::to_f(x[(((size_t)b * H + ih) * W + iw) * Cin + ci]);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[t][k] = fmaf(xv, d.v[k], acc[t][k]);
    }
#pragma unroll
    for (int k = 0; k < V; ++k) bs[k] += d.v[k];
  }
  // reduce across `rows` threads sharing a column, one tap at a time through shared memory
#pragma unroll
  for (int t = 0; t <= 9; ++t) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < V; ++k) sm[row * Cout + col * V + k] = (t < 9) ? acc[t < 9 ? t : 0][k] : bs[k];
    __syncthreads();
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
      double a = 0.0;  // x > 0 and sum(dy) ~ 0 after GroupNorm: keep the cancellation exact-ish
      for (int r = 0; r < rows; ++r) a += (double)sm[r * Cout + c];
      if (t < 9)
        part[(((size_t)chunk * Cin + ci) * 9 + t) * Cout + c] = a;   // double partials: the chunks cancel heavily
      else if (ci == 0 && bpart)
        bpart[(size_t)chunk * Cout + c] = a;
    }
  }
}

// dw[co][ci][t] = sum_chunk part[chunk][ci][t][co]; 64 elements x 4 chunk-groups per CTA, fixed-order combine
__global__ void __launch_bounds__(256)
    conv_smallcin_wgrad_reduce_kernel(const double* __restrict__ part, const double* __restrict__ bpart,
                                      float* __restrict__ dw, float* __restrict__ dbias, int Cin, int Cout,
                                      int chunks) {
  __shared__ double sm[4][64];
  const long long total = 9LL * Cin * Cout;
  const int e = threadIdx.x & 63, grp = threadIdx.x >> 6;
  for (long long base = (long long)blockIdx.x * 64; base < total + Cout; base += (long long)gridDim.x * 64) {
    const long long i = base + e;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < total) {
      int s = grp;
      for (; s + 12 < chunks; s += 16) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += part[(size_t)(s + 4 * u) * total + i];
      }
      for (; s < chunks; s += 4) acc[0] += part[(size_t)s * total + i];
    } else if (i < total + Cout && dbias) {
      const int co = (int)(i - total);
      for (int s = grp; s < chunks; s += 4) acc[0] += bpart[(size_t)s * Cout + co];
    }
    sm[grp][e] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (grp == 0) {
      const double v = (sm[0][e] + sm[1][e]) + (sm[2][e] + sm[3][e]);
      if (i < total) {
        const int co = (int)(i % Cout);
        const int t = (int)((i / Cout) % 9);
        const int ci = (int)(i / (9LL * Cout));
        dw[((size_t)co * Cin + ci) * 9 + t] = (float)v;
      } else if (i < total + Cout && dbias) {
        dbias[(int)(i - total)] = (float)v;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int conv_out(int H, int stride) { return (H - 1) / stride + 1; }  // k=3, pad=1

struct WgradPlan {
  bool small;
  int splits;         // generic: split-K count; small: chunk count
  int pix_per_split;  // pixels per split / chunk
  size_t bytes;
};

static WgradPlan wgrad_plan(int B, int H, int W, int Cin, int Cout, int stride) {
  WgradPlan pl;
  const int Ho = conv_out(H, stride), Wo = conv_out(W, stride);
  const long long Mpix = (long long)B * Ho * Wo;
  pl.small = (Cin % 4 != 0);
  if (pl.small) {
    long long pc = 32768 / (Cout > 0 ? Cout : 1);
    if (pc < 64) pc = 64;
    pl.pix_per_split = (int)pc;
    pl.splits = (int)((Mpix + pc - 1) / pc);
    pl.bytes = align_up((size_t)pl.splits * Cin * 9 * Cout * sizeof(double), 256) +
               align_up((size_t)pl.splits * Cout * sizeof(double), 256);
  } else {
    const long long tiles = (long long)cdiv(Cin, WM) * cdiv(Cout, WN) * 9;
    long long want = (4LL * 148 + tiles - 1) / tiles;  // ~4 waves worth of CTAs
    long long maxs = (Mpix + 255) / 256;                // at least 256 pixels per split
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    long long pps = (Mpix + want - 1) / want;
    pps = (pps + WK - 1) / WK * WK;
    pl.pix_per_split = (int)pps;
    pl.splits = (int)((Mpix + pps - 1) / pps);
    pl.bytes = align_up((size_t)pl.splits * 9 * Cin * Cout * sizeof(float), 256) +
               align_up((size_t)pl.splits * Cout * sizeof(float), 256);
  }
  return pl;
}

template <typename TX, typename TD>
static int wgrad_launch(const void* x, const void* dy, float* dw, float* dbias, void* ws, int B, int H, int W,
                        int Cin, int Cout, int stride, cudaStream_t st) {
  const int Ho = conv_out(H, stride), Wo = conv_out(W, stride);
  const WgradPlan pl = wgrad_plan(B, H, W, Cin, Cout, stride);
  float* part = reinterpret_cast<float*>(ws);
  if (pl.small) {
    constexpr int V = Vec16<TD>::N;
    BVAE_REQUIRE(Cout % V == 0 && Cout / V <= 256, BVAE_ERR_UNSUPPORTED, "conv3x3_wgrad(small Cin): Cout=%d", Cout);
    double* dpart = reinterpret_cast<double*>(ws);
    double* bpart = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) +
                                              align_up((size_t)pl.splits * Cin * 9 * Cout * sizeof(double), 256));
    const int cv = Cout / V, rows = 256 / cv > 0 ? 256 / cv : 1;
    BVAE_REQUIRE((long long)B * Ho * Wo < (1LL << 31), BVAE_ERR_UNSUPPORTED, "conv3x3_wgrad(small Cin): too many pixels");
    dim3 grid(pl.splits, Cin);
    conv_smallcin_wgrad_kernel<TX, TD><<<grid, cv * rows, (size_t)rows * Cout * sizeof(float), st>>>(
        (const TX*)x, (const TD*)dy, dpart, dbias ? bpart : nullptr, B, H, W, Cin, Ho, Wo, Cout, stride, cv, rows,
        pl.pix_per_split);
    BVAE_LAUNCH_CHECK("conv_smallcin_wgrad");
    const long long total = 9LL * Cin * Cout + Cout;
    conv_smallcin_wgrad_reduce_kernel<<<(int)((total + 63) / 64), 256, 0, st>>>(dpart, bpart, dw, dbias, Cin, Cout,
                                                                              pl.splits);
    BVAE_LAUNCH_CHECK("conv_smallcin_wgrad_reduce");
    return BVAE_OK;
  }
  BVAE_REQUIRE(Cout % 4 == 0, BVAE_ERR_UNSUPPORTED, "conv3x3_wgrad: Cout=%d must be a multiple of 4", Cout);
  float* bpart = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) +
                                          align_up((size_t)pl.splits * 9 * Cin * Cout * sizeof(float), 256));
  dim3 grid(cdiv(Cin, WM) * cdiv(Cout, WN), 9, pl.splits);
  conv_wgrad_kernel<TX, TD><<<grid, 256, 0, st>>>((const TX*)x, (const TD*)dy, part, dbias ? bpart : nullptr, B, H, W,
                                                  Cin, Ho, Wo, Cout, stride, pl.pix_per_split);
  BVAE_LAUNCH_CHECK("conv_wgrad");
  return launch_wgrad_reduce(part, bpart, dw, dbias, Cin, Cout, pl.splits, pl.splits, st);
}

}  // namespace bvae

using namespace bvae;

// 1 -> the tcgen05 kernels serve this shape and expect the [t][n][k] pack; 0 -> CUDA-core kernels, [t][k][n].
// BVAE_CONV_SIMT=1 in the environment pins everything to the CUDA-core kernels (debugging aid).
extern "C" int bvae_conv3x3_weight_layout(int H, int W, int Cin, int Cout, int stride, int dtype, int kind) {
  const char* env = getenv("BVAE_CONV_SIMT");  // read per call so tests can flip it
  if ((env != nullptr && env[0] == '1') || dtype != BVAE_BF16) return 0;
  if (stride != 1 && stride != 2) return 0;
  return conv_tc_supported(H, W, Cin, Cout, stride, kind) ? 1 : 0;
}

/* Number of per-sample (sum, sum of squares) partials bvae_conv3x3_fprop_stats writes for this shape; 0 = the kernel
 * that serves the shape has no statistics epilogue (run the fprop without it). */
extern "C" int bvae_conv3x3_fprop_stats_parts(int H, int W, int Cin, int Cout, int stride, int x_dtype, int y_dtype) {
  if (stride != 1 || x_dtype != BVAE_BF16 || y_dtype != BVAE_BF16) return 0;
  if (bvae_conv3x3_weight_layout(H, W, Cin, Cout, stride, BVAE_BF16, 0) != 1) return 0;
  if (!conv_halo_supported(H, W, Cin, Cout)) return 0;
  return conv_halo_stat_parts(H, W, Cout);
}

extern "C" int bvae_conv3x3_fprop_stats(const void* x, const void* wp, const float* bias, void* y, double* stat_part, int B,
                                        int H, int W, int Cin, int Cout, int stride, int x_dtype, int y_dtype,
                                        bvae_stream_t stream) {
  BVAE_REQUIRE(x && wp && y && stat_part, BVAE_ERR_ARG, "conv3x3_fprop_stats: null pointer");
  BVAE_REQUIRE(bvae_conv3x3_fprop_stats_parts(H, W, Cin, Cout, stride, x_dtype, y_dtype) > 0, BVAE_ERR_UNSUPPORTED,
               "conv3x3_fprop_stats: no statistics epilogue for this shape");
  return conv_tc_fprop(x, wp, bias, y, B, H, W, Cin, Cout, stride, as_stream(stream), stat_part);
}

extern "C" int bvae_conv3x3_fprop(const void* x, const void* wp, const float* bias, void* y, int B, int H, int W,
                                  int Cin, int Cout, int stride, int x_dtype, int y_dtype, bvae_stream_t stream) {
  BVAE_REQUIRE(x && wp && y, BVAE_ERR_ARG, "conv3x3_fprop: null pointer");
  BVAE_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, BVAE_ERR_ARG, "conv3x3_fprop: empty shape");
  BVAE_REQUIRE(stride == 1 || stride == 2, BVAE_ERR_UNSUPPORTED, "conv3x3_fprop: stride %d", stride);
  cudaStream_t st = as_stream(stream);
  const int Ho = conv_out(H, stride), Wo = conv_out(W, stride);
  if (bvae_conv3x3_weight_layout(H, W, Cin, Cout, stride, x_dtype == y_dtype ? x_dtype : BVAE_F32, 0) == 1)
    return conv_tc_fprop(x, wp, bias, y, B, H, W, Cin, Cout, stride, st);  // tcgen05 path, wp is the [t][n][k] pack
  if (Cin % 4 != 0) {  // first layer: direct kernel; x may be the fp32 input image
    BVAE_REQUIRE(Cin == 1 || Cin == 3, BVAE_ERR_UNSUPPORTED, "conv3x3_fprop: Cin=%d (need 1, 3 or a multiple of 4)", Cin);
    const size_t sm = (size_t)(9 * Cin + 1) * Cout * sizeof(float);
    BVAE_REQUIRE(sm <= 48 * 1024, BVAE_ERR_UNSUPPORTED, "conv3x3_fprop(small Cin): Cout=%d too large", Cout);
    const int V = (y_dtype == BVAE_BF16) ? 8 : 4;
    BVAE_REQUIRE(Cout % V == 0, BVAE_ERR_UNSUPPORTED, "conv3x3_fprop(small Cin): Cout=%d %% %d", Cout, V);
    const long long total = (long long)B * Ho * Wo * (Cout / V);
    BVAE_REQUIRE(total < (1LL << 31) - (1LL << 24), BVAE_ERR_UNSUPPORTED, "conv3x3_fprop(small Cin): too many outputs");
    long long blocks = (total + 255) / 256;
    if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
#define BVAE_SMALLCIN(TI, TO, CIN)                                                                              \
  conv_smallcin_fprop_kernel<TI, TO, CIN><<<(int)blocks, 256, sm, st>>>((const TI*)x, (const TO*)wp, bias, (TO*)y, \
                                                                       B, H, W, Ho, Wo, Cout, stride)
    if (x_dtype == BVAE_F32 && y_dtype == BVAE_F32) {
      if (Cin == 1) BVAE_SMALLCIN(float, float, 1); else BVAE_SMALLCIN(float, float, 3);
    } else if (x_dtype == BVAE_F32 && y_dtype == BVAE_BF16) {
      if (Cin == 1) BVAE_SMALLCIN(float, __nv_bfloat16, 1); else BVAE_SMALLCIN(float, __nv_bfloat16, 3);
    } else if (x_dtype == BVAE_BF16 && y_dtype == BVAE_BF16) {
      if (Cin == 1) BVAE_SMALLCIN(__nv_bfloat16, __nv_bfloat16, 1); else BVAE_SMALLCIN(__nv_bfloat16, __nv_bfloat16, 3);
    } else {
      BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "conv3x3_fprop: dtype pair (%d,%d)", x_dtype, y_dtype);
    }
#undef BVAE_SMALLCIN
    BVAE_LAUNCH_CHECK("conv_smallcin_fprop");
    return BVAE_OK;
  }
  BVAE_REQUIRE(Cout % 4 == 0, BVAE_ERR_UNSUPPORTED, "conv3x3_fprop: Cout=%d must be a multiple of 4", Cout);
  TapGemm p;
  p.src = x; p.wp = wp; p.bias = bias; p.dst = y;
  p.B = B; p.IH = H; p.IW = W; p.K = Cin;
  p.OH = Ho; p.OW = Wo; p.DH = Ho; p.DW = Wo; p.N = Cout;
  p.ss = stride; p.ds = 1; p.po = 0; p.pw = 0;
  p.ntaps = 9;
  for (int t = 0; t < 9; ++t) {
    p.di[t] = t / 3 - 1;
    p.dj[t] = t % 3 - 1;
    p.widx[t] = t;
  }
  return dispatch_tap_gemm(p, x_dtype, y_dtype, st);
}

extern "C" int bvae_conv3x3_dgrad(const void* dy, const void* wp_t, void* dx, int B, int H, int W, int Cin, int Cout,
                                  int stride, int dtype, bvae_stream_t stream) {
  BVAE_REQUIRE(dy && wp_t && dx, BVAE_ERR_ARG, "conv3x3_dgrad: null pointer");
  BVAE_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, BVAE_ERR_ARG, "conv3x3_dgrad: empty shape");
  BVAE_REQUIRE(stride == 1 || stride == 2, BVAE_ERR_UNSUPPORTED, "conv3x3_dgrad: stride %d", stride);
  BVAE_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0, BVAE_ERR_UNSUPPORTED,
               "conv3x3_dgrad: Cin=%d Cout=%d must be multiples of 4", Cin, Cout);
  cudaStream_t st = as_stream(stream);
  const int Ho = conv_out(H, stride), Wo = conv_out(W, stride);
  if (bvae_conv3x3_weight_layout(H, W, Cin, Cout, stride, dtype, 1) == 1)
    return conv_tc_dgrad(dy, wp_t, dx, B, H, W, Cin, Cout, stride, st);  // tcgen05 path, wp_t is the [t][n][k] pack
  TapGemm p;
  p.src = dy; p.wp = wp_t; p.bias = nullptr; p.dst = dx;
  p.B = B; p.IH = Ho; p.IW = Wo; p.K = Cout;
  p.DH = H; p.DW = W; p.N = Cin;
  p.ss = 1;
  if (stride == 1) {
    p.OH = H; p.OW = W; p.ds = 1; p.po = 0; p.pw = 0;
    p.ntaps = 9;
    for (int t = 0; t < 9; ++t) {
      p.di[t] = 1 - t / 3;
      p.dj[t] = 1 - t % 3;
      p.widx[t] = t;
    }
    return dispatch_tap_gemm(p, dtype, dtype, st);
  }
  for (int ph = 0; ph < 2; ++ph) {
    for (int pw = 0; pw < 2; ++pw) {
      p.OH = (H - ph + 1) / 2;
      p.OW = (W - pw + 1) / 2;
      if (p.OH <= 0 || p.OW <= 0) continue;
      p.ds = 2; p.po = ph; p.pw = pw;
      p.ntaps = 0;
      for (int r = 0; r < 3; ++r) {
        if ((ph + 1 - r) % 2 != 0) continue;
        for (int s = 0; s < 3; ++s) {
          if ((pw + 1 - s) % 2 != 0) continue;
          p.di[p.ntaps] = (ph + 1 - r) / 2;
          p.dj[p.ntaps] = (pw + 1 - s) / 2;
          p.widx[p.ntaps] = r * 3 + s;
          ++p.ntaps;
        }
      }
      const int rc = dispatch_tap_gemm(p, dtype, dtype, st);
      if (rc != BVAE_OK) return rc;
    }
  }
  return BVAE_OK;
}

static bool wgrad_uses_tc(int H, int W, int Cin, int Cout, int stride, int x_dtype, int dy_dtype) {
  const char* env = getenv("BVAE_CONV_SIMT");
  if (env != nullptr && env[0] == '1') return false;
  return x_dtype == BVAE_BF16 && dy_dtype == BVAE_BF16 && conv_tc_wgrad_supported(H, W, Cin, Cout, stride);
}

extern "C" size_t bvae_conv3x3_wgrad_workspace(int B, int H, int W, int Cin, int Cout, int stride) {
  if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (stride != 1 && stride != 2)) return 0;
  size_t bytes = wgrad_plan(B, H, W, Cin, Cout, stride).bytes;
  if (conv_tc_wgrad_supported(H, W, Cin, Cout, stride)) {  // dtype is not known here: cover both kernels
    const WgradTcPlan pl = conv_tc_wgrad_plan(B, H, W, Cin, Cout, stride);
    if (pl.part_bytes + pl.bpart_bytes > bytes) bytes = pl.part_bytes + pl.bpart_bytes;
  } else if (conv_wgrad_mma_small_supported(H, W, Cin, Cout, stride)) {
    const WgradTcPlan pl = conv_wgrad_mma_small_plan(B, H, W, Cin, Cout, stride);
    if (pl.part_bytes + pl.bpart_bytes > bytes) bytes = pl.part_bytes + pl.bpart_bytes;
  }
  return bytes;
}

extern "C" int bvae_conv3x3_wgrad(const void* x, const void* dy, float* dw, float* dbias, void* ws, size_t ws_bytes,
                                  int B, int H, int W, int Cin, int Cout, int stride, int x_dtype, int dy_dtype,
                                  bvae_stream_t stream) {
  BVAE_REQUIRE(x && dy && dw && ws, BVAE_ERR_ARG, "conv3x3_wgrad: null pointer");
  BVAE_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, BVAE_ERR_ARG, "conv3x3_wgrad: empty shape");
  BVAE_REQUIRE(stride == 1 || stride == 2, BVAE_ERR_UNSUPPORTED, "conv3x3_wgrad: stride %d", stride);
  BVAE_REQUIRE(ws_bytes >= bvae_conv3x3_wgrad_workspace(B, H, W, Cin, Cout, stride), BVAE_ERR_WORKSPACE,
               "conv3x3_wgrad: workspace too small");
  cudaStream_t st = as_stream(stream);
  if (wgrad_uses_tc(H, W, Cin, Cout, stride, x_dtype, dy_dtype)) {  // tcgen05 path
    const WgradTcPlan pl = conv_tc_wgrad_plan(B, H, W, Cin, Cout, stride);
    float* part = reinterpret_cast<float*>(ws);
    float* bpart = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + pl.part_bytes);
    const int rc = conv_tc_wgrad(x, dy, part, dbias ? bpart : nullptr, pl, B, H, W, Cin, Cout, stride, st);
    if (rc != BVAE_OK) return rc;
    if (conv_tc_wgrad_tile_partials(H, W, Cin, Cout, stride))
      return conv_tc_wgrad_tile_reduce(part, dw, dbias, pl, Cin, Cout, stride, st);
    return launch_wgrad_reduce(part, bpart, dw, dbias, Cin, Cout, pl.splits, pl.bias_chunks, st);
  }
  {
    const char* env = getenv("BVAE_CONV_SIMT");
    if (!(env != nullptr && env[0] == '1') && x_dtype == BVAE_BF16 && dy_dtype == BVAE_BF16 &&
        conv_wgrad_mma_small_supported(H, W, Cin, Cout, stride)) {   // narrow layers: warp-level MMA path
      const WgradTcPlan pl = conv_wgrad_mma_small_plan(B, H, W, Cin, Cout, stride);
      float* part = reinterpret_cast<float*>(ws);
      float* bpart = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + pl.part_bytes);
      const int rc = conv_wgrad_mma_small(x, dy, part, dbias ? bpart : nullptr, pl, B, H, W, Cin, Cout, stride, st);
      if (rc != BVAE_OK) return rc;
      return launch_wgrad_reduce(part, bpart, dw, dbias, Cin, Cout, pl.splits, pl.bias_chunks, st);
    }
  }
  if (x_dtype == BVAE_F32 && dy_dtype == BVAE_F32)
    return wgrad_launch<float, float>(x, dy, dw, dbias, ws, B, H, W, Cin, Cout, stride, st);
  if (x_dtype == BVAE_BF16 && dy_dtype == BVAE_BF16)
    return wgrad_launch<__nv_bfloat16, __nv_bfloat16>(x, dy, dw, dbias, ws, B, H, W, Cin, Cout, stride, st);
  if (x_dtype == BVAE_F32 && dy_dtype == BVAE_BF16)
    return wgrad_launch<float, __nv_bfloat16>(x, dy, dw, dbias, ws, B, H, W, Cin, Cout, stride, st);
  BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "conv3x3_wgrad: dtype pair (%d,%d)", x_dtype, dy_dtype);
}
