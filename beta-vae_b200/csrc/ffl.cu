// Focal-frequency loss, spectrum-side arithmetic (reference: src/training/losses.py:23-33):
//   dist = |F|^2;  scale_c = mean_{n,h,w} dist + eps;  w = clamp_min(dist / scale_c, eps)^alpha;  loss = mean(w * dist)
// with the weight NOT detached.  The 2-D FFT runs in shared memory, one CTA per image (bvae_ffl_image_fwd / _bwd below:
// the transform is linear, so F(pred) - F(target) = F(pred - target); power-of-two sides up to 128); the spectrum-only
// entry points (bvae_ffl_fwd / _bwd on a cuFFT spectrum) remain for other sizes.  The spectrum is ONE-SIDED [B][C][H][Wh = W/2+1] as
// interleaved (re, im) floats; column j stands for itself and its conjugate, multiplicity m_j = 1 for j = 0 and
// (W even) j = W/2, else 2.  With N = B*H*W elements per channel:
//   s_c = (1/N) sum m p + eps,   L = (1/(C N)) sum m w p,
//   dL/dp = m/(C N) * [ (1+alpha) w  -  alpha * U_c / s_c ]      (unclamped;  U_c = (1/N) sum_{unclamped} m w p)
//   dL/dp = m/(C N) * [ eps^alpha    -  alpha * U_c / s_c ]      (clamped: w is the constant eps^alpha, but s_c still
//                                                                 depends on this bin through the batch mean)
// and the gradient PyTorch expects for the complex input is dL/dp * 2 * z.
#include "common.cuh"

namespace bvae {

static constexpr int kFflCtas = 296;  // 2 per SM

__device__ __forceinline__ float ffl_mult(int j, int W) { return (j == 0 || (2 * j == W)) ? 1.f : 2.f; }

// part[c][cta] = sum over this CTA's elements of channel c of m * p
__global__ void __launch_bounds__(256)
    ffl_power_kernel(const float2* __restrict__ z, double* __restrict__ part, int B, int C, int H, int Wh, int W) {
  __shared__ double red[32];
  // the grid strides over ALL B * H * Wh elements of a channel (a per-sample inner loop left most CTAs idle and the
  // rest walking the batch one dependent load at a time)
  const long long per = (long long)H * Wh, n = (long long)B * per;
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const long long b = i / per, r = i - b * per;
      const float2 v = z[(b * C + c) * per + r];
      acc = fmaf(ffl_mult((int)(r % Wh), W), fmaf(v.x, v.x, v.y * v.y), acc);
    }
    const double s = block_sum((double)acc, red);
    if (threadIdx.x == 0) part[(size_t)c * gridDim.x + blockIdx.x] = s;
  }
}

// every CTA re-derives s_c from the partials, then part2[c][cta] = (sum m w p, sum_{unclamped} m w p)
__global__ void __launch_bounds__(256)
    ffl_weighted_kernel(const float2* __restrict__ z, const double* __restrict__ part, int nparts,
                        double* __restrict__ part2, float* __restrict__ scale_out, int B, int C, int H, int Wh, int W,
                        float alpha, float eps) {
  __shared__ double red[32];
  const long long per = (long long)H * Wh;
  const double N = (double)B * H * W;
  for (int c = 0; c < C; ++c) {
    double t = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) t += part[(size_t)c * nparts + i];
    t = block_sum(t, red);
    const float s = (float)(t / N) + eps;
    if (blockIdx.x == 0 && threadIdx.x == 0) scale_out[c] = s;
    float acc = 0.f, accu = 0.f;
    const long long n = (long long)B * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const long long b = i / per, q = i - b * per;
      const float2 v = z[(b * C + c) * per + q];
      const float p = fmaf(v.x, v.x, v.y * v.y);
      const float r = p / s;
      const bool clamped = r < eps;
      const float w = alpha == 1.f ? fmaxf(r, eps) : powf(fmaxf(r, eps), alpha);
      const float term = ffl_mult((int)(q % Wh), W) * w * p;
      acc += term;
      accu += clamped ? 0.f : term;
    }
    const double a = block_sum((double)acc, red);
    const double u = block_sum((double)accu, red);
    if (threadIdx.x == 0) {
      part2[((size_t)c * gridDim.x + blockIdx.x) * 2] = a;
      part2[((size_t)c * gridDim.x + blockIdx.x) * 2 + 1] = u;
    }
  }
}

// loss = (1/(C N)) sum_c sum;  U_c = (1/N) sum_unclamped
__global__ void __launch_bounds__(256) ffl_final_kernel(const double* __restrict__ part2, int nctas, int C, double N,
                                                        float weight, float* __restrict__ loss, float* __restrict__ U) {
  __shared__ double red[32];
  double total = 0.0;
  for (int c = 0; c < C; ++c) {
    double a = 0.0, u = 0.0;
    for (int i = threadIdx.x; i < nctas; i += blockDim.x) {
      a += part2[((size_t)c * nctas + i) * 2];
      u += part2[((size_t)c * nctas + i) * 2 + 1];
    }
    a = block_sum(a, red);
    u = block_sum(u, red);
    total += a;
    if (threadIdx.x == 0) U[c] = (float)(u / N);
  }
  if (threadIdx.x == 0) loss[0] = weight * (float)(total / (N * C));
}

__global__ void __launch_bounds__(256)
    ffl_bwd_kernel(const float2* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ U,
                   const float* __restrict__ g_loss, float2* __restrict__ gz, int B, int C, int H, int Wh, int W,
                   float alpha, float eps) {
  const long long per = (long long)H * Wh, total = (long long)B * C * per;
  const float g = g_loss[0];
  const float invCN = 1.0f / ((float)C * (float)B * (float)H * (float)W);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i / per) % C);
    const float2 v = z[i];
    const float p = fmaf(v.x, v.x, v.y * v.y);
    const float s = scale[c];
    const float r = p / s;
    // the scale s_c depends on every bin: the -alpha U_c / s_c term applies to clamped bins (constant weight eps^alpha) too
    const float wgt = r < eps ? (alpha == 1.f ? eps : powf(eps, alpha)) : (1.f + alpha) * (alpha == 1.f ? r : powf(r, alpha));
    const float dp = wgt - alpha * U[c] / s;
    const float k = 2.f * g * invCN * ffl_mult((int)(i % Wh), W) * dp;
    gz[i] = make_float2(k * v.x, k * v.y);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// 2-D FFT of the difference image in shared memory (reference: torch.fft.fft2(norm="ortho") at
// src/training/losses.py:23-24).  Radix-2 decimation in time on lines of complex [N + 1]: the loader stores element k at
// bitrev(k), so a pass runs in place and delivers natural order.  A line is owned by ONE warp for all of its log2(N)
// stages (__syncwarp between stages).  Twiddles exp(-2 pi i k / N) come from sincospif (1 ulp), tabulated per CTA.
// ---------------------------------------------------------------------------------------------------------------
static constexpr int kFftMaxN = 128;

__device__ __forceinline__ int ffl_bitrev(int v, int bits) { return (int)(__brev((unsigned)v) >> (32 - bits)); }
__device__ __forceinline__ int ffl_log2(int n) { return 31 - __clz(n); }

// in-place DIT pass over `count` lines of length n; element k of line l is s[l * lstride + k * kstride].
// sign = -1: forward (exp(-i...)), +1: inverse.  tw[j] = exp(-2 pi i j / ntw), j < ntw / 2.
__device__ __forceinline__ void ffl_fft_lines(float2* s, int count, int n, int lstride, int kstride, const float2* tw,
                                              int ntw, float sign) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int bits = ffl_log2(n);
  for (int l = warp; l < count; l += nwarp) {
    float2* line = s + (size_t)l * lstride;
    for (int st = 0; st < bits; ++st) {
      const int half = 1 << st, tstep = ntw >> (st + 1);
      for (int j = lane; j < n / 2; j += 32) {
        const int k = j & (half - 1), i0 = ((j >> st) << (st + 1)) + k, i1 = i0 + half;
        float2 w = tw[k * tstep];
        w.y *= -sign;                         // table holds exp(-i t): conjugate for the inverse
        const float2 a = line[(size_t)i0 * kstride], b = line[(size_t)i1 * kstride];
        const float2 t = make_float2(w.x * b.x - w.y * b.y, w.x * b.y + w.y * b.x);
        line[(size_t)i0 * kstride] = make_float2(a.x + t.x, a.y + t.y);
        line[(size_t)i1 * kstride] = make_float2(a.x - t.x, a.y - t.y);
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ void ffl_twiddles(float2* tw, int ntw) {
  for (int j = threadIdx.x; j < ntw / 2; j += blockDim.x) {
    float sn, cs;
    sincospif(-2.0f * (float)j / (float)ntw, &sn, &cs);
    tw[j] = make_float2(cs, sn);
  }
}

// The 2-D transform is two passes of 1-D line transforms with the one-sided intermediate in global memory (2 MB at
// batch 32, L2-resident): rows along w, then the W/2+1 retained columns along h.  A CTA owns kFftLines lines (one warp
// per line), so a 128 x 128 image is 16 + 9 CTAs per pass and a batch fills every SM many times over -- the single
// CTA-per-image version of this kernel kept 32 SMs busy with 16 latency-bound warps each and took 40 + 52 us.
static constexpr int kFftLines = 8;

// pass 1 (forward): T[img][h][kw] = sum_w (pred - target)[h][w] e^{-2 pi i kw w / W},  kw <= W/2
__global__ void __launch_bounds__(kFftLines * 32)
    ffl_rows_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, float2* __restrict__ T, int H,
                        int W) {
  __shared__ float2 s[kFftLines * (kFftMaxN + 1)];
  __shared__ float2 tw[kFftMaxN / 2];
  const int Wp = W + 1, Wh = W / 2 + 1, wb = ffl_log2(W);
  const int img = blockIdx.y, h0 = blockIdx.x * kFftLines, nl = min(kFftLines, H - h0);
  ffl_twiddles(tw, W);
  const float* pp = pred + ((size_t)img * H + h0) * W;
  const float* tp = target + ((size_t)img * H + h0) * W;
  for (int i = threadIdx.x; i < nl * W; i += blockDim.x) {
    const int l = i / W, w = i - l * W;
    s[l * Wp + ffl_bitrev(w, wb)] = make_float2(pp[i] - tp[i], 0.f);
  }
  __syncthreads();
  ffl_fft_lines(s, nl, W, Wp, 1, tw, W, -1.f);
  __syncthreads();
  float2* out = T + ((size_t)img * H + h0) * Wh;
  for (int i = threadIdx.x; i < nl * Wh; i += blockDim.x) {
    const int l = i / Wh, kw = i - l * Wh;
    out[i] = s[l * Wp + kw];
  }
}

// pass 2 (forward): spec[img][kh][kw] = (1/sqrt(HW)) sum_h T[img][h][kw] e^{-2 pi i kh h / H};  part[c][...] = sum m |F|^2
__global__ void __launch_bounds__(kFftLines * 32)
    ffl_cols_fwd_kernel(const float2* __restrict__ T, float2* __restrict__ spec, double* __restrict__ part, int B, int C,
                        int H, int W) {
  __shared__ float2 s[kFftLines * (kFftMaxN + 1)];
  __shared__ float2 tw[kFftMaxN / 2];
  __shared__ double red[32];
  const int Hp = H + 1, Wh = W / 2 + 1, hb = ffl_log2(H);
  const int img = blockIdx.y, b = img / C, c = img % C;
  const int k0 = blockIdx.x * kFftLines, nl = min(kFftLines, Wh - k0);
  ffl_twiddles(tw, H);
  const float2* in = T + (size_t)img * H * Wh + k0;
  for (int i = threadIdx.x; i < H * nl; i += blockDim.x) {          // nl consecutive kw of one h: a 64-byte segment
    const int h = i / nl, l = i - h * nl;
    s[l * Hp + ffl_bitrev(h, hb)] = in[(size_t)h * Wh + l];
  }
  __syncthreads();
  ffl_fft_lines(s, nl, H, Hp, 1, tw, H, -1.f);
  __syncthreads();
  const float norm = rsqrtf((float)H * (float)W);
  float acc = 0.f;
  float2* out = spec + (size_t)img * H * Wh + k0;
  for (int i = threadIdx.x; i < H * nl; i += blockDim.x) {
    const int kh = i / nl, l = i - kh * nl;
    float2 v = s[l * Hp + kh];
    v.x *= norm;
    v.y *= norm;
    out[(size_t)kh * Wh + l] = v;
    acc = fmaf(ffl_mult(k0 + l, W), fmaf(v.x, v.x, v.y * v.y), acc);
  }
  const double tot = block_sum((double)acc, red);
  if (threadIdx.x == 0) part[((size_t)c * B + b) * gridDim.x + blockIdx.x] = tot;
}

// pass 1 (backward): T[img][h][kw] = sum_kh g(kh, kw) e^{+2 pi i kh h / H} for the retained columns, with
// g = 2 g_loss m(kw) dp z / (C N) the gradient w.r.t. the one-sided spectrum (dp as in ffl_bwd_kernel)
__global__ void __launch_bounds__(kFftLines * 32)
    ffl_cols_bwd_kernel(const float2* __restrict__ spec, const float* __restrict__ scale, const float* __restrict__ U,
                        const float* __restrict__ g_loss, float gmul, float2* __restrict__ T, int B, int C, int H, int W,
                        float alpha, float eps) {
  __shared__ float2 s[kFftLines * (kFftMaxN + 1)];
  __shared__ float2 tw[kFftMaxN / 2];
  const int Hp = H + 1, Wh = W / 2 + 1, hb = ffl_log2(H);
  const int img = blockIdx.y, c = img % C;
  const int k0 = blockIdx.x * kFftLines, nl = min(kFftLines, Wh - k0);
  ffl_twiddles(tw, H);
  const float g = g_loss[0] * gmul;
  const float invCN = 1.0f / ((float)C * (float)B * (float)H * (float)W);
  const float sc = scale[c], uc = U[c];
  const float2* z = spec + (size_t)img * H * Wh + k0;
  for (int i = threadIdx.x; i < H * nl; i += blockDim.x) {
    const int kh = i / nl, l = i - kh * nl;
    const float2 v = z[(size_t)kh * Wh + l];
    const float p = fmaf(v.x, v.x, v.y * v.y);
    const float r = p / sc;
    // w = clamp_min(r, eps)^alpha: constant where clamped; the scale s_c depends on EVERY bin, so the
    // -alpha U_c / s_c term applies to clamped bins too
    const float wgt = r < eps ? (alpha == 1.f ? eps : powf(eps, alpha)) : (1.f + alpha) * (alpha == 1.f ? r : powf(r, alpha));
    const float k = 2.f * g * invCN * ffl_mult(k0 + l, W) * (wgt - alpha * uc / sc);
    s[l * Hp + ffl_bitrev(kh, hb)] = make_float2(k * v.x, k * v.y);
  }
  __syncthreads();
  ffl_fft_lines(s, nl, H, Hp, 1, tw, H, 1.f);
  __syncthreads();
  float2* out = T + (size_t)img * H * Wh + k0;
  for (int i = threadIdx.x; i < H * nl; i += blockDim.x) {
    const int h = i / nl, l = i - h * nl;
    out[(size_t)h * Wh + l] = s[l * Hp + h];
  }
}

// pass 2 (backward): d_pred[img][h][w] = (1/sqrt(HW)) Re sum_{kw <= W/2} T[img][h][kw] e^{+2 pi i kw w / W}
// (the adjoint of the one-sided forward transform: columns beyond W/2 are zero, no conjugate extension)
__global__ void __launch_bounds__(kFftLines * 32)
    ffl_rows_bwd_kernel(const float2* __restrict__ T, float* __restrict__ d_pred, int H, int W) {
  __shared__ float2 s[kFftLines * (kFftMaxN + 1)];
  __shared__ float2 tw[kFftMaxN / 2];
  const int Wp = W + 1, Wh = W / 2 + 1, wb = ffl_log2(W);
  const int img = blockIdx.y, h0 = blockIdx.x * kFftLines, nl = min(kFftLines, H - h0);
  ffl_twiddles(tw, W);
  const float2* in = T + ((size_t)img * H + h0) * Wh;
  for (int i = threadIdx.x; i < nl * W; i += blockDim.x) {
    const int l = i / W, kw = i - l * W;
    s[l * Wp + ffl_bitrev(kw, wb)] = kw < Wh ? in[(size_t)l * Wh + kw] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  ffl_fft_lines(s, nl, W, Wp, 1, tw, W, 1.f);
  __syncthreads();
  const float norm = rsqrtf((float)H * (float)W);
  float* out = d_pred + ((size_t)img * H + h0) * W;
  for (int i = threadIdx.x; i < nl * W; i += blockDim.x) {
    const int l = i / W, w = i - l * W;
    out[i] = s[l * Wp + w].x * norm;
  }
}

static bool ffl_image_ok(int H, int W) {
  auto p2 = [](int v) { return v >= 2 && (v & (v - 1)) == 0 && v <= kFftMaxN; };
  return p2(H) && p2(W);
}
static int ffl_col_blocks(int W) { return cdiv(W / 2 + 1, kFftLines); }

}  // namespace bvae

using namespace bvae;

extern "C" size_t bvae_ffl_workspace(int C) {
  if (C <= 0) return 0;
  return align_up((size_t)C * kFflCtas * 3 * sizeof(double), 256);
}

extern "C" int bvae_ffl_fwd(const float* spec, int B, int C, int H, int Wh, int W, float alpha, float eps, float* loss,
                            float* scale, float* U, void* ws, size_t ws_bytes, bvae_stream_t stream) {
  BVAE_REQUIRE(spec && loss && scale && U && ws, BVAE_ERR_ARG, "ffl_fwd: null pointer");
  BVAE_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && Wh == W / 2 + 1, BVAE_ERR_ARG, "ffl_fwd: bad shape (Wh must be W/2+1)");
  BVAE_REQUIRE(ws_bytes >= bvae_ffl_workspace(C), BVAE_ERR_WORKSPACE, "ffl_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  double* part = reinterpret_cast<double*>(ws);
  double* part2 = part + (size_t)C * kFflCtas;
  const float2* z = reinterpret_cast<const float2*>(spec);
  ffl_power_kernel<<<kFflCtas, 256, 0, st>>>(z, part, B, C, H, Wh, W);
  BVAE_LAUNCH_CHECK("ffl_power");
  ffl_weighted_kernel<<<kFflCtas, 256, 0, st>>>(z, part, kFflCtas, part2, scale, B, C, H, Wh, W, alpha, eps);
  BVAE_LAUNCH_CHECK("ffl_weighted");
  ffl_final_kernel<<<1, 256, 0, st>>>(part2, kFflCtas, C, (double)B * H * W, 1.0f, loss, U);
  BVAE_LAUNCH_CHECK("ffl_final");
  return BVAE_OK;
}

extern "C" int bvae_ffl_bwd(const float* spec, const float* scale, const float* U, const float* g_loss, float* g_spec,
                            int B, int C, int H, int Wh, int W, float alpha, float eps, bvae_stream_t stream) {
  BVAE_REQUIRE(spec && scale && U && g_loss && g_spec, BVAE_ERR_ARG, "ffl_bwd: null pointer");
  BVAE_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && Wh == W / 2 + 1, BVAE_ERR_ARG, "ffl_bwd: bad shape");
  const long long total = (long long)B * C * H * Wh;
  long long blocks = (total + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  ffl_bwd_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(spec), scale, U, g_loss,
                                                           reinterpret_cast<float2*>(g_spec), B, C, H, Wh, W, alpha, eps);
  BVAE_LAUNCH_CHECK("ffl_bwd");
  return BVAE_OK;
}

/* Whole focal-frequency term on images (losses.py:16-35): loss_out = weight * FFL(pred, target); the 2-D FFT runs in
 * shared memory.  pred / target / d_pred are [B][C][H][W] fp32 (NCHW, as the reference holds them); spec is the saved
 * one-sided spectrum [B][C][H][W/2+1][2].  Returns BVAE_ERR_UNSUPPORTED for sides that are not powers of two <= 128
 * (use the spectrum-only entry points on a library FFT then). */
extern "C" int bvae_ffl_image_supported(int H, int W) { return ffl_image_ok(H, W) ? 1 : 0; }

extern "C" size_t bvae_ffl_image_workspace(int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  const size_t parts = ((size_t)C * B * ffl_col_blocks(W) + (size_t)C * kFflCtas * 2) * sizeof(double);
  return align_up(parts, 256) + align_up((size_t)B * C * H * (W / 2 + 1) * sizeof(float2), 256);
}

extern "C" int bvae_ffl_image_fwd(const float* pred, const float* target, int B, int C, int H, int W, float alpha,
                                  float eps, float weight, float* spec, float* loss, float* scale, float* U, void* ws,
                                  size_t ws_bytes, bvae_stream_t stream) {
  BVAE_REQUIRE(pred && target && spec && loss && scale && U && ws, BVAE_ERR_ARG, "ffl_image_fwd: null pointer");
  BVAE_REQUIRE(B > 0 && C > 0 && ffl_image_ok(H, W), BVAE_ERR_UNSUPPORTED, "ffl_image_fwd: %dx%d (powers of two <= %d)", H, W,
               kFftMaxN);
  BVAE_REQUIRE(ws_bytes >= bvae_ffl_image_workspace(B, C, H, W), BVAE_ERR_WORKSPACE, "ffl_image_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int ncb = ffl_col_blocks(W), Wh = W / 2 + 1;
  double* part = reinterpret_cast<double*>(ws);
  double* part2 = part + (size_t)C * B * ncb;
  float2* T = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) +
                                        align_up(((size_t)C * B * ncb + (size_t)C * kFflCtas * 2) * sizeof(double), 256));
  ffl_rows_fwd_kernel<<<dim3(cdiv(H, kFftLines), B * C), kFftLines * 32, 0, st>>>(pred, target, T, H, W);
  BVAE_LAUNCH_CHECK("ffl_rows_fwd");
  ffl_cols_fwd_kernel<<<dim3(ncb, B * C), kFftLines * 32, 0, st>>>(T, reinterpret_cast<float2*>(spec), part, B, C, H, W);
  BVAE_LAUNCH_CHECK("ffl_cols_fwd");
  ffl_weighted_kernel<<<kFflCtas, 256, 0, st>>>(reinterpret_cast<const float2*>(spec), part, B * ncb, part2, scale, B, C, H,
                                                Wh, W, alpha, eps);
  BVAE_LAUNCH_CHECK("ffl_weighted");
  ffl_final_kernel<<<1, 256, 0, st>>>(part2, kFflCtas, C, (double)B * H * W, weight, loss, U);
  BVAE_LAUNCH_CHECK("ffl_final");
  return BVAE_OK;
}

extern "C" int bvae_ffl_image_bwd(const float* spec, const float* scale, const float* U, const float* g_loss, float weight,
                                  float* d_pred, void* ws, size_t ws_bytes, int B, int C, int H, int W, float alpha,
                                  float eps, bvae_stream_t stream) {
  BVAE_REQUIRE(spec && scale && U && g_loss && d_pred && ws, BVAE_ERR_ARG, "ffl_image_bwd: null pointer");
  BVAE_REQUIRE(B > 0 && C > 0 && ffl_image_ok(H, W), BVAE_ERR_UNSUPPORTED, "ffl_image_bwd: %dx%d", H, W);
  BVAE_REQUIRE(ws_bytes >= bvae_ffl_image_workspace(B, C, H, W), BVAE_ERR_WORKSPACE, "ffl_image_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int ncb = ffl_col_blocks(W);
  float2* T = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) +
                                        align_up(((size_t)C * B * ncb + (size_t)C * kFflCtas * 2) * sizeof(double), 256));
  ffl_cols_bwd_kernel<<<dim3(ncb, B * C), kFftLines * 32, 0, st>>>(reinterpret_cast<const float2*>(spec), scale, U, g_loss,
                                                                  weight, T, B, C, H, W, alpha, eps);
  BVAE_LAUNCH_CHECK("ffl_cols_bwd");
  ffl_rows_bwd_kernel<<<dim3(cdiv(H, kFftLines), B * C), kFftLines * 32, 0, st>>>(T, d_pred, H, W);
  BVAE_LAUNCH_CHECK("ffl_rows_bwd");
  return BVAE_OK;
}
