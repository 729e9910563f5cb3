// Focal-frequency loss, spectrum-side arithmetic (reference: src/training/losses.py:23-33):
//   dist = |F|^2;  scale_c = mean_{n,h,w} dist + eps;  w = clamp_min(dist / scale_c, eps)^alpha;  loss = mean(w * dist)
// with the weight NOT detached.  The 2-D FFT itself stays cuFFT (torch.fft.rfft2 on the real difference image --
// the transform is linear, so F(pred) - F(target) = F(pred - target)); these kernels replace the ~25 elementwise /
// reduction launches around it and their autograd.  Input is the ONE-SIDED spectrum [B][C][H][Wh = W/2+1] as
// interleaved (re, im) floats; column j stands for itself and its conjugate, multiplicity m_j = 1 for j = 0 and
// (W even) j = W/2, else 2.  With N = B*H*W elements per channel:
//   s_c = (1/N) sum m p + eps,   L = (1/(C N)) sum m w p,
//   dL/dp = m/(C N) * [ (1+alpha) w  -  alpha * U_c / s_c ]      (unclamped;  U_c = (1/N) sum_{unclamped} m w p)
//   dL/dp = m/(C N) * eps                                         (clamped: w is the constant eps)
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
    ffl_weighted_kernel(const float2* __restrict__ z, const double* __restrict__ part, double* __restrict__ part2,
                        float* __restrict__ scale_out, int B, int C, int H, int Wh, int W, float alpha, float eps) {
  __shared__ double red[32];
  const long long per = (long long)H * Wh;
  const double N = (double)B * H * W;
  for (int c = 0; c < C; ++c) {
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += part[(size_t)c * gridDim.x + i];
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
                                                        float* __restrict__ loss, float* __restrict__ U) {
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
  if (threadIdx.x == 0) loss[0] = (float)(total / (N * C));
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
    float dp;
    if (r < eps) {
      dp = eps;
    } else {
      const float w = alpha == 1.f ? r : powf(r, alpha);
      dp = (1.f + alpha) * w - alpha * U[c] / s;
    }
    const float k = 2.f * g * invCN * ffl_mult((int)(i % Wh), W) * dp;
    gz[i] = make_float2(k * v.x, k * v.y);
  }
}

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
  ffl_weighted_kernel<<<kFflCtas, 256, 0, st>>>(z, part, part2, scale, B, C, H, Wh, W, alpha, eps);
  BVAE_LAUNCH_CHECK("ffl_weighted");
  ffl_final_kernel<<<1, 256, 0, st>>>(part2, kFflCtas, C, (double)B * H * W, loss, U);
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
