// Bilinear x2 upsample (align_corners=False), NHWC, forward and backward (memory bound).
// Replaces upsample_bilinear2d / upsample_bilinear2d_backward issued by
// nn.Upsample(scale_factor=2, mode='bilinear') in the reference's DeconvBlock
// (src/models/beta_vae.py:52-53).
//
// With scale 2 the interpolation weights are the fixed separable pattern
//   out[2i]   = 0.25 * x[max(i-1,0)] + 0.75 * x[i]
//   out[2i+1] = 0.75 * x[i]          + 0.25 * x[min(i+1,n-1)]
// The backward is the transposed stencil, written as a gather so it needs no atomics:
//   dx[i] = 0.25*d[clamp(2i-1)] + 0.75*d[2i] + 0.75*d[2i+1] + 0.25*d[clamp(2i+2)]
// (clamping the out-of-range tap onto the border row reproduces the border weights).
// Each thread moves one 16-byte channel vector; consecutive threads take consecutive vectors
// of the same pixel, then the next pixel, so loads and stores are fully coalesced.
#include "common.cuh"

namespace bvae {

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int B,
                                                             int H, int W, int cv) {
  constexpr int V = Vec16<T>::N;
  const int C = cv * V;
  // one thread per INPUT vector -> 4 outputs.  32-bit index math: the launcher guarantees B*H*W*cv < 2^31 (64-bit
  // div / mod sequences cost more instructions than the stencil itself)
  const unsigned total = (unsigned)B * H * W * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % (unsigned)cv);
    unsigned r = i / (unsigned)cv;
    const int w = (int)(r % (unsigned)W);
    r /= (unsigned)W;
    const int h = (int)(r % (unsigned)H);
    const int b = (int)(r / (unsigned)H);
    const int hm = max(h - 1, 0), hp = min(h + 1, H - 1), wm = max(w - 1, 0), wp = min(w + 1, W - 1);
    const T* xb = x + (size_t)b * H * W * C + (size_t)c * V;
    Vec16<T> n[3][3];
    const int hs[3] = {hm, h, hp}, ws[3] = {wm, w, wp};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int d = 0; d < 3; ++d) n[a][d].load(xb + ((size_t)hs[a] * W + ws[d]) * C);
    T* yb = y + (size_t)b * 4 * H * W * C + (size_t)c * V;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh) {
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        // rows: dh=0 -> (.25 hm, .75 h); dh=1 -> (.75 h, .25 hp)
        const int ra = dh == 0 ? 0 : 1, rb = dh == 0 ? 1 : 2;
        const float wa = dh == 0 ? 0.25f : 0.75f, wb = dh == 0 ? 0.75f : 0.25f;
        const int ca = dw == 0 ? 0 : 1, cb = dw == 0 ? 1 : 2;
        const float va = dw == 0 ? 0.25f : 0.75f, vb = dw == 0 ? 0.75f : 0.25f;
        Vec16<T> o;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          // separable: rows first then columns, matching ATen's (h-lerp of w-lerps) up to fp32 rounding
          const float top = va * n[ra][ca].v[k] + vb * n[ra][cb].v[k];
          const float bot = va * n[rb][ca].v[k] + vb * n[rb][cb].v[k];
          o.v[k] = wa * top + wb * bot;
        }
        o.store(yb + ((size_t)(2 * h + dh) * (2 * W) + (2 * w + dw)) * C);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B,
                                                             int H, int W, int cv) {
  constexpr int V = Vec16<T>::N;
  const int C = cv * V;
  const int H2 = 2 * H, W2 = 2 * W;
  const unsigned total = (unsigned)B * H * W * cv;     // 32-bit index math, see the forward kernel
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % (unsigned)cv);
    unsigned r = i / (unsigned)cv;
    const int w = (int)(r % (unsigned)W);
    r /= (unsigned)W;
    const int h = (int)(r % (unsigned)H);
    const int b = (int)(r / (unsigned)H);
    const T* db = dy + (size_t)b * H2 * W2 * C + (size_t)c * V;
    const int hs[4] = {max(2 * h - 1, 0), 2 * h, 2 * h + 1, min(2 * h + 2, H2 - 1)};
    const int ws[4] = {max(2 * w - 1, 0), 2 * w, 2 * w + 1, min(2 * w + 2, W2 - 1)};
    const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float rowacc[V];
#pragma unroll
      for (int k = 0; k < V; ++k) rowacc[k] = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        Vec16<T> v;
        v.load(db + ((size_t)hs[a] * W2 + ws[d]) * C);
#pragma unroll
        for (int k = 0; k < V; ++k) rowacc[k] = fmaf(wt[d], v.v[k], rowacc[k]);
      }
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = fmaf(wt[a], rowacc[k], acc[k]);
    }
    Vec16<T> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = acc[k];
    o.store(dx + ((size_t)b * H * W + (size_t)h * W + w) * C + (size_t)c * V);
  }
}

template <typename T>
static int upsample_launch(bool fwd, const void* in, void* out, int B, int H, int W, int C, cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  BVAE_REQUIRE(C % V == 0, BVAE_ERR_UNSUPPORTED, "upsample2x: C=%d must be a multiple of %d", C, V);
  const long long total = (long long)B * H * W * (C / V);
  BVAE_REQUIRE(total < (1LL << 31), BVAE_ERR_UNSUPPORTED, "upsample2x: tensor too large for 32-bit indexing");
  long long blocks = (total + 255) / 256;
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (fwd)
    upsample2x_fwd_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)in, (T*)out, B, H, W, C / V);
  else
    upsample2x_bwd_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)in, (T*)out, B, H, W, C / V);
  BVAE_LAUNCH_CHECK("upsample2x");
  return BVAE_OK;
}

}  // namespace bvae

using namespace bvae;

extern "C" int bvae_upsample2x_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype,
                                   bvae_stream_t stream) {
  BVAE_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0, BVAE_ERR_ARG, "upsample2x_fwd: bad argument");
  if (dtype == BVAE_F32) return upsample_launch<float>(true, x, y, B, H, W, C, as_stream(stream));
  if (dtype == BVAE_BF16) return upsample_launch<__nv_bfloat16>(true, x, y, B, H, W, C, as_stream(stream));
  BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "upsample2x_fwd: dtype %d", dtype);
}

extern "C" int bvae_upsample2x_bwd(const void* dy, void* dx, int B, int H, int W, int C, int dtype,
                                   bvae_stream_t stream) {
  BVAE_REQUIRE(dy && dx && B > 0 && H > 0 && W > 0 && C > 0, BVAE_ERR_ARG, "upsample2x_bwd: bad argument");
  if (dtype == BVAE_F32) return upsample_launch<float>(false, dy, dx, B, H, W, C, as_stream(stream));
  if (dtype == BVAE_BF16) return upsample_launch<__nv_bfloat16>(false, dy, dx, B, H, W, C, as_stream(stream));
  BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "upsample2x_bwd: dtype %d", dtype);
}
