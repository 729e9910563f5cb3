// final_conv (C -> Co in {1,3}, 3x3, pad 1) + sigmoid + reconstruction loss, fused (memory bound:
// the 64-channel input is read once, the output is 1/64th of it).
//
// Replaces conv2d + sigmoid (reference: src/models/beta_vae.py:163-164) and
// mse_loss / l1_loss / binary_cross_entropy with reduction="sum" divided by the batch size
// (src/models/beta_vae.py:185-192), and their backward.
//
// Thread mapping: `cv` = C / V consecutive threads share one output pixel, each owning one
// 16-byte channel vector; the 3x3xC dot product is finished with log2(cv) warp shuffles.
// In the backward the same 9 neighbouring dlogit values give both products at pixel q:
//   dh[q][c]    = sum_t dlogit[q - off_t] * w[c][t]
//   dw[c][t]   += h[q][c] * dlogit[q - off_t]
// so h is read exactly once and dh written exactly once.  dw/db leave each CTA as fp32
// partials that a second kernel sums in fixed order (deterministic).
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.cuh"

namespace bvae {

static constexpr int kFcThreads = 256;
static constexpr int kFcMaxCtas = 148 * 4;

__device__ __forceinline__ float recon_elem_loss(float r, float x, int type) {
  if (type == BVAE_LOSS_MSE) {
    const float d = r - x;
    return d * d;
  }
  if (type == BVAE_LOSS_L1) return fabsf(r - x);
  // F.binary_cross_entropy clamps both log terms at -100
  const float lp = fmaxf(logf(r), -100.f), l1p = fmaxf(logf(1.f - r), -100.f);
  return -(x * lp + (1.f - x) * l1p);
}

// d loss_sum / d r  (before the 1/B factor)
__device__ __forceinline__ float recon_elem_grad(float r, float x, int type) {
  if (type == BVAE_LOSS_MSE) return 2.f * (r - x);
  if (type == BVAE_LOSS_L1) return r > x ? 1.f : (r < x ? -1.f : 0.f);
  return (r - x) / fmaxf((1.f - r) * r, 1e-12f);  // ATen binary_cross_entropy_backward
}

template <typename T>
__global__ void __launch_bounds__(kFcThreads)
    finalconv_fwd_kernel(const T* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias,
                         const float* __restrict__ target, float* __restrict__ recon, double* __restrict__ lpart,
                         int loss_type, int B, int H, int W, int C, int Co, int cv) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float wsm[];  // [Co][9][C]
  __shared__ double red[32];
  for (int i = threadIdx.x; i < Co * 9 * C; i += blockDim.x) {
    const int c = i % C, t = (i / C) % 9, co = i / (9 * C);
    wsm[i] = w[((size_t)co * C + c) * 9 + t];
  }
  __syncthreads();
  const int col = threadIdx.x % cv, row = threadIdx.x / cv, rows = blockDim.x / cv;
  const long long npix = (long long)B * H * W;
  float lsum = 0.f;
  // every thread of a group runs the same trip count so the shuffles below are convergent
  for (long long p0 = (long long)blockIdx.x * rows; p0 < npix; p0 += (long long)gridDim.x * rows) {
    const long long p = p0 + row;
    const bool live = p < npix;
    const long long pp = live ? p : 0;
    const int x = (int)(pp % W), y = (int)((pp / W) % H), b = (int)(pp / ((long long)W * H));
    float acc[3] = {0.f, 0.f, 0.f};
    if (live) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int ih = y + t / 3 - 1, iw = x + t % 3 - 1;
        if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
        Vec16<T> v;
        v.load(h + (((size_t)b * H + ih) * W + iw) * C + (size_t)col * V);
        for (int co = 0; co < Co; ++co) {
          const float* wr = wsm + (co * 9 + t) * C + col * V;
#pragma unroll
          for (int k = 0; k < V; ++k) acc[co] = fmaf(v.v[k], wr[k], acc[co]);
        }
      }
    }
    for (int co = 0; co < Co; ++co) {
      float a = acc[co];
      for (int o = cv >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (live && col == 0) {
        const float logit = a + bias[co];
        const float r = 1.f / (1.f + expf(-logit));
        const size_t o = (((size_t)b * Co + co) * H + y) * W + x;
        recon[o] = r;
        if (target) lsum += recon_elem_loss(r, target[o], loss_type);
      }
    }
  }
  if (lpart) {
    const double s = block_sum((double)lsum, red);
    if (threadIdx.x == 0) lpart[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) loss_final_kernel(const double* __restrict__ part, int n, double scale,
                                                         float* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = (float)(acc * scale);
}

// dlogit[b][co][y][x] = (g_recon + g_loss/B * dL/dr) * r(1-r)
__global__ void __launch_bounds__(256)
    dlogit_kernel(const float* __restrict__ recon, const float* __restrict__ target, const float* __restrict__ g_recon,
                  const float* __restrict__ g_loss, float* __restrict__ dlogit, int loss_type, float invB,
                  long long n) {
  const float gl = g_loss ? g_loss[0] * invB : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float r = recon[i];
    float g = g_recon ? g_recon[i] : 0.f;
    if (target && g_loss) g = fmaf(gl, recon_elem_grad(r, target[i], loss_type), g);
    dlogit[i] = g * r * (1.f - r);
  }
}

template <typename T>
__global__ void __launch_bounds__(kFcThreads)
    finalconv_bwd_kernel(const T* __restrict__ h, const float* __restrict__ w, const float* __restrict__ dlogit,
                         T* __restrict__ dh, float* __restrict__ wpart, float* __restrict__ bpart, int B, int H, int W,
                         int C, int Co, int cv) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sm[];  // weights [Co][9][C], then reduction scratch [rows][C]
  float* wsm = sm;
  float* scratch = sm + Co * 9 * C;
  for (int i = threadIdx.x; i < Co * 9 * C; i += blockDim.x) {
    const int c = i % C, t = (i / C) % 9, co = i / (9 * C);
    wsm[i] = w[((size_t)co * C + c) * 9 + t];
  }
  __syncthreads();
  const int col = threadIdx.x % cv, row = threadIdx.x / cv, rows = blockDim.x / cv;
  const long long npix = (long long)B * H * W;
  for (int co = 0; co < Co; ++co) {
    float wacc[9][V];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < V; ++k) wacc[t][k] = 0.f;
    float bacc = 0.f;
    // 32-bit index arithmetic (host guarantees npix < 2^31): 64-bit div/mod per pixel dominated this loop
    const int npix32 = (int)npix, pstep = (int)gridDim.x * rows;
    for (int p = (int)blockIdx.x * rows + row; p < npix32; p += pstep) {
      const int x = p % W, yb = p / W, y = yb % H, b = yb / H;
      const float* dl = dlogit + ((size_t)b * Co + co) * H * W;
      float d[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);  // q - off_t
        d[t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? dl[(size_t)yy * W + xx] : 0.f;
      }
      const size_t off = (size_t)p * C + (size_t)col * V;
      Vec16<T> hv;
      hv.load(h + off);
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < V; ++k) wacc[t][k] = fmaf(hv.v[k], d[t], wacc[t][k]);
      if (col == 0) bacc += d[4];
      if (dh) {
        Vec16<T> o;
        if (co > 0) o.load(dh + off);
        float a[V];
#pragma unroll
        for (int k = 0; k < V; ++k) a[k] = co > 0 ? o.v[k] : 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          // 16-byte shared-memory reads (the scalar form made this loop LDS-issue bound)
          const float4* wq = reinterpret_cast<const float4*>(wsm + (co * 9 + t) * C + col * V);
#pragma unroll
          for (int q = 0; q < V / 4; ++q) {
            const float4 w4 = wq[q];
            a[4 * q] = fmaf(d[t], w4.x, a[4 * q]);
            a[4 * q + 1] = fmaf(d[t], w4.y, a[4 * q + 1]);
            a[4 * q + 2] = fmaf(d[t], w4.z, a[4 * q + 2]);
            a[4 * q + 3] = fmaf(d[t], w4.w, a[4 * q + 3]);
          }
        }
#pragma unroll
        for (int k = 0; k < V; ++k) o.v[k] = a[k];
        o.store(dh + off);
      }
    }
    // block reduction of dw (one tap at a time) and db
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < V; ++k) scratch[row * C + col * V + k] = wacc[t][k];
      __syncthreads();
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
        for (int r = 0; r < rows; ++r) a += scratch[r * C + c];
        wpart[(((size_t)blockIdx.x * Co + co) * 9 + t) * C + c] = a;
      }
    }
    __syncthreads();
    if (col == 0) scratch[row] = bacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f;
      for (int r = 0; r < rows; ++r) a += scratch[r];
      bpart[(size_t)blockIdx.x * Co + co] = a;
    }
  }
}

// dw[co][c][t] = sum_cta wpart[cta][co][t][c];  db[co] = sum_cta bpart[cta][co]
// 32 elements x 8 CTA-groups per block, four loads in flight per thread, fixed-order combine.
__global__ void __launch_bounds__(256)
    finalconv_wreduce_kernel(const float* __restrict__ wpart, const float* __restrict__ bpart, float* __restrict__ dw,
                             float* __restrict__ db, int nctas, int C, int Co) {
  __shared__ float sm[8][32];
  const int total = Co * 9 * C;
  const int e = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + e;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  if (i < total + Co) {
    const float* src = i < total ? wpart + i : bpart + (i - total);
    const size_t stride = i < total ? (size_t)total : (size_t)Co;
    int k = grp;
    for (; k + 24 < nctas; k += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] += src[(size_t)(k + 8 * u) * stride];
    }
    for (; k < nctas; k += 8) a[0] += src[(size_t)k * stride];
  }
  sm[grp][e] = (a[0] + a[1]) + (a[2] + a[3]);
  __syncthreads();
  if (grp == 0 && i < total + Co) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) v += sm[g][e];
    if (i < total) {
      const int c = i % C, t = (i / C) % 9, co = i / (9 * C);
      dw[((size_t)co * C + c) * 9 + t] = v;
    } else {
      db[i - total] = v;
    }
  }
}

struct FinalGeom {
  int cv, rows, threads, ctas;
};
static bool final_geom(int B, int H, int W, int C, int vec, FinalGeom* g) {
  if (C % vec != 0) return false;
  g->cv = C / vec;
  if (g->cv > 32 || (g->cv & (g->cv - 1)) != 0) return false;  // shuffle reduction needs a power of two <= 32
  g->rows = kFcThreads / g->cv;
  g->threads = kFcThreads;
  const long long npix = (long long)B * H * W;
  long long ctas = (npix + g->rows - 1) / g->rows;
  if (ctas > kFcMaxCtas) ctas = kFcMaxCtas;
  g->ctas = (int)ctas;
  return true;
}

struct FinalWs {
  double* lpart;  // [kFcMaxCtas]
  float* dlogit;  // [B][Co][H][W]
  float* wpart;   // [kFcMaxCtas][Co][9][C]
  float* bpart;   // [kFcMaxCtas][Co]
  __nv_bfloat16* wp16;  // tcgen05 forward weights: [rows <= 32][C] bf16 (conv_halo_final_rows); sized 9*16*C
  void* dcol;           // tcgen05 backward: bf16 im2col of dlogit [B][H][W][16] + packed weights
  size_t bytes;
};
static FinalWs final_ws_carve(void* base, int B, int H, int W, int C, int Co) {
  FinalWs w;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.lpart = reinterpret_cast<double*>(take((size_t)(kFcMaxCtas > 8 * 148 ? kFcMaxCtas : 8 * 148) * sizeof(double)));   // tcgen05 path: 8 epilogue warps per CTA
  w.dlogit = reinterpret_cast<float*>(take((size_t)B * Co * H * W * sizeof(float)));
  w.wpart = reinterpret_cast<float*>(take((size_t)kFcMaxCtas * Co * 9 * C * sizeof(float)));
  w.bpart = reinterpret_cast<float*>(take((size_t)kFcMaxCtas * Co * sizeof(float)));
  w.wp16 = reinterpret_cast<__nv_bfloat16*>(take((size_t)9 * 16 * C * sizeof(__nv_bfloat16)));
  w.dcol = take(finalconv_bwd_tc_supported(H, W, C, Co) ? finalconv_bwd_tc_dcol_bytes(B, H, W) : 0);
  w.bytes = off;
  return w;
}

// w [Co][C][3][3] fp32 -> wg [rows][C] bf16 with row t * Co + co = w[co][:][t]; rows beyond 9 * Co are zero
__global__ void __launch_bounds__(256) final_pack_g_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wg,
                                                           int C, int Co, int rows) {
  const int total = rows * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, n = i / C, t = n / Co, co = n % Co;
    wg[i] = __float2bfloat16_rn(n < 9 * Co ? w[((size_t)co * C + c) * 9 + t] : 0.f);
  }
}

static bool final_tc_env_off() {
  const char* env = getenv("BVAE_CONV_SIMT");
  return env != nullptr && env[0] == '1';
}

template <typename T>
static int final_fwd_impl(const void* h, const float* w, const float* b, const float* target, float* recon,
                          float* loss_out, int loss_type, void* ws, int B, int H, int W, int C, int Co,
                          cudaStream_t st) {
  FinalGeom g;
  BVAE_REQUIRE(final_geom(B, H, W, C, Vec16<T>::N, &g), BVAE_ERR_UNSUPPORTED,
               "finalconv: C=%d must be %d * (power of two <= 32)", C, Vec16<T>::N);
  FinalWs fw = final_ws_carve(ws, B, H, W, C, Co);
  const size_t sm = (size_t)Co * 9 * C * sizeof(float);
  BVAE_REQUIRE(sm <= 48 * 1024, BVAE_ERR_UNSUPPORTED, "finalconv: C=%d too large", C);
  const bool want_loss = target != nullptr && loss_out != nullptr;
  if (sizeof(T) == 2 && !final_tc_env_off() && conv_halo_final_supported(H, W, C, Co)) {
    // tcgen05 path: channel contraction for all nine taps at once on the tensor cores, 3x3 stencil + sigmoid + loss
    // in the epilogue (conv_tc_halo.cu)
    const int rows = conv_halo_final_rows(Co);
    final_pack_g_kernel<<<cdiv(rows * C, 256), 256, 0, st>>>(w, fw.wp16, C, Co, rows);
    BVAE_LAUNCH_CHECK("final_pack_g");
    int n_lpart = 0;
    const int rc = conv_halo_final_fwd(h, fw.wp16, b, want_loss ? target : nullptr, recon, want_loss ? fw.lpart : nullptr,
                                       &n_lpart, loss_type, B, H, W, C, Co, st);
    if (rc != BVAE_OK) return rc;
    if (want_loss) {
      loss_final_kernel<<<1, 256, 0, st>>>(fw.lpart, n_lpart, 1.0 / (double)B, loss_out);
      BVAE_LAUNCH_CHECK("loss_final");
    }
    return BVAE_OK;
  }
  finalconv_fwd_kernel<T><<<g.ctas, g.threads, sm, st>>>((const T*)h, w, b, want_loss ? target : nullptr, recon,
                                                         want_loss ? fw.lpart : nullptr, loss_type, B, H, W, C, Co,
                                                         g.cv);
  BVAE_LAUNCH_CHECK("finalconv_fwd");
  if (want_loss) {
    loss_final_kernel<<<1, 256, 0, st>>>(fw.lpart, g.ctas, 1.0 / (double)B, loss_out);
    BVAE_LAUNCH_CHECK("loss_final");
  }
  return BVAE_OK;
}

template <typename T>
static int final_bwd_impl(const void* h, const float* w, const float* recon, const float* target,
                          const float* g_recon, const float* g_loss, void* dh, float* dw, float* db, int loss_type,
                          void* ws, int B, int H, int W, int C, int Co, cudaStream_t st) {
  FinalGeom g;
  BVAE_REQUIRE(final_geom(B, H, W, C, Vec16<T>::N, &g), BVAE_ERR_UNSUPPORTED,
               "finalconv: C=%d must be %d * (power of two <= 32)", C, Vec16<T>::N);
  FinalWs fw = final_ws_carve(ws, B, H, W, C, Co);
  const long long n = (long long)B * Co * H * W;
  long long blocks = (n + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  dlogit_kernel<<<(int)blocks, 256, 0, st>>>(recon, target, g_recon, g_loss, fw.dlogit, loss_type, 1.f / (float)B, n);
  BVAE_LAUNCH_CHECK("dlogit");
  if (sizeof(T) == 2 && !final_tc_env_off() && finalconv_bwd_tc_supported(H, W, C, Co)) {
    // tcgen05 path: both contractions (dh, dW) run against the 16-column im2col of dlogit
    int ncta = 0;
    const int rc = finalconv_bwd_tc(h, w, fw.dlogit, dh, fw.wpart, fw.bpart, fw.dcol, &ncta, B, H, W, st);
    if (rc != BVAE_OK) return rc;
    finalconv_wreduce_kernel<<<cdiv(Co * 9 * C + Co, 32), 256, 0, st>>>(fw.wpart, fw.bpart, dw, db, ncta, C, Co);
    BVAE_LAUNCH_CHECK("finalconv_wreduce");
    return BVAE_OK;
  }
  const size_t sm = ((size_t)Co * 9 * C + (size_t)g.rows * C) * sizeof(float);
  BVAE_REQUIRE(sm <= 48 * 1024, BVAE_ERR_UNSUPPORTED, "finalconv: C=%d too large", C);
  BVAE_REQUIRE((long long)B * H * W < (1LL << 31) - 65536 * 1024LL, BVAE_ERR_UNSUPPORTED, "finalconv_bwd: too many pixels");
  finalconv_bwd_kernel<T><<<g.ctas, g.threads, sm, st>>>((const T*)h, w, fw.dlogit, (T*)dh, fw.wpart, fw.bpart, B, H,
                                                         W, C, Co, g.cv);
  BVAE_LAUNCH_CHECK("finalconv_bwd");
  finalconv_wreduce_kernel<<<cdiv(Co * 9 * C + Co, 32), 256, 0, st>>>(fw.wpart, fw.bpart, dw, db, g.ctas, C, Co);
  BVAE_LAUNCH_CHECK("finalconv_wreduce");
  return BVAE_OK;
}

}  // namespace bvae

using namespace bvae;

extern "C" size_t bvae_finalconv_workspace(int B, int H, int W, int C, int Co) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || Co <= 0) return 0;
  return final_ws_carve(nullptr, B, H, W, C, Co).bytes;
}

extern "C" int bvae_finalconv_sigmoid_recon_fwd(const void* h, const float* w, const float* b, const float* target,
                                                float* recon, float* loss_out, int loss_type, void* ws,
                                                size_t ws_bytes, int B, int H, int W, int C, int Co, int dtype,
                                                bvae_stream_t stream) {
  BVAE_REQUIRE(h && w && b && recon && ws, BVAE_ERR_ARG, "finalconv_fwd: null pointer");
  BVAE_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, BVAE_ERR_ARG, "finalconv_fwd: empty shape");
  BVAE_REQUIRE(Co >= 1 && Co <= 3, BVAE_ERR_UNSUPPORTED, "finalconv_fwd: Co=%d (1..3)", Co);
  BVAE_REQUIRE(loss_type >= 0 && loss_type <= 2, BVAE_ERR_ARG, "finalconv_fwd: loss_type %d", loss_type);
  BVAE_REQUIRE(ws_bytes >= bvae_finalconv_workspace(B, H, W, C, Co), BVAE_ERR_WORKSPACE, "finalconv_fwd: workspace");
  if (dtype == BVAE_F32)
    return final_fwd_impl<float>(h, w, b, target, recon, loss_out, loss_type, ws, B, H, W, C, Co, as_stream(stream));
  if (dtype == BVAE_BF16)
    return final_fwd_impl<__nv_bfloat16>(h, w, b, target, recon, loss_out, loss_type, ws, B, H, W, C, Co,
                                         as_stream(stream));
  BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "finalconv_fwd: dtype %d", dtype);
}

extern "C" int bvae_finalconv_sigmoid_recon_bwd(const void* h, const float* w, const float* recon,
                                                const float* target, const float* g_recon, const float* g_loss,
                                                void* dh, float* dw, float* db, int loss_type, void* ws,
                                                size_t ws_bytes, int B, int H, int W, int C, int Co, int dtype,
                                                bvae_stream_t stream) {
  BVAE_REQUIRE(h && w && recon && dw && db && ws, BVAE_ERR_ARG, "finalconv_bwd: null pointer");
  BVAE_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, BVAE_ERR_ARG, "finalconv_bwd: empty shape");
  BVAE_REQUIRE(Co >= 1 && Co <= 3, BVAE_ERR_UNSUPPORTED, "finalconv_bwd: Co=%d (1..3)", Co);
  BVAE_REQUIRE(ws_bytes >= bvae_finalconv_workspace(B, H, W, C, Co), BVAE_ERR_WORKSPACE, "finalconv_bwd: workspace");
  if (dtype == BVAE_F32)
    return final_bwd_impl<float>(h, w, recon, target, g_recon, g_loss, dh, dw, db, loss_type, ws, B, H, W, C, Co,
                                 as_stream(stream));
  if (dtype == BVAE_BF16)
    return final_bwd_impl<__nv_bfloat16>(h, w, recon, target, g_recon, g_loss, dh, dw, db, loss_type, ws, B, H, W, C,
                                         Co, as_stream(stream));
  BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "finalconv_bwd: dtype %d", dtype);
}
