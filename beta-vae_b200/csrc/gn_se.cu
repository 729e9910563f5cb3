// GroupNorm(1,C) + ReLU + Squeeze-and-Excitation, fused forward / backward (memory bound): the multi-pass
// kernels and the C-ABI entry points.  Samples that fit a thread-block cluster's shared memory (<= 1 MB) take the
// single-launch cluster-resident kernels of gn_se_cluster.cu instead; these kernels serve larger samples (dec3 of the
// default config), channel counts the cluster mapping does not cover, and BVAE_GN_NO_CLUSTER=1.
//
// Replaces native_group_norm + relu_ + adaptive_avg_pool2d + 2x linear + sigmoid + mul
// (reference: src/models/beta_vae.py:41-46, 61-66; src/models/se_blocks.py:15-19) and their
// autograd backward.
//
// Data layout: y/out/dout/dy are [B][P][C] (NHWC, P = H*W).  One CTA owns a contiguous run
// of pixels of one sample ("chunk"); thread t owns one 16-byte channel vector (column
// t % CV) and walks pixel rows t / CV, +R, +2R ... so every warp-level access is a run of
// consecutive 16-byte vectors.  Per-(sample, channel) sums leave the CTA as fp32 partials
// [B][chunk][C] and are combined by the tiny per-sample kernel that also does the SE
// mat-vecs with warp shuffles; no atomics, so results are run-to-run deterministic.
//
// Forward  : stats pass (read y) -> pool pass (read y) -> excite (tiny) -> scale pass (read y, write out)
// Backward : reduce pass (read dout,y) -> mid (tiny, per sample) -> params (tiny) -> apply pass (read dout,y, write dy)
#include "common.cuh"
#include "gn_cluster.cuh"

namespace bvae {

static constexpr int kGnThreads = 256;
static constexpr float kGnEps = 1e-5f;  // nn.GroupNorm default

struct GnGeom {
  int cv;       // 16-byte vectors per pixel
  int rows;     // pixel rows walked in parallel by one CTA
  int threads;  // cv * rows
  int pc;       // pixels per chunk
  int nchunks;
};

// Pixels per chunk (= per CTA).  The passes are pure streaming, so what matters is bytes in flight and whole waves:
// four CTAs are resident per SM (64 registers x 256 threads), so the grid is sized to AT MOST 4 x SMs CTAs -- one wave --
// as long as a thread then walks no more than 32 vectors; beyond that the tensor is large enough for several waves.
// (Round 1: a fixed 64 KB per CTA left the small deep layers with 32 CTAs on 148 SMs.  Round 2 first capped a thread at
// 16 vectors, which gave dec3's 67 MB tensors 1024 CTAs = 1.73 waves of short-lived CTAs.)
static inline int gn_pixels_per_chunk(int B, int P, int C, int rows) {
  if (rows < 1) rows = 1;
  static const int one_wave = getenv("BVAE_GN_WAVE") == nullptr || atoi(getenv("BVAE_GN_WAVE")) != 0;
  const long long want = 4LL * num_sms();                 // CTAs
  long long nch = one_wave ? want / (B > 0 ? B : 1) : (want + B - 1) / (B > 0 ? B : 1);
  if (nch < 1) nch = 1;
  long long pc = (P + nch - 1) / nch;
  pc = (pc + rows - 1) / rows * rows;                      // whole rounds of the CTA's pixel rows
  if (pc < rows) pc = rows;
  const long long cap = (one_wave ? 32LL : 16LL) * rows;
  if (pc > cap) pc = cap;
  const long long floor_pc = (32768 / (C > 0 ? C : 1)) / 64;   // never below ~1 KB per CTA
  if (pc < floor_pc) pc = floor_pc;
  if (pc > P) pc = P;
  return (int)pc;
}

static inline bool gn_geom(int B, int P, int C, int vec, GnGeom* g) {
  if (C % vec != 0) return false;
  g->cv = C / vec;
  if (g->cv > kGnThreads) return false;
  g->rows = kGnThreads / g->cv;
  if (g->rows < 1) g->rows = 1;
  g->threads = (g->cv * g->rows + 31) / 32 * 32;  // whole warps; surplus threads idle (row >= rows)
  g->pc = gn_pixels_per_chunk(B, P, C, g->rows);
  g->nchunks = cdiv(P, g->pc);
  return true;
}
// upper bound of the chunk count over both element types (workspace sizing does not know the dtype)
static inline int gn_max_chunks(int B, int P, int C) {
  int n = 1;
  for (int vec : {4, 8}) {
    GnGeom g;
    if (gn_geom(B, P, C, vec, &g) && g.nchunks > n) n = g.nchunks;
  }
  return n;
}

// Every CTA re-derives (mean, rstd) of its sample from the stats partials: nchunks doubles.
__device__ __forceinline__ void sample_stats(const double* __restrict__ part, int nchunks, double count,
                                             float* mean_out, float* rstd_out, double* red) {
  double s = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < nchunks; i += blockDim.x) {
    s += part[2 * i];
    ss += part[2 * i + 1];
  }
  s = block_sum(s, red);
  ss = block_sum(ss, red);
  const double mean = s / count;
  double var = ss / count - mean * mean;
  if (var < 0.0) var = 0.0;
  *mean_out = (float)mean;
  *rstd_out = (float)(1.0 / sqrt(var + (double)kGnEps));
}

// ------------------------------------------------------------------ forward: statistics
template <typename T>
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const T* __restrict__ y, double* __restrict__ part,
                                                              int P, int C, int cv, int rows, int pc) {
  __shared__ double red[32];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int col = threadIdx.x % cv, row = threadIdx.x / cv;
  const int p0 = chunk * pc, p1 = min(P, p0 + pc);
  const T* base = y + ((size_t)b * P) * C + (size_t)col * Vec16<T>::N;
  float s = 0.f, ss = 0.f;
#pragma unroll 4
  for (int p = p0 + row; row < rows && p < p1; p += rows) {
    Vec16<T> v;
    v.load(base + (size_t)p * C);
#pragma unroll
    for (int i = 0; i < Vec16<T>::N; ++i) {
      s += v.v[i];
      ss = fmaf(v.v[i], v.v[i], ss);
    }
  }
  double ds = block_sum((double)s, red);
  double dss = block_sum((double)ss, red);
  if (threadIdx.x == 0) {
    double* o = part + ((size_t)b * gridDim.x + chunk) * 2;
    o[0] = ds;
    o[1] = dss;
  }
}

// ------------------------------------------------------------------ forward: pooled sums over positive pixels
template <typename T>
__global__ void __launch_bounds__(kGnThreads)
    gn_pool_kernel(const T* __restrict__ y, const double* __restrict__ spart, int nsp, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ cpart, float* __restrict__ xpart,
                   float* __restrict__ stats, int P, int C, int cv, int rows, int pc) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sm[];  // [rows][cv*V] x 2
  __shared__ double red[32];
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  float mean, rstd;
  sample_stats(spart + (size_t)b * nsp * 2, nsp, (double)P * C, &mean, &rstd, red);
  if (chunk == 0 && threadIdx.x == 0) {
    stats[2 * b] = mean;
    stats[2 * b + 1] = rstd;
  }
  const int col = threadIdx.x % cv, row = threadIdx.x / cv;
  const int p0 = chunk * pc, p1 = min(P, p0 + pc);
  // z = (y - mean) * (rstd * gamma) + beta is the pre-activation in every pass of both kernel families (see
  // gn_se_cluster.cu): the relu mask z > 0 is identical everywhere
  float za[V], bt[V], cnt[V], sx[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    za[i] = rstd * gamma[col * V + i];
    bt[i] = beta[col * V + i];
    cnt[i] = 0.f;
    sx[i] = 0.f;
  }
  const T* base = y + ((size_t)b * P) * C + (size_t)col * V;
  auto accumulate = [&](const uint4& raw) {
    Vec16<T> v;
    v.unpack(raw);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float vm = v.v[i] - mean;
      const bool pos = fmaf(vm, za[i], bt[i]) > 0.f;
      cnt[i] += pos ? 1.f : 0.f;
      sx[i] += pos ? vm : 0.f;
    }
  };
  if (row < rows) {
    int p = p0 + row;
    for (; p + 3 * rows < p1; p += 4 * rows) {     // four independent 16-byte loads in flight per thread
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = Vec16<T>::load_raw(base + (size_t)(p + u * rows) * C);
#pragma unroll
      for (int u = 0; u < 4; ++u) accumulate(raw[u]);
    }
    for (; p < p1; p += rows) accumulate(Vec16<T>::load_raw(base + (size_t)p * C));
  }
#pragma unroll
  for (int i = 0; i < V; ++i) sx[i] *= rstd;
  float* s0 = sm;
  float* s1 = sm + rows * C;
  if (row < rows) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s0[row * C + col * V + i] = cnt[i];
      s1[row * C + col * V + i] = sx[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0, q = 0.0;  // fixed-order, double: these sums cancel heavily downstream
    for (int r = 0; r < rows; ++r) {
      a += (double)s0[r * C + c];
      q += (double)s1[r * C + c];
    }
    const size_t o = ((size_t)b * nchunks + chunk) * C + c;
    cpart[o] = (float)a;
    xpart[o] = (float)q;
  }
}

// ------------------------------------------------------------------ forward: (mean, rstd) from the partials, blocks without SE
// one warp per sample
__global__ void __launch_bounds__(256)
    gn_finalize_stats_kernel(const double* __restrict__ spart, int nsp, double count, float* __restrict__ stats, int B) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  double s = 0.0, ss = 0.0;
  for (int i = lane; i < nsp; i += 32) {
    s += spart[((size_t)b * nsp + i) * 2];
    ss += spart[((size_t)b * nsp + i) * 2 + 1];
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (lane == 0) {
    const double mean = s / count;
    double var = ss / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * b] = (float)mean;
    stats[2 * b + 1] = (float)(1.0 / sqrt(var + (double)kGnEps));
  }
}

// ------------------------------------------------------------------ forward: excitation (one CTA per sample)
// Latency is what this kernel costs (one CTA per sample, everything else waits for the gate): every phase is written as
// ONE batch of independent loads per thread -- chunk partials, W1 rows (a warp per hidden unit, 16-byte lane loads), W2
// rows (a thread per channel, 16-byte loads) -- instead of load-use chains.
static constexpr int kExThreads = 512;
__global__ void __launch_bounds__(kExThreads)
    se_excite_kernel(const float* __restrict__ cpart, const float* __restrict__ xpart, int nchunks,
                     const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ w1,
                     const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                     float* __restrict__ cpos, float* __restrict__ spos, float* __restrict__ hidden,
                     float* __restrict__ gate, int P, int C, int R) {
  extern __shared__ float sm[];  // pool[C] + hid[R]
  float* pool = sm;
  float* hid = sm + C;
  const int b = blockIdx.x;
  __shared__ double cs0[kExThreads], cs1[kExThreads];
  // chunk sums: thread (c, g) adds chunks g, g+G, ... (G thread groups when C is small), then a fixed-order combine
  const int G = (C < (int)blockDim.x) ? (int)blockDim.x / C : 1;
  for (int c0 = 0; c0 < C; c0 += blockDim.x) {
    const int c = (G > 1) ? (int)threadIdx.x % C : c0 + (int)threadIdx.x;
    const int g = (G > 1) ? (int)threadIdx.x / C : 0;
    double a = 0.0, q = 0.0;
    if (c < C && g < G) {
      const float* cp = cpart + (size_t)b * nchunks * C + c;
      const float* xp = xpart + (size_t)b * nchunks * C + c;
      double av[4] = {0.0, 0.0, 0.0, 0.0}, qv[4] = {0.0, 0.0, 0.0, 0.0};
      int k = g;
      for (; k + 3 * G < nchunks; k += 4 * G) {
        float cf[4], xf[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          cf[u] = cp[(size_t)(k + u * G) * C];
          xf[u] = xp[(size_t)(k + u * G) * C];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          av[u] += (double)cf[u];
          qv[u] += (double)xf[u];
        }
      }
      for (; k < nchunks; k += G) {
        av[0] += (double)cp[(size_t)k * C];
        qv[0] += (double)xp[(size_t)k * C];
      }
      a = (av[0] + av[1]) + (av[2] + av[3]);
      q = (qv[0] + qv[1]) + (qv[2] + qv[3]);
    }
    if (G > 1) {
      cs0[threadIdx.x] = a;
      cs1[threadIdx.x] = q;
      __syncthreads();
      if (g == 0 && c < C) {
        for (int gg = 1; gg < G; ++gg) {
          a += cs0[gg * C + c];
          q += cs1[gg * C + c];
        }
      }
    }
    if (g == 0 && c < C) {
      const float as = (float)a, qs = (float)q;
      cpos[(size_t)b * C + c] = as;
      spos[(size_t)b * C + c] = qs;
      pool[c] = (gamma[c] * qs + beta[c] * as) / (float)P;  // mean over pixels of relu(gn(y))
    }
    if (G > 1) break;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int j = warp; j < R; j += nwarp) {  // warp-shuffle mat-vec: hidden = relu(W1 pool + b1)
    const float* wr = w1 + (size_t)j * C;
    float acc = 0.f;
    if ((C & 127) == 0) {              // 16-byte lane loads, all of the row's loads in flight at once
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int c = lane * 4; c < C; c += 128) {
        const float4 wv = *reinterpret_cast<const float4*>(wr + c);
        const float4 pv = *reinterpret_cast<const float4*>(pool + c);
        a4[0] = fmaf(wv.x, pv.x, a4[0]);
        a4[1] = fmaf(wv.y, pv.y, a4[1]);
        a4[2] = fmaf(wv.z, pv.z, a4[2]);
        a4[3] = fmaf(wv.w, pv.w, a4[3]);
      }
      acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    } else {
      for (int c = lane; c < C; c += 32) acc = fmaf(wr[c], pool[c], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float h = fmaxf(acc + b1[j], 0.f);
      hid[j] = h;
      hidden[(size_t)b * R + j] = h;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {  // gate = sigmoid(W2 hidden + b2)
    float acc = b2[c];
    const float* wr = w2 + (size_t)c * R;
    if ((R & 3) == 0) {
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
      const float4* wq = reinterpret_cast<const float4*>(wr);
      const float4* hq = reinterpret_cast<const float4*>(hid);
#pragma unroll 8
      for (int j = 0; j < R / 4; ++j) {
        const float4 w4 = wq[j], h4 = hq[j];
        a4[0] = fmaf(w4.x, h4.x, a4[0]);
        a4[1] = fmaf(w4.y, h4.y, a4[1]);
        a4[2] = fmaf(w4.z, h4.z, a4[2]);
        a4[3] = fmaf(w4.w, h4.w, a4[3]);
      }
      acc += (a4[0] + a4[1]) + (a4[2] + a4[3]);
    } else {
#pragma unroll 8
      for (int j = 0; j < R; ++j) acc = fmaf(wr[j], hid[j], acc);
    }
    gate[(size_t)b * C + c] = 1.0f / (1.0f + expf(-acc));
  }
}

// ------------------------------------------------------------------ forward: normalise + relu + rescale
template <typename T>
__global__ void __launch_bounds__(kGnThreads)
    gn_scale_kernel(const T* __restrict__ y, T* __restrict__ out, const float* __restrict__ stats,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ gate,
                    int P, int C, int cv, int rows, int pc) {
  constexpr int V = Vec16<T>::N;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];   // written by the pool pass / gn_finalize_stats_kernel
  const int col = threadIdx.x % cv, row = threadIdx.x / cv;
  const int p0 = chunk * pc, p1 = min(P, p0 + pc);
  float zs[V], bs_[V];   // gate folded in: relu(z) * s = max(z * s, 0) for the sigmoid gate s > 0
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float sc = gate ? gate[(size_t)b * C + col * V + i] : 1.f;
    zs[i] = rstd * gamma[col * V + i] * sc;
    bs_[i] = beta[col * V + i] * sc;
  }
  const size_t off = ((size_t)b * P) * C + (size_t)col * V;
#pragma unroll 4
  for (int p = p0 + row; row < rows && p < p1; p += rows) {
    Vec16<T> v;
    v.load(y + off + (size_t)p * C);
#pragma unroll
    for (int i = 0; i < V; ++i) v.v[i] = fmaxf(fmaf(v.v[i] - mean, zs[i], bs_[i]), 0.f);
    v.store(out + off + (size_t)p * C);
  }
}

// ------------------------------------------------------------------ backward: per-(sample,channel) reductions
template <typename T>
__global__ void __launch_bounds__(kGnThreads)
    gn_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ y, const float* __restrict__ stats,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ a1part, float* __restrict__ a2part, int P, int C, int cv, int rows,
                         int pc) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sm[];
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  const int col = threadIdx.x % cv, row = threadIdx.x / cv;
  const int p0 = chunk * pc, p1 = min(P, p0 + pc);
  float za[V], bt[V], a1[V], a2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    za[i] = rstd * gamma[col * V + i];
    bt[i] = beta[col * V + i];
    a1[i] = 0.f;
    a2[i] = 0.f;
  }
  const size_t off = ((size_t)b * P) * C + (size_t)col * V;
  auto accumulate = [&](const uint4& yraw, const uint4& draw) {
    Vec16<T> v, d;
    v.unpack(yraw);
    d.unpack(draw);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float vm = v.v[i] - mean;
      const float dm = (fmaf(vm, za[i], bt[i]) > 0.f) ? d.v[i] : 0.f;
      a1[i] += dm;
      a2[i] = fmaf(dm, vm, a2[i]);
    }
  };
  if (row < rows) {
    int p = p0 + row;
    for (; p + 3 * rows < p1; p += 4 * rows) {     // eight independent 16-byte loads in flight per thread
      uint4 yr[4], dr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        yr[u] = Vec16<T>::load_raw(y + off + (size_t)(p + u * rows) * C);
        dr[u] = Vec16<T>::load_raw(dout + off + (size_t)(p + u * rows) * C);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) accumulate(yr[u], dr[u]);
    }
    for (; p < p1; p += rows)
      accumulate(Vec16<T>::load_raw(y + off + (size_t)p * C), Vec16<T>::load_raw(dout + off + (size_t)p * C));
  }
#pragma unroll
  for (int i = 0; i < V; ++i) a2[i] *= rstd;
  float* s0 = sm;
  float* s1 = sm + rows * C;
  if (row < rows) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s0[row * C + col * V + i] = a1[i];
      s1[row * C + col * V + i] = a2[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0, q = 0.0;  // fixed-order, double: these sums cancel heavily downstream
    for (int r = 0; r < rows; ++r) {
      a += (double)s0[r * C + c];
      q += (double)s1[r * C + c];
    }
    const size_t o = ((size_t)b * nchunks + chunk) * C + c;
    a1part[o] = (float)a;
    a2part[o] = (float)q;
  }
}

// ------------------------------------------------------------------ backward: per-sample SE + GN coefficients
// Per sample b (one CTA):  A1[c] = sum_p dout*m, A2[c] = sum_p dout*m*xhat  (m = relu mask)
//   ds      = gamma*A2 + beta*A1                     (d gate, since a = relu(gamma*xhat+beta))
//   dz2     = ds * s(1-s);  dh = W2^T dz2;  dz1 = dh * (hidden>0);  dpool = W1^T dz1
//   G1      = s*A1 + dpool/P * cpos    (= sum_p g,       g = d relu-input)
//   G2      = s*A2 + dpool/P * spos    (= sum_p g*xhat)
//   S1, S2  = sum_c gamma*G1, sum_c gamma*G2          (GroupNorm backward sums)
__global__ void __launch_bounds__(kExThreads)
    gn_bwd_mid_kernel(const float* __restrict__ a1part, const float* __restrict__ a2part, int nchunks,
                      const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ w1,
                      const float* __restrict__ w2, const float* __restrict__ cpos, const float* __restrict__ spos,
                      const float* __restrict__ hidden, const float* __restrict__ gate, float* __restrict__ G1o,
                      float* __restrict__ G2o, float* __restrict__ dz2o, float* __restrict__ dz1o,
                      float* __restrict__ dpoolo, float* __restrict__ coef, int P, int C, int R) {
  extern __shared__ float sm[];  // A1[C] A2[C] dz2[C] dz1[R]
  __shared__ double red[32];
  float* A1 = sm;
  float* A2 = sm + C;
  float* dz2 = sm + 2 * C;
  float* dz1 = sm + 3 * C;
  const int b = blockIdx.x;
  const bool has_se = (w1 != nullptr);
  __shared__ double cs0[kExThreads], cs1[kExThreads];
  const int G = (C < (int)blockDim.x) ? (int)blockDim.x / C : 1;   // thread groups sharing the chunk loop (see se_excite_kernel)
  for (int c0 = 0; c0 < C; c0 += blockDim.x) {
    const int c = (G > 1) ? (int)threadIdx.x % C : c0 + (int)threadIdx.x;
    const int g = (G > 1) ? (int)threadIdx.x / C : 0;
    double ad = 0.0, qd = 0.0;
    if (c < C && g < G) {
      const float* p1 = a1part + (size_t)b * nchunks * C + c;
      const float* p2 = a2part + (size_t)b * nchunks * C + c;
      double adv[4] = {0.0, 0.0, 0.0, 0.0}, qdv[4] = {0.0, 0.0, 0.0, 0.0};
      int k = g;
      for (; k + 3 * G < nchunks; k += 4 * G) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          adv[u] += (double)p1[(size_t)(k + u * G) * C];
          qdv[u] += (double)p2[(size_t)(k + u * G) * C];
        }
      }
      for (; k < nchunks; k += G) {
        adv[0] += (double)p1[(size_t)k * C];
        qdv[0] += (double)p2[(size_t)k * C];
      }
      ad = (adv[0] + adv[1]) + (adv[2] + adv[3]);
      qd = (qdv[0] + qdv[1]) + (qdv[2] + qdv[3]);
    }
    if (G > 1) {
      cs0[threadIdx.x] = ad;
      cs1[threadIdx.x] = qd;
      __syncthreads();
      if (g == 0 && c < C) {
        for (int gg = 1; gg < G; ++gg) {
          ad += cs0[gg * C + c];
          qd += cs1[gg * C + c];
        }
      }
    }
    if (g == 0 && c < C) {
      const float a = (float)ad, q = (float)qd;
      A1[c] = a;
      A2[c] = q;
      if (has_se) {
        const double s = (double)gate[(size_t)b * C + c];
        const double ds = (double)gamma[c] * qd + (double)beta[c] * ad;   // the two terms cancel: keep the double sums
        const float t = (float)(ds * s * (1.0 - s));
        dz2[c] = t;
        dz2o[(size_t)b * C + c] = t;
      }
    }
    if (G > 1) break;
  }
  __syncthreads();
  if (has_se) {
    // dh[j] = sum_c dz2[c] * W2[c][j].  W2 is [C][R]: threads run along j (coalesced rows), channel groups in
    // parallel, eight independent loads per trip; the group partials are combined in double (signed terms cancel).
    const int ng = max(1, (int)blockDim.x / R);
    const int jj = threadIdx.x % R, cgp = threadIdx.x / R;
    float part = 0.f;
    if (cgp < ng) {
      int c = cgp;
      for (; c + 7 * ng < C; c += 8 * ng) {
        float wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) wv[u] = w2[(size_t)(c + u * ng) * R + jj];
#pragma unroll
        for (int u = 0; u < 8; ++u) part = fmaf(dz2[c + u * ng], wv[u], part);
      }
      for (; c < C; c += ng) part = fmaf(dz2[c], w2[(size_t)c * R + jj], part);
    }
    __syncthreads();                       // A1/A2/dz2 are complete; reuse cs0 as the group scratch
    float* scr = reinterpret_cast<float*>(cs0);
    if (cgp < ng) scr[cgp * R + jj] = part;
    __syncthreads();
    if ((int)threadIdx.x < R) {
      double acc = 0.0;
      for (int q = 0; q < ng; ++q) acc += (double)scr[q * R + threadIdx.x];
      const float t = hidden[(size_t)b * R + threadIdx.x] > 0.f ? (float)acc : 0.f;
      dz1[threadIdx.x] = t;
      dz1o[(size_t)b * R + threadIdx.x] = t;
    }
    __syncthreads();
  }
  double s1 = 0.0, s2 = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float g1, g2;
    if (has_se) {
      float dp = 0.f;
      {
        int j = 0;
        for (; j + 8 <= R; j += 8) {            // eight independent (coalesced) loads per trip
          float wv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) wv[u] = w1[(size_t)(j + u) * C + c];
#pragma unroll
          for (int u = 0; u < 8; ++u) dp = fmaf(dz1[j + u], wv[u], dp);
        }
        for (; j < R; ++j) dp = fmaf(dz1[j], w1[(size_t)j * C + c], dp);
      }
      dp /= (float)P;
      dpoolo[(size_t)b * C + c] = dp;
      const float s = gate[(size_t)b * C + c];
      g1 = fmaf(dp, cpos[(size_t)b * C + c], s * A1[c]);
      g2 = fmaf(dp, spos[(size_t)b * C + c], s * A2[c]);
    } else {
      g1 = A1[c];
      g2 = A2[c];
    }
    G1o[(size_t)b * C + c] = g1;
    G2o[(size_t)b * C + c] = g2;
    s1 += (double)gamma[c] * (double)g1;
    s2 += (double)gamma[c] * (double)g2;
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    coef[2 * b] = (float)s1;
    coef[2 * b + 1] = (float)s2;
  }
}

// ------------------------------------------------------------------ backward: parameter gradients (sum over batch)
// One WARP per job (lanes over the batch): the form for small job counts (a thread-per-job loop over B was latency
// bound); job space: [0, C) gamma/beta(+db2); then dw2 [C*R]; then dw1 [R*C]; then db1 [R].
// (`block` of `nblocks`: the CTA's place inside this layer's share of the grid -- the whole grid for a stand-alone
//  launch, a slice of it in gn_bwd_params_batched_kernel)
__device__ __forceinline__ void gn_bwd_params_warp_body(
    const float* __restrict__ G1, const float* __restrict__ G2, const float* __restrict__ dz2,
    const float* __restrict__ dz1, const float* __restrict__ cpos, const float* __restrict__ spos,
    const float* __restrict__ hidden, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw1, float* __restrict__ db1,
    float* __restrict__ dw2, float* __restrict__ db2, int B, int P, int C, int R, int has_se, int block, int nblocks) {
  // job space: [0, C) gamma/beta(+db2); then dw2 [C*R]; then dw1 [R*C]; then db1 [R].
  // One WARP per job, lanes stride over the batch (the per-thread serial loop over B was latency bound).
  const long long n_gb = C, n_w2 = has_se ? (long long)C * R : 0, n_w1 = n_w2, n_b1 = has_se ? R : 0;
  const long long total = n_gb + n_w2 + n_w1 + n_b1;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)block * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)nblocks * blockDim.x) >> 5;
  for (long long i = warp0; i < total; i += nwarps) {
    if (i < n_gb) {
      const int c = (int)i;
      float dg = 0.f, db = 0.f, d2 = 0.f;
      for (int b = lane; b < B; b += 32) {
        dg += G2[(size_t)b * C + c];
        db += G1[(size_t)b * C + c];
        if (has_se) d2 += dz2[(size_t)b * C + c];
      }
      dg = warp_sum(dg);
      db = warp_sum(db);
      d2 = warp_sum(d2);
      if (lane == 0) {
        dgamma[c] = dg;
        dbeta[c] = db;
        if (has_se) db2[c] = d2;
      }
    } else if (i < n_gb + n_w2) {
      const long long k = i - n_gb;
      const int c = (int)(k / R), j = (int)(k % R);
      float acc = 0.f;
      for (int b = lane; b < B; b += 32) acc = fmaf(dz2[(size_t)b * C + c], hidden[(size_t)b * R + j], acc);
      acc = warp_sum(acc);
      if (lane == 0) dw2[k] = acc;
    } else if (i < n_gb + n_w2 + n_w1) {
      const long long k = i - n_gb - n_w2;
      const int j = (int)(k / C), c = (int)(k % C);
      float acc = 0.f;
      const float gc = gamma[c], bc = beta[c];
      for (int b = lane; b < B; b += 32) {
        const float pool = (gc * spos[(size_t)b * C + c] + bc * cpos[(size_t)b * C + c]) / (float)P;
        acc = fmaf(dz1[(size_t)b * R + j], pool, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) dw1[k] = acc;
    } else {
      const int j = (int)(i - n_gb - n_w2 - n_w1);
      float acc = 0.f;
      for (int b = lane; b < B; b += 32) acc += dz1[(size_t)b * R + j];
      acc = warp_sum(acc);
      if (lane == 0) db1[j] = acc;
    }
  }
}

// The two SE weight gradients are small GEMMs over the batch -- dw2[c][j] = sum_b dz2[b][c] hidden[b][j] and
// dw1[j][c] = sum_b dz1[b][j] pool[b][c] -- so with many (c, j) pairs a shared-memory tile beats one job per warp
// (C = 512, R = 64: 66 k jobs took 22.7 us thread-per-job, 29 us warp-per-job).  CTA = 32 channels x 8 hidden units
// (grid C/32 x R/8: enough CTAs that each one is a single round of loads, 32 FMAs per thread and a store); lane =
// channel, warp = hidden unit; batch in chunks of 32 through shared memory.  Fixed summation order (b ascending).
__device__ __forceinline__ void gn_bwd_params_tiled_body(
    const float* __restrict__ G1, const float* __restrict__ G2, const float* __restrict__ dz2,
    const float* __restrict__ dz1, const float* __restrict__ cpos, const float* __restrict__ spos,
    const float* __restrict__ hidden, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw1, float* __restrict__ db1,
    float* __restrict__ dw2, float* __restrict__ db2, int B, int P, int C, int R, int bx, int by) {
  __shared__ float dz2s[32][33], pools[32][33];
  __shared__ float hids[32][9], dz1s[32][9];
  const int c0 = bx * 32, j0 = by * 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + lane, j = j0 + warp;
  const bool cok = c < C, jok = j < R;
  const float gc = cok ? gamma[c] : 0.f, bc = cok ? beta[c] : 0.f, invP = 1.f / (float)P;
  const bool sums = by == 0;                  // these CTAs also produce dgamma / dbeta / db2 for their channels
  float acc2 = 0.f, acc1 = 0.f, dg = 0.f, db = 0.f, d2 = 0.f;
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int nb = min(32, B - b0);
    __syncthreads();
    {
      // one round of independent loads: warp w stages samples w, w+8, w+16, w+24 (lane = channel), and the
      // 32 x 8 hidden / dz1 blocks (thread = (sample, unit))
      float v2[4], vs[4], vc[4], g1[4], g2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = warp + 8 * u;
        const bool ok = cok && r < nb;
        const size_t o = (size_t)(b0 + r) * C + c;
        v2[u] = ok ? dz2[o] : 0.f;
        vs[u] = ok ? spos[o] : 0.f;
        vc[u] = ok ? cpos[o] : 0.f;
        g1[u] = (ok && sums) ? G1[o] : 0.f;
        g2[u] = (ok && sums) ? G2[o] : 0.f;
      }
      const int hr = threadIdx.x >> 3, hj = threadIdx.x & 7;
      const bool hok = hr < nb && j0 + hj < R;
      const float hv = hok ? hidden[(size_t)(b0 + hr) * R + j0 + hj] : 0.f;
      const float zv = hok ? dz1[(size_t)(b0 + hr) * R + j0 + hj] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dz2s[warp + 8 * u][lane] = v2[u];
        pools[warp + 8 * u][lane] = (gc * vs[u] + bc * vc[u]) * invP;
        dg += g2[u];                              // warp w holds the partial over its four samples
        db += g1[u];
        d2 += v2[u];
      }
      hids[hr][hj] = hv;
      dz1s[hr][hj] = zv;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {                // rows >= nb are zero
      acc2 = fmaf(dz2s[r][lane], hids[r][warp], acc2);
      acc1 = fmaf(dz1s[r][warp], pools[r][lane], acc1);
    }
  }
  if (cok && jok) dw1[(size_t)j * C + c] = acc1;
  // dw2 tile [32 c][8 j]: rows of 8 contiguous floats; plus the per-channel sums (8 warp partials each)
  __syncthreads();
  float* stage = &dz2s[0][0];                     // [32][9]
  float* red = &pools[0][0];                      // [3][8][32]
  stage[lane * 9 + warp] = acc2;
  if (sums) {
    red[(0 * 8 + warp) * 32 + lane] = dg;
    red[(1 * 8 + warp) * 32 + lane] = db;
    red[(2 * 8 + warp) * 32 + lane] = d2;
  }
  __syncthreads();
  {
    const int cc = threadIdx.x >> 3, jj = threadIdx.x & 7;
    if (c0 + cc < C && j0 + jj < R) dw2[(size_t)(c0 + cc) * R + j0 + jj] = stage[cc * 9 + jj];
  }
  if (sums && warp < 3 && cok) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[(warp * 8 + w) * 32 + lane];
    (warp == 0 ? dgamma : (warp == 1 ? dbeta : db2))[c] = a;
  }
  if (bx == 0 && warp == 0) {                     // db1 for this CTA's hidden units: lanes over the batch
    for (int jj = 0; jj < 8 && j0 + jj < R; ++jj) {
      float a = 0.f;
      for (int b = lane; b < B; b += 32) a += dz1[(size_t)b * R + j0 + jj];
      a = warp_sum(a);
      if (lane == 0) db1[j0 + jj] = a;
    }
  }
}

__global__ void __launch_bounds__(256)
    gn_bwd_params_warp_kernel(const float* __restrict__ G1, const float* __restrict__ G2, const float* __restrict__ dz2,
                              const float* __restrict__ dz1, const float* __restrict__ cpos, const float* __restrict__ spos,
                              const float* __restrict__ hidden, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* __restrict__ dgamma, float* __restrict__ dbeta,
                              float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                              float* __restrict__ db2, int B, int P, int C, int R, int has_se) {
  gn_bwd_params_warp_body(G1, G2, dz2, dz1, cpos, spos, hidden, gamma, beta, dgamma, dbeta, dw1, db1, dw2, db2, B, P, C, R,
                          has_se, (int)blockIdx.x, (int)gridDim.x);
}
__global__ void __launch_bounds__(256)
    gn_bwd_params_tiled_kernel(const float* __restrict__ G1, const float* __restrict__ G2, const float* __restrict__ dz2,
                               const float* __restrict__ dz1, const float* __restrict__ cpos,
                               const float* __restrict__ spos, const float* __restrict__ hidden,
                               const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ dgamma,
                               float* __restrict__ dbeta, float* __restrict__ dw1, float* __restrict__ db1,
                               float* __restrict__ dw2, float* __restrict__ db2, int B, int P, int C, int R) {
  gn_bwd_params_tiled_body(G1, G2, dz2, dz1, cpos, spos, hidden, gamma, beta, dgamma, dbeta, dw1, db1, dw2, db2, B, P, C, R,
                           (int)blockIdx.x, (int)blockIdx.y);
}

// The parameter gradients of SEVERAL blocks in one launch.  Each block's reduction over the batch is a few
// microseconds of work behind a launch, and nothing downstream of it runs before the optimizer: issued per block they
// are eight serialised small kernels on the backward's critical path (~35 us of a 1.4 ms step).  The caller collects one
// job per block (bvae_gn_relu_se_bwd_deferred) and flushes them together when backward has been enqueued.
struct GnParamJob {
  const float *G1, *G2, *dz2, *dz1, *cpos, *spos, *hidden, *gamma, *beta;
  float *dgamma, *dbeta, *dw1, *db1, *dw2, *db2;
  int B, P, C, R, has_se;
  int tiled;       // 1: tiled kernel body, grid (C/32) x (R/8); 0: warp-per-job body over `blocks` CTAs
  int blocks;
  int pad_;
};
static constexpr int kGnMaxJobs = 16;
struct GnParamBatch {
  GnParamJob job[kGnMaxJobs];
  int first[kGnMaxJobs + 1];   // first CTA of job i; first[n] = grid size
  int n;
};
__global__ void __launch_bounds__(256) gn_bwd_params_batched_kernel(const __grid_constant__ GnParamBatch bt) {
  int l = 0;
  while (l + 1 < bt.n && (int)blockIdx.x >= bt.first[l + 1]) ++l;
  const GnParamJob& j = bt.job[l];
  const int lb = (int)blockIdx.x - bt.first[l];
  if (j.tiled) {
    const int gx = (j.C + 31) / 32;
    gn_bwd_params_tiled_body(j.G1, j.G2, j.dz2, j.dz1, j.cpos, j.spos, j.hidden, j.gamma, j.beta, j.dgamma, j.dbeta, j.dw1,
                             j.db1, j.dw2, j.db2, j.B, j.P, j.C, j.R, lb % gx, lb / gx);
  } else {
    gn_bwd_params_warp_body(j.G1, j.G2, j.dz2, j.dz1, j.cpos, j.spos, j.hidden, j.gamma, j.beta, j.dgamma, j.dbeta, j.dw1,
                            j.db1, j.dw2, j.db2, j.B, j.P, j.C, j.R, j.has_se, lb, j.blocks);
  }
}

// ------------------------------------------------------------------ backward: write dy
template <typename T>
__global__ void __launch_bounds__(kGnThreads)
    gn_bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ y, T* __restrict__ dy,
                        const float* __restrict__ stats, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const float* __restrict__ gate,
                        const float* __restrict__ dpool, const float* __restrict__ coef, int P, int C, int cv,
                        int rows, int pc) {
  constexpr int V = Vec16<T>::N;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  const float invM = 1.0f / ((float)P * (float)C);
  const float S1 = coef[2 * b] * invM, S2 = coef[2 * b + 1] * invM;
  const int col = threadIdx.x % cv, row = threadIdx.x / cv;
  const int p0 = chunk * pc, p1 = min(P, p0 + pc);
  // dy = rstd * (gg*gamma - mean(gg*gamma) - xhat * mean(gg*gamma*xhat)),  gg = mask ? dout*gate + dpool : 0
  //    = (mask ? dout * c1 + c2 : 0) + (y - mean) * k2 + k1
  float za[V], bt[V], c1[V], c2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float rg = rstd * gamma[col * V + i];
    za[i] = rg;
    bt[i] = beta[col * V + i];
    c1[i] = rg * (gate ? gate[(size_t)b * C + col * V + i] : 1.f);
    c2[i] = rg * (gate ? dpool[(size_t)b * C + col * V + i] : 0.f);
  }
  const float k1 = -rstd * S1, k2 = -rstd * rstd * S2;
  const size_t off = ((size_t)b * P) * C + (size_t)col * V;
#pragma unroll 2
  for (int p = p0 + row; row < rows && p < p1; p += rows) {
    Vec16<T> v, d;
    v.load(y + off + (size_t)p * C);
    d.load(dout + off + (size_t)p * C);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float vm = v.v[i] - mean;
      const float t = (fmaf(vm, za[i], bt[i]) > 0.f) ? fmaf(d.v[i], c1[i], c2[i]) : 0.f;
      d.v[i] = t + fmaf(vm, k2, k1);
    }
    d.store(dy + off + (size_t)p * C);
  }
}

// ------------------------------------------------------------------ forward: normalise + relu + rescale + bilinear x2
// Decoder blocks feed an nn.Upsample(x2, bilinear, align_corners=False) (reference: src/models/beta_vae.py:52-53): the
// activation relu(gn(y)) * gate is never written at its own resolution -- every thread owns one INPUT vector (b, h, w, c),
// activates its 3x3 neighbourhood on the fly (one FMA + max per element, coefficients of its (sample, channel) in
// registers) and writes the four output pixels it is the centre of.  Stencil as in upsample.cu:
//   out[2i] = .25 a[i-1] + .75 a[i],  out[2i+1] = .75 a[i] + .25 a[i+1]   (edge-clamped, rows then columns).
template <typename T>
__global__ void __launch_bounds__(256)
    gn_apply_up_kernel(const T* __restrict__ y, T* __restrict__ up, const float* __restrict__ stats,
                       const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ gate,
                       int H, int W, int C, int cv, int vec_per_sample) {
  constexpr int V = Vec16<T>::N;
  const int b = blockIdx.y;
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  const T* yb = y + (size_t)b * H * W * C;
  T* ub = up + (size_t)b * 4 * H * W * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < vec_per_sample; i += gridDim.x * blockDim.x) {
    const int c = i % cv;
    const int pix = i / cv;
    const int w = pix % W, h = pix / W;
    float zs[V], bs_[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float sc = gate ? gate[(size_t)b * C + c * V + k] : 1.f;
      zs[k] = rstd * gamma[c * V + k] * sc;
      bs_[k] = beta[c * V + k] * sc;
    }
    const int hs[3] = {max(h - 1, 0), h, min(h + 1, H - 1)}, ws[3] = {max(w - 1, 0), w, min(w + 1, W - 1)};
    Vec16<T> n[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int d = 0; d < 3; ++d) n[a][d].load(yb + ((size_t)hs[a] * W + ws[d]) * C + (size_t)c * V);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int k = 0; k < V; ++k) {
          // round to the storage type first: bit-identical to materialising the activation and upsampling it
          const float act = fmaxf(fmaf(n[a][d].v[k] - mean, zs[k], bs_[k]), 0.f);
          n[a][d].v[k] = Elem<T>::to_f(Elem<T>::from_f(act));
        }
#pragma unroll
    for (int dh = 0; dh < 2; ++dh) {
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const int ra = dh == 0 ? 0 : 1, rb = dh == 0 ? 1 : 2;
        const float wa = dh == 0 ? 0.25f : 0.75f, wb = dh == 0 ? 0.75f : 0.25f;
        const int ca = dw == 0 ? 0 : 1, cb = dw == 0 ? 1 : 2;
        const float va = dw == 0 ? 0.25f : 0.75f, vb = dw == 0 ? 0.75f : 0.25f;
        Vec16<T> o;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const float top = va * n[ra][ca].v[k] + vb * n[ra][cb].v[k];
          const float bot = va * n[rb][ca].v[k] + vb * n[rb][cb].v[k];
          o.v[k] = wa * top + wb * bot;
        }
        o.store(ub + ((size_t)(2 * h + dh) * (2 * W) + (2 * w + dw)) * C + (size_t)c * V);
      }
    }
  }
}

// ------------------------------------------------------------------ backward: bilinear x2 backward + per-(sample,channel) sums
// d_up [B][2H][2W][C] is the gradient w.r.t. the upsampled activation.  dout = upsample2x_bwd(d_up) (the transposed
// stencil of upsample.cu, gathered) is written once as a workspace tensor and, in the same pass, enters the
// A1 = sum dout*m and A2 = sum dout*m*xhat partial sums of gn_bwd_reduce_kernel (same chunk geometry, same order).
template <typename T>
__global__ void __launch_bounds__(kGnThreads)
    gn_bwd_reduce_up_kernel(const T* __restrict__ dup, const T* __restrict__ y, T* __restrict__ dout,
                            const float* __restrict__ stats, const float* __restrict__ gamma,
                            const float* __restrict__ beta, float* __restrict__ a1part, float* __restrict__ a2part,
                            int H, int W, int C, int cv, int rows, int pc) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sm[];
  const int P = H * W, H2 = 2 * H, W2 = 2 * W;
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  const int col = threadIdx.x % cv, row = threadIdx.x / cv;
  const int p0 = chunk * pc, p1 = min(P, p0 + pc);
  float za[V], bt[V], a1[V], a2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    za[i] = rstd * gamma[col * V + i];
    bt[i] = beta[col * V + i];
    a1[i] = 0.f;
    a2[i] = 0.f;
  }
  const size_t off = ((size_t)b * P) * C + (size_t)col * V;
  const T* db = dup + (size_t)b * H2 * W2 * C + (size_t)col * V;
  const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  for (int p = p0 + row; row < rows && p < p1; p += rows) {
    const int h = p / W, w = p % W;
    const int hs[4] = {max(2 * h - 1, 0), 2 * h, 2 * h + 1, min(2 * h + 2, H2 - 1)};
    const int ws[4] = {max(2 * w - 1, 0), 2 * w, 2 * w + 1, min(2 * w + 2, W2 - 1)};
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
        Vec16<T> g;
        g.load(db + ((size_t)hs[a] * W2 + ws[d]) * C);
#pragma unroll
        for (int k = 0; k < V; ++k) rowacc[k] = fmaf(wt[d], g.v[k], rowacc[k]);
      }
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = fmaf(wt[a], rowacc[k], acc[k]);
    }
    Vec16<T> v, d;
    v.load(y + off + (size_t)p * C);
#pragma unroll
    for (int k = 0; k < V; ++k) d.v[k] = Elem<T>::to_f(Elem<T>::from_f(acc[k]));   // what the apply pass will read back
    d.store(dout + off + (size_t)p * C);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float vm = v.v[i] - mean;
      const float dm = (fmaf(vm, za[i], bt[i]) > 0.f) ? d.v[i] : 0.f;
      a1[i] += dm;
      a2[i] = fmaf(dm, vm, a2[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) a2[i] *= rstd;
  float* s0 = sm;
  float* s1 = sm + rows * C;
  if (row < rows) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s0[row * C + col * V + i] = a1[i];
      s1[row * C + col * V + i] = a2[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0, q = 0.0;  // fixed-order, double: these sums cancel heavily downstream
    for (int r = 0; r < rows; ++r) {
      a += (double)s0[r * C + c];
      q += (double)s1[r * C + c];
    }
    const size_t o = ((size_t)b * nchunks + chunk) * C + c;
    a1part[o] = (float)a;
    a2part[o] = (float)q;
  }
}

// `job` non-null: describe the launch instead of issuing it (deferred, see gn_bwd_params_batched_kernel)
static int launch_gn_params(const float* G1, const float* G2, const float* dz2, const float* dz1, const float* cpos,
                            const float* spos, const float* hidden, const float* gamma, const float* beta,
                            float* dgamma, float* dbeta, float* dw1, float* db1, float* dw2, float* db2, int B, int P,
                            int C, int R, bool has_se, cudaStream_t st, GnParamJob* job = nullptr) {
  const long long jobs = (long long)C + (has_se ? 2LL * C * R + R : 0);
  const bool tiled = has_se && jobs > 4096;
  int blocks;
  if (tiled) {
    blocks = ((C + 31) / 32) * ((R + 7) / 8);
  } else {
    blocks = (int)((jobs + 7) / 8);  // one warp per job, 8 warps per CTA
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  }
  if (job) {
    *job = GnParamJob{G1, G2, dz2, dz1, cpos, spos, hidden, gamma, beta, dgamma, dbeta, dw1, db1, dw2, db2,
                      B, P, C, R, has_se ? 1 : 0, tiled ? 1 : 0, blocks, 0};
    return BVAE_OK;
  }
  if (tiled) {
    gn_bwd_params_tiled_kernel<<<dim3((C + 31) / 32, (R + 7) / 8), 256, 0, st>>>(G1, G2, dz2, dz1, cpos, spos, hidden, gamma, beta, dgamma,
                                                              dbeta, dw1, db1, dw2, db2, B, P, C, R);
    BVAE_LAUNCH_CHECK("gn_bwd_params_tiled");
    return BVAE_OK;
  }
  gn_bwd_params_warp_kernel<<<blocks, 256, 0, st>>>(G1, G2, dz2, dz1, cpos, spos, hidden, gamma, beta, dgamma, dbeta, dw1,
                                                    db1, dw2, db2, B, P, C, R, has_se ? 1 : 0);
  BVAE_LAUNCH_CHECK("gn_bwd_params");
  return BVAE_OK;
}

// ------------------------------------------------------------------ workspace layout
struct GnWs {
  double* spart;  // [B][nchunks][2]
  float* p0;      // [B][nchunks][C]
  float* p1;      // [B][nchunks][C]
  float* G1;      // [B][C]
  float* G2;      // [B][C]
  float* dz2;     // [B][C]
  float* dpool;   // [B][C]
  float* dz1;     // [B][Rmax=C]
  float* coef;    // [B][2]
  void* dout;     // [B][P][C] elements (fused-upsample backward only)
  size_t bytes;
};

static GnWs gn_ws_carve(void* base, int B, int P, int C, size_t dout_bytes = 0) {
  const int nch = gn_max_chunks(B, P, C);
  GnWs w;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.spart = reinterpret_cast<double*>(take((size_t)B * nch * 2 * sizeof(double)));
  w.p0 = reinterpret_cast<float*>(take((size_t)B * nch * C * sizeof(float)));
  w.p1 = reinterpret_cast<float*>(take((size_t)B * nch * C * sizeof(float)));
  w.G1 = reinterpret_cast<float*>(take((size_t)B * C * sizeof(float)));
  w.G2 = reinterpret_cast<float*>(take((size_t)B * C * sizeof(float)));
  w.dz2 = reinterpret_cast<float*>(take((size_t)B * C * sizeof(float)));
  w.dpool = reinterpret_cast<float*>(take((size_t)B * C * sizeof(float)));
  w.dz1 = reinterpret_cast<float*>(take((size_t)B * C * sizeof(float)));
  w.coef = reinterpret_cast<float*>(take((size_t)B * 2 * sizeof(float)));
  w.dout = dout_bytes ? take(dout_bytes) : nullptr;
  w.bytes = off;
  return w;
}

// Which kernels serve a shape.  Measured on B200 (profiles/r2_gn_paths.txt): every dependent phase of these latency-bound
// ops costs ~1 us, so for samples that fit a cluster's shared memory (<= 1 MB) the ONE-launch cluster-resident kernels
// (gn_se_cluster.cu, ~8 phases) beat the streaming passes (4 launches x ~3 phases) even with the bilinear x2 fused into
// the latter: 13-17 us vs 22-32 us forward at batch 32.  Larger samples (dec3 of the default config: 2 MB) stream.
// BVAE_GN_CLUSTER=0 / 1 forces the streaming / cluster kernels (the tests run both).
static int gn_cluster_pref() {
  const char* e = getenv("BVAE_GN_CLUSTER");
  return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0);
}
static bool gn_want_cluster(bool /*exact_fp32*/) { return gn_cluster_pref() != 0; }

// up != 0: `out` is the bilinear x2 upsample of the activation, [B][2H][2W][C].
// stat_part: optional [B][n_stat_part][2] doubles (sum, sum of squares) produced by the convolution's epilogue.
template <typename T>
static int gn_fwd_impl(const T* y, T* out, const float* gamma, const float* beta, const float* w1, const float* b1,
                       const float* w2, const float* b2, int B, int H, int W, int C, int R, int up, float* stats,
                       float* cpos, float* spos, float* hidden, float* gate, const double* stat_part, int n_stat_part,
                       void* ws, cudaStream_t st) {
  const int P = H * W;
  GnGeom g;
  BVAE_REQUIRE(gn_geom(B, P, C, Vec16<T>::N, &g), BVAE_ERR_UNSUPPORTED,
               "gn_relu_se: C=%d must be a multiple of %d with C/%d <= 256", C, Vec16<T>::N, Vec16<T>::N);
  GnWs w = gn_ws_carve(ws, B, P, C);
  if (!up && !stat_part && gn_want_cluster(sizeof(T) == 4)) {
    bool handled = false;   // one cluster-resident kernel when the sample fits a cluster's shared memory
    const int rc = gn_cluster_fwd<T>(y, out, gamma, beta, w1, b1, w2, b2, B, P, C, R, stats, cpos, spos, hidden, gate, st,
                                     &handled);
    if (rc != BVAE_OK || handled) return rc;
  }
  dim3 grid(g.nchunks, B);
  const double* sp = stat_part;
  int nsp = n_stat_part;
  if (!sp) {
    gn_stats_kernel<T><<<grid, g.threads, 0, st>>>(y, w.spart, P, C, g.cv, g.rows, g.pc);
    BVAE_LAUNCH_CHECK("gn_stats");
    sp = w.spart;
    nsp = g.nchunks;
  }
  const bool has_se = w1 != nullptr;
  if (has_se) {
    const size_t sm = (size_t)2 * g.rows * C * sizeof(float);
    gn_pool_kernel<T><<<grid, g.threads, sm, st>>>(y, sp, nsp, gamma, beta, w.p0, w.p1, stats, P, C, g.cv, g.rows, g.pc);
    BVAE_LAUNCH_CHECK("gn_pool");
    se_excite_kernel<<<B, kExThreads, (size_t)(C + R + 4) * sizeof(float), st>>>(w.p0, w.p1, g.nchunks, gamma, beta, w1, b1, w2,
                                                                      b2, cpos, spos, hidden, gate, P, C, R);
    BVAE_LAUNCH_CHECK("se_excite");
  } else {
    gn_finalize_stats_kernel<<<cdiv(B, 8), 256, 0, st>>>(sp, nsp, (double)P * C, stats, B);
    BVAE_LAUNCH_CHECK("gn_finalize_stats");
  }
  if (up) {
    const int vps = P * g.cv;
    int bx = cdiv(vps, 256);
    const int cap = cdiv(16 * num_sms(), B);
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    gn_apply_up_kernel<T><<<dim3(bx, B), 256, 0, st>>>(y, out, stats, gamma, beta, has_se ? gate : nullptr, H, W, C, g.cv,
                                                        vps);
    BVAE_LAUNCH_CHECK("gn_apply_up");
  } else {
    gn_scale_kernel<T><<<grid, g.threads, 0, st>>>(y, out, stats, gamma, beta, has_se ? gate : nullptr, P, C, g.cv, g.rows,
                                                   g.pc);
    BVAE_LAUNCH_CHECK("gn_scale");
  }
  return BVAE_OK;
}

// up != 0: `dout` is the gradient w.r.t. the UPSAMPLED activation, [B][2H][2W][C]; ws then also holds dout at [B][H][W][C].
template <typename T>
static int gn_bwd_impl(const T* dout, const T* y, T* dy, const float* gamma, const float* beta, const float* w1,
                       const float* w2, int B, int H, int W, int C, int R, int up, const float* stats, const float* cpos,
                       const float* spos, const float* hidden, const float* gate, float* dgamma, float* dbeta,
                       float* dw1, float* db1, float* dw2, float* db2, void* ws, cudaStream_t st, GnParamJob* job) {
  const int P = H * W;
  GnGeom g;
  BVAE_REQUIRE(gn_geom(B, P, C, Vec16<T>::N, &g), BVAE_ERR_UNSUPPORTED,
               "gn_relu_se: C=%d must be a multiple of %d with C/%d <= 256", C, Vec16<T>::N, Vec16<T>::N);
  GnWs w = gn_ws_carve(ws, B, P, C, up ? (size_t)B * P * C * sizeof(T) : 0);
  const bool has_se = w1 != nullptr;
  if (!up && gn_want_cluster(sizeof(T) == 4)) {
    bool handled = false;
    const int rc = gn_cluster_bwd<T>(dout, y, dy, gamma, beta, w1, w2, B, P, C, R, stats, cpos, spos, hidden, gate, w.G1,
                                     w.G2, w.dz2, w.dz1, st, &handled);
    if (rc != BVAE_OK) return rc;
    if (handled)
      return launch_gn_params(w.G1, w.G2, w.dz2, w.dz1, cpos, spos, hidden, gamma, beta, dgamma, dbeta, dw1, db1, dw2, db2,
                              B, P, C, R, has_se, st, job);
  }
  dim3 grid(g.nchunks, B);
  const size_t sm = (size_t)2 * g.rows * C * sizeof(float);
  const T* d_lo = dout;
  if (up) {
    T* dlo = reinterpret_cast<T*>(w.dout);
    gn_bwd_reduce_up_kernel<T><<<grid, g.threads, sm, st>>>(dout, y, dlo, stats, gamma, beta, w.p0, w.p1, H, W, C, g.cv,
                                                            g.rows, g.pc);
    BVAE_LAUNCH_CHECK("gn_bwd_reduce_up");
    d_lo = dlo;
  } else {
    gn_bwd_reduce_kernel<T><<<grid, g.threads, sm, st>>>(dout, y, stats, gamma, beta, w.p0, w.p1, P, C, g.cv, g.rows,
                                                         g.pc);
    BVAE_LAUNCH_CHECK("gn_bwd_reduce");
  }
  gn_bwd_mid_kernel<<<B, kExThreads, (size_t)(3 * C + (has_se ? R : 0)) * sizeof(float), st>>>(
      w.p0, w.p1, g.nchunks, gamma, beta, w1, w2, cpos, spos, hidden, gate, w.G1, w.G2, w.dz2, w.dz1, w.dpool, w.coef,
      P, C, R);
  BVAE_LAUNCH_CHECK("gn_bwd_mid");
  gn_bwd_apply_kernel<T><<<grid, g.threads, 0, st>>>(d_lo, y, dy, stats, gamma, beta, has_se ? gate : nullptr,
                                                     w.dpool, w.coef, P, C, g.cv, g.rows, g.pc);
  BVAE_LAUNCH_CHECK("gn_bwd_apply");
  // parameter gradients last: they are off the critical path of the data gradient
  return launch_gn_params(w.G1, w.G2, w.dz2, w.dz1, cpos, spos, hidden, gamma, beta, dgamma, dbeta, dw1, db1, dw2, db2, B,
                          P, C, R, has_se, st, job);
}

}  // namespace bvae

using namespace bvae;

extern "C" size_t bvae_gn_relu_se_workspace(int B, int P, int C) {
  if (B <= 0 || P <= 0 || C <= 0) return 0;
  return gn_ws_carve(nullptr, B, P, C).bytes;
}

extern "C" size_t bvae_gn_relu_se_up_workspace(int B, int H, int W, int C) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0) return 0;
  return gn_ws_carve(nullptr, B, H * W, C, (size_t)B * H * W * C * sizeof(float)).bytes;
}

static int gn_fwd_entry(const void* y, void* out, const float* gamma, const float* beta, const float* w1, const float* b1,
                        const float* w2, const float* b2, int B, int H, int W, int C, int R, int up, float* stats,
                        float* cpos, float* spos, float* hidden, float* gate, const double* stat_part, int n_stat_part,
                        void* ws, size_t ws_bytes, int dtype, bvae_stream_t stream) {
  BVAE_REQUIRE(y && out && gamma && beta && stats && ws, BVAE_ERR_ARG, "gn_relu_se_fwd: null pointer");
  BVAE_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, BVAE_ERR_ARG, "gn_relu_se_fwd: empty shape B=%d P=%d C=%d", B, H * W, C);
  BVAE_REQUIRE(stat_part == nullptr || n_stat_part > 0, BVAE_ERR_ARG, "gn_relu_se_fwd: stat_part without a count");
  if (w1) {
    BVAE_REQUIRE(b1 && w2 && b2 && cpos && spos && hidden && gate && R > 0, BVAE_ERR_ARG,
                 "gn_relu_se_fwd: SE requested but SE buffers missing");
    BVAE_REQUIRE((size_t)(C + R) * 4 <= 48 * 1024, BVAE_ERR_UNSUPPORTED, "gn_relu_se: C+R too large");
  }
  const size_t need = up ? bvae_gn_relu_se_up_workspace(B, H, W, C) : bvae_gn_relu_se_workspace(B, H * W, C);
  BVAE_REQUIRE(ws_bytes >= need, BVAE_ERR_WORKSPACE, "gn_relu_se_fwd: workspace %zu < %zu", ws_bytes, need);
  cudaStream_t st = as_stream(stream);
  if (dtype == BVAE_F32)
    return gn_fwd_impl<float>((const float*)y, (float*)out, gamma, beta, w1, b1, w2, b2, B, H, W, C, R, up, stats, cpos,
                              spos, hidden, gate, stat_part, n_stat_part, ws, st);
  if (dtype == BVAE_BF16)
    return gn_fwd_impl<__nv_bfloat16>((const __nv_bfloat16*)y, (__nv_bfloat16*)out, gamma, beta, w1, b1, w2, b2, B, H, W,
                                      C, R, up, stats, cpos, spos, hidden, gate, stat_part, n_stat_part, ws, st);
  BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "gn_relu_se_fwd: dtype %d", dtype);
}

static int gn_bwd_entry(const void* dout, const void* y, void* dy, const float* gamma, const float* beta, const float* w1,
                        const float* w2, int B, int H, int W, int C, int R, int up, const float* stats, const float* cpos,
                        const float* spos, const float* hidden, const float* gate, float* dgamma, float* dbeta,
                        float* dw1, float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes, int dtype,
                        bvae_stream_t stream, GnParamJob* job = nullptr) {
  BVAE_REQUIRE(dout && y && dy && gamma && beta && stats && dgamma && dbeta && ws, BVAE_ERR_ARG,
               "gn_relu_se_bwd: null pointer");
  BVAE_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, BVAE_ERR_ARG, "gn_relu_se_bwd: empty shape");
  if (w1) {
    BVAE_REQUIRE(w2 && cpos && spos && hidden && gate && dw1 && db1 && dw2 && db2 && R > 0, BVAE_ERR_ARG,
                 "gn_relu_se_bwd: SE requested but SE buffers missing");
    BVAE_REQUIRE((size_t)(3 * C + R) * 4 <= 48 * 1024 && R <= kExThreads, BVAE_ERR_UNSUPPORTED, "gn_relu_se: C / R too large");
  }
  const size_t need = up ? bvae_gn_relu_se_up_workspace(B, H, W, C) : bvae_gn_relu_se_workspace(B, H * W, C);
  BVAE_REQUIRE(ws_bytes >= need, BVAE_ERR_WORKSPACE, "gn_relu_se_bwd: workspace small");
  cudaStream_t st = as_stream(stream);
  if (dtype == BVAE_F32)
    return gn_bwd_impl<float>((const float*)dout, (const float*)y, (float*)dy, gamma, beta, w1, w2, B, H, W, C, R, up,
                              stats, cpos, spos, hidden, gate, dgamma, dbeta, dw1, db1, dw2, db2, ws, st, job);
  if (dtype == BVAE_BF16)
    return gn_bwd_impl<__nv_bfloat16>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)y, (__nv_bfloat16*)dy, gamma,
                                      beta, w1, w2, B, H, W, C, R, up, stats, cpos, spos, hidden, gate, dgamma, dbeta,
                                      dw1, db1, dw2, db2, ws, st, job);
  BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "gn_relu_se_bwd: dtype %d", dtype);
}

extern "C" int bvae_gn_relu_se_fwd(const void* y, void* out, const float* gamma, const float* beta, const float* w1,
                                   const float* b1, const float* w2, const float* b2, int B, int P, int C, int R,
                                   float* stats, float* cpos, float* spos, float* hidden, float* gate,
                                   const double* stat_part, int n_stat_part, void* ws, size_t ws_bytes, int dtype,
                                   bvae_stream_t stream) {
  return gn_fwd_entry(y, out, gamma, beta, w1, b1, w2, b2, B, P, 1, C, R, 0, stats, cpos, spos, hidden, gate, stat_part,
                      n_stat_part, ws, ws_bytes, dtype, stream);
}

extern "C" int bvae_gn_relu_se_up_fwd(const void* y, void* out_up, const float* gamma, const float* beta, const float* w1,
                                      const float* b1, const float* w2, const float* b2, int B, int H, int W, int C,
                                      int R, float* stats, float* cpos, float* spos, float* hidden, float* gate,
                                      const double* stat_part, int n_stat_part, void* ws, size_t ws_bytes, int dtype,
                                      bvae_stream_t stream) {
  return gn_fwd_entry(y, out_up, gamma, beta, w1, b1, w2, b2, B, H, W, C, R, 1, stats, cpos, spos, hidden, gate, stat_part,
                      n_stat_part, ws, ws_bytes, dtype, stream);
}

extern "C" int bvae_gn_relu_se_bwd(const void* dout, const void* y, void* dy, const float* gamma, const float* beta,
                                   const float* w1, const float* b1, const float* w2, const float* b2, int B, int P,
                                   int C, int R, const float* stats, const float* cpos, const float* spos,
                                   const float* hidden, const float* gate, float* dgamma, float* dbeta, float* dw1,
                                   float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes, int dtype,
                                   bvae_stream_t stream) {
  (void)b1;
  (void)b2;
  return gn_bwd_entry(dout, y, dy, gamma, beta, w1, w2, B, P, 1, C, R, 0, stats, cpos, spos, hidden, gate, dgamma, dbeta,
                      dw1, db1, dw2, db2, ws, ws_bytes, dtype, stream);
}

extern "C" int bvae_gn_relu_se_up_bwd(const void* d_up, const void* y, void* dy, const float* gamma, const float* beta,
                                      const float* w1, const float* b1, const float* w2, const float* b2, int B, int H,
                                      int W, int C, int R, const float* stats, const float* cpos, const float* spos,
                                      const float* hidden, const float* gate, float* dgamma, float* dbeta, float* dw1,
                                      float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes, int dtype,
                                      bvae_stream_t stream) {
  (void)b1;
  (void)b2;
  return gn_bwd_entry(d_up, y, dy, gamma, beta, w1, w2, B, H, W, C, R, 1, stats, cpos, spos, hidden, gate, dgamma, dbeta,
                      dw1, db1, dw2, db2, ws, ws_bytes, dtype, stream);
}

extern "C" size_t bvae_gn_params_job_bytes(void) { return sizeof(GnParamJob); }

/* bvae_gn_relu_se_bwd / _up_bwd (up != 0: dout is [B][2H][2W][C]) WITHOUT the reduction of the parameter gradients over
 * the batch: dy is produced, the per-sample terms stay in `ws`, and `job_out` (HOST memory, bvae_gn_params_job_bytes())
 * receives the description of the missing launch.  `ws` and every pointer passed here must stay valid until the job has
 * gone through bvae_gn_bwd_params_batched on the same stream. */
extern "C" int bvae_gn_relu_se_bwd_deferred(const void* dout, const void* y, void* dy, const float* gamma, const float* beta,
                                            const float* w1, const float* w2, int B, int H, int W, int C, int R, int up,
                                            const float* stats, const float* cpos, const float* spos,
                                            const float* hidden, const float* gate, float* dgamma, float* dbeta,
                                            float* dw1, float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes,
                                            int dtype, void* job_out, bvae_stream_t stream) {
  BVAE_REQUIRE(job_out, BVAE_ERR_ARG, "gn_relu_se_bwd_deferred: null job_out");
  return gn_bwd_entry(dout, y, dy, gamma, beta, w1, w2, B, H, W, C, R, up ? 1 : 0, stats, cpos, spos, hidden, gate, dgamma,
                      dbeta, dw1, db1, dw2, db2, ws, ws_bytes, dtype, stream, reinterpret_cast<GnParamJob*>(job_out));
}

/* One launch for the parameter gradients of `njobs` deferred blocks (`jobs`: HOST array of job records, back to back). */
extern "C" int bvae_gn_bwd_params_batched(const void* jobs, int njobs, bvae_stream_t stream) {
  BVAE_REQUIRE(jobs && njobs > 0, BVAE_ERR_ARG, "gn_bwd_params_batched: no jobs");
  const GnParamJob* src = reinterpret_cast<const GnParamJob*>(jobs);
  cudaStream_t st = as_stream(stream);
  for (int i0 = 0; i0 < njobs; i0 += kGnMaxJobs) {
    GnParamBatch bt;
    memset(&bt, 0, sizeof(bt));
    bt.n = njobs - i0 < kGnMaxJobs ? njobs - i0 : kGnMaxJobs;
    int total = 0;
    for (int i = 0; i < bt.n; ++i) {
      bt.job[i] = src[i0 + i];
      BVAE_REQUIRE(bt.job[i].blocks > 0 && bt.job[i].G1 && bt.job[i].dgamma, BVAE_ERR_ARG, "gn_bwd_params_batched: bad job %d",
                   i0 + i);
      bt.first[i] = total;
      total += bt.job[i].blocks;
    }
    bt.first[bt.n] = total;
    gn_bwd_params_batched_kernel<<<total, 256, 0, st>>>(bt);
    BVAE_LAUNCH_CHECK("gn_bwd_params_batched");
  }
  return BVAE_OK;
}

/* 1 when bvae_gn_relu_se_fwd would run the single-launch cluster-resident kernel for this shape (the caller then keeps
 * the bilinear x2 as its own pass); 0 when the streaming passes run (the fused *_up_* entry points then save a pass). */
extern "C" int bvae_gn_relu_se_prefers_unfused(int B, int P, int C, int R, int dtype) {
  if (gn_cluster_pref() == 0) return 0;
  return gn_cluster_eligible(B, P, C, R, dtype == BVAE_BF16 ? 2 : 4) ? 1 : 0;
}
