// Latent heads (fc_mu, fc_logvar) + clamp + reparameterisation + KL + KL-side objective,
// and fc_dec; forward and backward.
//
// Replaces, from the reference: the two nn.Linear heads and torch.clamp in BetaVAE.encode
// (src/models/beta_vae.py:137-145), reparameterize (:147-150), the KL / free-bits / capacity /
// latent-reg arithmetic of BetaVAE.loss (:219-253), fc_dec and the latent clamp in decode
// (:153-161), plus everything autograd derives for them.
//
// Two GEMM shapes cover all six products.  F = C*P (32768 for the default config) is the
// long axis, the other two (batch, 2L) are short:
//   reduce-F ("tall"):  out[m][n] = sum_f X[m][f] * W[n][f]       heads fwd, fc_dec dz
//       split over F into chunks; fp32 partials [chunk][m][n]; a finalize kernel sums the
//       chunks in a fixed order (deterministic) and applies the fused epilogue.
//   wide-F:             out[m][f] = sum_k A[k or m][..] * Bm[k][f]  heads dh / dW, fc_dec fwd / dW
//       K is tiny (L, 2L or B); every CTA streams a 128-feature slab of Bm exactly once.
// Activations use NHWC feature order f' = p*C + c; master parameters / gradients use the
// reference's NCHW order f = c*P + p.  The permutation is folded into the pack kernels
// (core.cu) and into the store addressing of the wide-F kernel.
#include "common.cuh"

namespace bvae {

// ------------------------------------------------------------------------------------------
// reduce-F
// ------------------------------------------------------------------------------------------
static constexpr int RF_CHUNK = 128;  // features per CTA (one fp32 partial [M][N] per chunk)
static constexpr int RF_SUB = 64;     // features staged in shared memory per pass
static constexpr int RF_MT = 32;      // rows (batch) per CTA
static constexpr int RF_NT = 128;     // output columns per pass
static constexpr int RF_LDX = RF_MT + 4, RF_LDW = RF_NT + 4;

// CTA tile 32 (m) x 128 (n) x 128 (f), thread tile 4 x 4.  Both operand tiles are stored TRANSPOSED in shared memory
// ([f][m], [f][n]) so that one feature step is one broadcast 16-byte read of X and one conflict-free 16-byte read of W
// per thread for 16 FMAs (the earlier [row][f] layout needed five 16-byte reads per 16 FMAs and ran at the
// shared-memory pipe's limit).  Lanes map to rows when loading, so the transposed stores are conflict free.
template <typename T>
__global__ void __launch_bounds__(256) fc_reduce_f_kernel(const T* __restrict__ X, const T* __restrict__ Wm,
                                                          float* __restrict__ part, int M, int N, int F) {
  __shared__ __align__(16) float Xs[RF_SUB][RF_LDX];
  __shared__ __align__(16) float Ws[RF_SUB][RF_LDW];
  const int chunk = blockIdx.x, m0 = blockIdx.y * RF_MT;
  const int f0 = chunk * RF_CHUNK;
  const int tid = threadIdx.x;
  constexpr int V = Vec16<T>::N;
  constexpr int QV = RF_SUB / V;           // 16-byte vectors per staged row
  const int tm = tid >> 5, tn = tid & 31;  // rows tm*4..+3, columns tn*4..+3 of the pass
  for (int n0 = 0; n0 < N; n0 += RF_NT) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int fs = f0; fs < f0 + RF_CHUNK && fs < F; fs += RF_SUB) {
      const bool vec_ok = (F % V == 0) && (fs + RF_SUB <= F);
      __syncthreads();
      if (vec_ok) {
        for (int i = tid; i < RF_MT * QV; i += 256) {        // lane = row, warp-uniform vector index
          const int r = i % RF_MT, q = i / RF_MT;
          Vec16<T> v;
          if (m0 + r < M) {
            v.load(X + (size_t)(m0 + r) * F + fs + q * V);
          } else {
#pragma unroll
            for (int e = 0; e < V; ++e) v.v[e] = 0.f;
          }
#pragma unroll
          for (int e = 0; e < V; ++e) Xs[q * V + e][r] = v.v[e];
        }
        for (int i = tid; i < RF_NT * QV; i += 256) {
          const int r = i % RF_NT, q = i / RF_NT;
          Vec16<T> v;
          if (n0 + r < N) {
            v.load(Wm + (size_t)(n0 + r) * F + fs + q * V);
          } else {
#pragma unroll
            for (int e = 0; e < V; ++e) v.v[e] = 0.f;
          }
#pragma unroll
          for (int e = 0; e < V; ++e) Ws[q * V + e][r] = v.v[e];
        }
      } else {
        for (int i = tid; i < RF_MT * RF_SUB; i += 256) {
          const int r = i % RF_MT, f = i / RF_MT;
          Xs[f][r] = (m0 + r < M && fs + f < F) ? Elem<T>::to_f(X[(size_t)(m0 + r) * F + fs + f]) : 0.f;
        }
        for (int i = tid; i < RF_NT * RF_SUB; i += 256) {
          const int r = i % RF_NT, f = i / RF_NT;
          Ws[f][r] = (n0 + r < N && fs + f < F) ? Elem<T>::to_f(Wm[(size_t)(n0 + r) * F + fs + f]) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll 8
      for (int f = 0; f < RF_SUB; ++f) {
        const float4 a = *reinterpret_cast<const float4*>(&Xs[f][tm * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Ws[f][tn * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + tm * 4 + i, n = n0 + tn * 4;
      if (m >= M) continue;
      float* o = part + ((size_t)chunk * M + m) * N + n;
      if ((N & 3) == 0 && n + 3 < N) {
        *reinterpret_cast<float4*>(o) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) o[j] = acc[i][j];
      }
    }
  }
}

// ---- bf16: the same product on the warp-level tensor-core path -----------------------------------------------
// These GEMMs have M = batch (32) and stream their long operand exactly once, so they are bound by HBM, not by math;
// what the CUDA-core kernel above cannot do is stay out of the way of the loads (16 FMA issue slots per 8 bytes of
// W).  mma.sync m16n8k16 needs one issue slot per 4096 MACs, which leaves the kernel a single load phase: every
// thread issues all of its 16-byte loads, one barrier, 32 MMAs per warp, store.  (tcgen05 would add a TMEM
// allocation and an mbarrier pipeline around 16 KB of work per CTA; the tensor-pipe rate is irrelevant here.)
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

static constexpr int RM_LD = RF_CHUNK + 8;   // bf16 elements per staged row (272 B: conflict-free fragment reads)

// grid (F / 128, ceil(M / 32)); requires F % 128 == 0
__global__ void __launch_bounds__(256) fc_reduce_f_mma_kernel(const __nv_bfloat16* __restrict__ X,
                                                              const __nv_bfloat16* __restrict__ Wm,
                                                              float* __restrict__ part, int M, int N, int F) {
  __shared__ __align__(16) __nv_bfloat16 Xs[RF_MT][RM_LD];
  __shared__ __align__(16) __nv_bfloat16 Ws[RF_NT][RM_LD];
  const int chunk = blockIdx.x, m0 = blockIdx.y * RF_MT;
  const size_t f0 = (size_t)chunk * RF_CHUNK;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int n0 = 0; n0 < N; n0 += RF_NT) {
    if (n0) __syncthreads();
    if (n0 == 0) {
      for (int i = tid; i < RF_MT * (RF_CHUNK / 8); i += 256) {
        const int r = i / (RF_CHUNK / 8), q = i % (RF_CHUNK / 8);
        *reinterpret_cast<uint4*>(&Xs[r][q * 8]) =
            (m0 + r < M) ? *reinterpret_cast<const uint4*>(X + (size_t)(m0 + r) * F + f0 + q * 8) : zero;
      }
    }
    for (int i = tid; i < RF_NT * (RF_CHUNK / 8); i += 256) {
      const int r = i / (RF_CHUNK / 8), q = i % (RF_CHUNK / 8);
      *reinterpret_cast<uint4*>(&Ws[r][q * 8]) =
          (n0 + r < N) ? *reinterpret_cast<const uint4*>(Wm + (size_t)(n0 + r) * F + f0 + q * 8) : zero;
    }
    __syncthreads();
    const int nw = warp * 16;                 // this warp's 16 output columns of the pass
    if (n0 + nw >= N) continue;
    float acc[2][2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
#pragma unroll
    for (int k0 = 0; k0 < RF_CHUNK; k0 += 16) {
      uint32_t a[2][4], b[2][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        a[i][0] = *reinterpret_cast<const uint32_t*>(&Xs[i * 16 + g][k0 + 2 * t]);
        a[i][1] = *reinterpret_cast<const uint32_t*>(&Xs[i * 16 + g + 8][k0 + 2 * t]);
        a[i][2] = *reinterpret_cast<const uint32_t*>(&Xs[i * 16 + g][k0 + 2 * t + 8]);
        a[i][3] = *reinterpret_cast<const uint32_t*>(&Xs[i * 16 + g + 8][k0 + 2 * t + 8]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        b[j][0] = *reinterpret_cast<const uint32_t*>(&Ws[nw + j * 8 + g][k0 + 2 * t]);
        b[j][1] = *reinterpret_cast<const uint32_t*>(&Ws[nw + j * 8 + g][k0 + 2 * t + 8]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) mma_bf16_16816(acc[i][j], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int m = m0 + i * 16 + g + hrow * 8, n = n0 + nw + j * 8 + 2 * t;
          if (m >= M) continue;
          float* o = part + ((size_t)chunk * M + m) * N + n;
          if (n < N) o[0] = acc[i][j][hrow * 2];
          if (n + 1 < N) o[1] = acc[i][j][hrow * 2 + 1];
        }
  }
}

template <typename T>
static int reduce_f_launch(const void* X, const void* Wm, float* part, int M, int N, int F, cudaStream_t st) {
  dim3 grid(cdiv(F, RF_CHUNK), cdiv(M, RF_MT));
  if (sizeof(T) == 2 && F % RF_CHUNK == 0) {
    fc_reduce_f_mma_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)X, (const __nv_bfloat16*)Wm, part, M, N, F);
    BVAE_LAUNCH_CHECK("fc_reduce_f_mma");
    return BVAE_OK;
  }
  fc_reduce_f_kernel<T><<<grid, 256, 0, st>>>((const T*)X, (const T*)Wm, part, M, N, F);
  BVAE_LAUNCH_CHECK("fc_reduce_f");
  return BVAE_OK;
}

// ------------------------------------------------------------------------------------------
// wide-F
// ------------------------------------------------------------------------------------------
struct FcWide {
  const float* A;          // element (m, k) at A[m*a_sm + k*a_sk]
  long long a_sm, a_sk;
  int M, K;
  int ones_row;            // row index whose A entries are all 1 (bias-gradient row), or -1
  float a_clamp;           // > 0: A values are clamped to [-a_clamp, a_clamp] on load (latent clamp)
  const void* Bm;          // [K][F]
  int F;
  const float* bias;       // [F] added to every row, or NULL
  void* out0;              // rows m <  split: out0[m*o0_sm + g(f)*o0_sf]
  long long o0_sm, o0_sf;
  void* out1;              // rows m >= split: out1[(m-split)*o1_sm + g(f)*o1_sf]
  long long o1_sm, o1_sf;
  int split;
  int permC, permP;        // g(f) = (f % C)*P + f / C when permC > 0 else f
};

static constexpr int WF_FT = 128, WF_KMAX = 128;

// With the NHWC -> NCHW feature permutation (permC > 0: the weight-gradient outputs, reference layout [..][C*P]) a
// CTA's 128 features are a (16 pixels x 8 channels) patch when P % 16 == 0 and C % 8 == 0: Bm is read as 16-byte
// channel runs, and every thread stores four consecutive pixels of one channel -- 16 contiguous bytes -- instead of
// four 4-byte stores P floats apart.  Otherwise the tile is 128 consecutive features.
//
// The whole K extent (<= 128 per pass; K is L, 2L or the batch) of both operands is staged in shared memory in ONE
// phase, every thread issuing all of its 16-byte loads before the barrier: the kernel is a stream over Bm, and a
// K loop of small tiles with a barrier pair per tile exposed DRAM latency once per tile.  TM = rows per thread
// (CTA tile = 8*TM rows x 128 features): 4 for M <= 32, 8 for the weight-gradient shapes so the Bm slab is re-read by
// half as many CTAs.
template <typename TB, typename TO, int TM>
__global__ void __launch_bounds__(256) fc_wide_f_kernel(const FcWide p, int kt) {
  constexpr int MT = 8 * TM;
  constexpr int LDA = MT + 4, LDB = WF_FT + 4;
  extern __shared__ __align__(16) float wf_sm[];
  float* As = wf_sm;                        // [kt][LDA]
  float* Bs = wf_sm + (size_t)kt * LDA;     // [kt][LDB]
  const TB* __restrict__ Bm = reinterpret_cast<const TB*>(p.Bm);
  constexpr int VB = Vec16<TB>::N;
  const int tid = threadIdx.x;
  const bool patch = p.permC > 0 && p.permP % 16 == 0 && p.permC % 8 == 0;
  const int f0 = blockIdx.x * WF_FT, m0 = blockIdx.y * MT;
  const bool bvec = !patch && (p.F % VB == 0) && (f0 + WF_FT <= p.F);   // 16-byte loads of the B tile
  int pc0 = 0, pp0 = 0;                       // patch origin (channel, pixel)
  if (patch) {
    const int ctiles = p.permC / 8;
    pc0 = (blockIdx.x % ctiles) * 8;
    pp0 = (blockIdx.x / ctiles) * 16;
  }
  const int ty = tid >> 5, tx = tid & 31;  // rows ty*TM..+TM-1, tile-local features tx*4..+3
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += kt) {
    const int kn = min(kt, p.K - k0);
    if (k0) __syncthreads();
    for (int i = tid; i < kn * MT; i += 256) {
      const int k = i / MT, m = i % MT;
      float v = 0.f;
      if (m0 + m < p.M) v = (m0 + m == p.ones_row) ? 1.f : p.A[(size_t)(m0 + m) * p.a_sm + (size_t)(k0 + k) * p.a_sk];
      if (p.a_clamp > 0.f && m0 + m != p.ones_row) v = fminf(fmaxf(v, -p.a_clamp), p.a_clamp);
      As[k * LDA + m] = v;
    }
    if (patch) {
      // memory order: channel fastest (8-channel runs, one or two 16-byte vectors); tile-local order: pixel fastest
      constexpr int NV = 8 / VB;
      for (int i = tid; i < kn * 16 * NV; i += 256) {
        const int k = i / (16 * NV), r = i % (16 * NV), pl = r / NV, hv = r % NV;
        Vec16<TB> v;
        v.load(Bm + (size_t)(k0 + k) * p.F + (size_t)(pp0 + pl) * p.permC + pc0 + hv * VB);
#pragma unroll
        for (int e = 0; e < VB; ++e) Bs[k * LDB + (hv * VB + e) * 16 + pl] = v.v[e];
      }
    } else if (bvec) {
      for (int i = tid; i < kn * (WF_FT / VB); i += 256) {
        const int k = i / (WF_FT / VB), q = i % (WF_FT / VB);
        Vec16<TB> v;
        v.load(Bm + (size_t)(k0 + k) * p.F + f0 + q * VB);
#pragma unroll
        for (int e = 0; e < VB; ++e) Bs[k * LDB + q * VB + e] = v.v[e];
      }
    } else {
      for (int i = tid; i < kn * WF_FT; i += 256) {
        const int k = i / WF_FT, fl = i % WF_FT;
        Bs[k * LDB + fl] = (f0 + fl < p.F) ? Elem<TB>::to_f(Bm[(size_t)(k0 + k) * p.F + f0 + fl]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < kn; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k * LDB + tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int h = 0; h < TM / 4; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k * LDA + ty * TM + h * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[h * 4 + i][j] = fmaf(av[i], bv[j], acc[h * 4 + i][j]);
      }
    }
  }
  if (p.o0_sm == 1) {
    // output is row-contiguous in m (fc_dec dW [F][L]): stage the tile in shared memory (reusing the operand
    // buffers) and let each warp write one feature's rows as contiguous segments
    __syncthreads();
    float* Cs = wf_sm;   // [WF_FT][MT + 1]
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Cs[(tx * 4 + j) * (MT + 1) + ty * TM + i] = acc[i][j];
    __syncthreads();
    for (int fl = ty; fl < WF_FT; fl += 8) {
      long long g;
      long long fsrc;        // NHWC feature index (bias lookup)
      if (patch) {
        const int c = pc0 + (fl >> 4), px = pp0 + (fl & 15);
        g = (long long)c * p.permP + px;
        fsrc = (long long)px * p.permC + c;
      } else {
        const int f = f0 + fl;
        if (f >= p.F) continue;
        g = p.permC > 0 ? (long long)(f % p.permC) * p.permP + f / p.permC : f;
        fsrc = f;
      }
      for (int ml = tx; ml < MT; ml += 32) {   // lane = row
        const int m = m0 + ml;
        if (m >= p.M) continue;
        const bool lo = m < p.split;
        TO* base = reinterpret_cast<TO*>(lo ? p.out0 : p.out1);
        const long long sm = lo ? p.o0_sm : p.o1_sm, sf = lo ? p.o0_sf : p.o1_sf;
        const long long mm = lo ? m : m - p.split;
        float v = Cs[fl * (MT + 1) + ml];
        if (p.bias) v += p.bias[fsrc];
        base[mm * sm + g * sf] = Elem<TO>::from_f(v);
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
    const bool lo = m < p.split;
    TO* base = reinterpret_cast<TO*>(lo ? p.out0 : p.out1);
    const long long sm = lo ? p.o0_sm : p.o1_sm, sf = lo ? p.o0_sf : p.o1_sf;
    const long long mm = lo ? m : m - p.split;
    if (patch) {
      // tile-local features tx*4..+3 = channel pc0 + tx/4, pixels pp0 + (tx%4)*4 .. +3  ->  g = c*P + pixel
      const int c = pc0 + (tx >> 2), px = pp0 + (tx & 3) * 4;
      const long long g = (long long)c * p.permP + px;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + (p.bias ? p.bias[(size_t)(px + j) * p.permC + c] : 0.f);
      if (sf == 1 && sizeof(TO) == 4) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + mm * sm + g) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) base[mm * sm + (g + j) * sf] = Elem<TO>::from_f(v[j]);
      }
      continue;
    }
    if (bvec && sf == 1 && p.permC == 0 && sizeof(TO) == 2) {
      // four consecutive features of one row: one 8-byte store
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + (p.bias ? p.bias[f0 + tx * 4 + j] : 0.f);
      __nv_bfloat162 lo2 = __floats2bfloat162_rn(v[0], v[1]), hi2 = __floats2bfloat162_rn(v[2], v[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo2);
      pk.y = *reinterpret_cast<uint32_t*>(&hi2);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + mm * sm + f0 + tx * 4) = pk;
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = f0 + tx * 4 + j;
      if (f >= p.F) continue;
      const long long g = p.permC > 0 ? (long long)(f % p.permC) * p.permP + f / p.permC : f;
      float v = acc[i][j];
      if (p.bias) v += p.bias[f];
      base[mm * sm + g * sf] = Elem<TO>::from_f(v);
    }
  }
}

// ---- bf16: wide-F on the warp-level tensor-core path (see fc_reduce_f_mma_kernel for why mma.sync) ------------
// CTA = 128 features x all M rows (<= 16 * MTILES); warp w owns features [16w, 16w + 16).  Bm is staged as
// [k][feature] and read through ldmatrix.trans (the contraction index is its slow dimension); the small fp32
// operand A (gradients / latents) is split into bf16 hi + lo parts, two MMAs per tile, so the product keeps fp32-level
// accuracy in A while Bm -- already bf16 in memory -- is used as is.
static constexpr int WM_LDB = WF_FT + 8;   // bf16 elements per staged Bm row

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <typename TO>
__device__ __forceinline__ void wm_store2(TO* dst, float a, float b);
template <>
__device__ __forceinline__ void wm_store2<float>(float* dst, float a, float b) {
  *reinterpret_cast<float2*>(dst) = make_float2(a, b);
}
template <>
__device__ __forceinline__ void wm_store2<__nv_bfloat16>(__nv_bfloat16* dst, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(a, b);
}

// generic permuted scatter (permC > 0 without the patch mapping): rare, kept out of line so the hot epilogue stays small
template <typename TO>
__device__ __noinline__ void wm_scatter2(TO* rowp, long long sf, long long f, int C, int P, float v0, float v1) {
  rowp[((f % C) * P + f / C) * sf] = Elem<TO>::from_f(v0);
  rowp[(((f + 1) % C) * P + (f + 1) / C) * sf] = Elem<TO>::from_f(v1);
}

// kt = 1 << ktlog = K rounded up to a power of two in [16, 128] (rows beyond K are zero), at most 128 per pass; a
// CTA stages up to 128 rows of A and walks them in blocks of 32 (two m16 tiles).  Requires every tile inside F
// (F % 128 == 0).  The code is deliberately compact -- shift/mask indexing, precomputed row pointers, a real loop over
// the row blocks: every warp runs it once, and a fully unrolled 8-tile version spent 53 % of its issue slots waiting
// for instruction fetch (ncu: stall "no_instructions").  Passes after the first (K > 128: the batch of a weight
// gradient) add into the fp32 output, same thread, same order: deterministic.
template <typename TO>
__global__ void __launch_bounds__(256) fc_wide_f_mma_kernel(const FcWide p, int ktlog, int mrlog) {
  const int kt = 1 << ktlog, mr = 1 << mrlog;   // rows staged per CTA: 32, 64 or 128
  extern __shared__ __align__(16) uint8_t wm_raw[];
  const int lda = kt + 8;
  __nv_bfloat16* Bs = reinterpret_cast<__nv_bfloat16*>(wm_raw);   // [kt][WM_LDB]
  __nv_bfloat16* Ahi = Bs + (size_t)kt * WM_LDB;                  // [mr][lda]
  __nv_bfloat16* Alo = Ahi + (size_t)mr * lda;
  const __nv_bfloat16* __restrict__ Bm = reinterpret_cast<const __nv_bfloat16*>(p.Bm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const bool patch = p.permC > 0 && p.permP % 16 == 0 && p.permC % 8 == 0;
  const int f0 = blockIdx.x * WF_FT, m0 = blockIdx.y * mr;
  int pc0 = 0, pp0 = 0;
  if (patch) {
    const int ctiles = p.permC / 8;
    pc0 = (blockIdx.x % ctiles) * 8;
    pp0 = (blockIdx.x / ctiles) * 16;
  }
  // output columns of this thread: tile-local features nw + 8j + 2t, +1
  const int nw = warp * 16;
  const bool pair = patch || p.permC == 0;   // the two features of a fragment pair are adjacent in the output
  long long goff[2];
  float bias0[2], bias1[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int fl = nw + j * 8 + 2 * t;
    long long fsrc0, fsrc1;
    if (patch) {
      const int c = pc0 + (fl >> 4), px = pp0 + (fl & 15);
      goff[j] = (long long)c * p.permP + px;       // the next local feature is the next pixel of the same channel
      fsrc0 = (long long)px * p.permC + c;
      fsrc1 = fsrc0 + p.permC;
    } else {
      goff[j] = f0 + fl;
      fsrc0 = goff[j];
      fsrc1 = fsrc0 + 1;
    }
    bias0[j] = p.bias ? p.bias[fsrc0] : 0.f;
    bias1[j] = p.bias ? p.bias[fsrc1] : 0.f;
  }
  const int rows = min(mr, p.M - m0);
  const bool kfast = p.a_sk == 1;
  for (int k0 = 0; k0 < p.K; k0 += kt) {
    const int kn = min(kt, p.K - k0);     // live rows of this pass; rows kn..kt-1 are zero
    if (k0) __syncthreads();
    // Every staging loop issues a batch of eight independent loads before it touches shared memory: with one load per
    // iteration each trip paid a full L2 / DRAM round trip and the CTA spent most of its life waiting.
    // Bm -> [k][tile-local feature]; patch order: pixel fastest within a channel.  (k, pl) / (k, q) = i / 16, i % 16.
    for (int base = 0; base < kt * 16; base += 256 * 8) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * 256 + tid, k = i >> 4, r = i & 15;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (i < kt * 16 && k < kn)
          v[u] = patch ? *reinterpret_cast<const uint4*>(Bm + (size_t)(k0 + k) * p.F + (size_t)(pp0 + r) * p.permC + pc0)
                       : *reinterpret_cast<const uint4*>(Bm + (size_t)(k0 + k) * p.F + f0 + r * 8);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * 256 + tid, k = i >> 4, r = i & 15;
        if (i >= kt * 16) break;
        if (patch) {
          const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&v[u]);
#pragma unroll
          for (int c = 0; c < 8; ++c) Bs[k * WM_LDB + c * 16 + r] = e[c];
        } else {
          *reinterpret_cast<uint4*>(&Bs[k * WM_LDB + r * 8]) = v[u];
        }
      }
    }
    // A -> bf16 hi / lo, [m][k]
    for (int base = 0; base < mr * kt; base += 256 * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * 256 + tid;
        const int k = kfast ? (i & (kt - 1)) : (i >> mrlog), m = kfast ? (i >> ktlog) : (i & (mr - 1));
        v[u] = 0.f;
        if (i < mr * kt && k < kn && m < rows && m0 + m != p.ones_row)
          v[u] = p.A[(size_t)(m0 + m) * p.a_sm + (size_t)(k0 + k) * p.a_sk];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * 256 + tid;
        if (i >= mr * kt) break;
        const int k = kfast ? (i & (kt - 1)) : (i >> mrlog), m = kfast ? (i >> ktlog) : (i & (mr - 1));
        float x = v[u];
        if (p.a_clamp > 0.f) x = fminf(fmaxf(x, -p.a_clamp), p.a_clamp);
        if (m0 + m == p.ones_row && k < kn) x = 1.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(x);
        Ahi[m * lda + k] = hi;
        Alo[m * lda + k] = __float2bfloat16_rn(x - __bfloat162float(hi));
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int mb = 0; mb < rows; mb += 32) {
      float acc[2][2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
#pragma unroll 1
      for (int kk = 0; kk < kt; kk += 16) {
        uint32_t b[4];   // {n-tile 0: k lo, k hi; n-tile 1: k lo, k hi}
        ldmatrix_x4_trans(b, &Bs[(kk + ((lane >> 3) & 1) * 8 + (lane & 7)) * WM_LDB + nw + (lane >> 4) * 8]);
        const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint32_t ah[4], al[4];
          const int r0 = (mb + i * 16 + g) * lda + kk + 2 * t, r1 = r0 + 8 * lda;
          ah[0] = *reinterpret_cast<const uint32_t*>(&Ahi[r0]);
          ah[1] = *reinterpret_cast<const uint32_t*>(&Ahi[r1]);
          ah[2] = *reinterpret_cast<const uint32_t*>(&Ahi[r0 + 8]);
          ah[3] = *reinterpret_cast<const uint32_t*>(&Ahi[r1 + 8]);
          al[0] = *reinterpret_cast<const uint32_t*>(&Alo[r0]);
          al[1] = *reinterpret_cast<const uint32_t*>(&Alo[r1]);
          al[2] = *reinterpret_cast<const uint32_t*>(&Alo[r0 + 8]);
          al[3] = *reinterpret_cast<const uint32_t*>(&Alo[r1 + 8]);
          mma_bf16_16816(acc[i][0], ah, b0);
          mma_bf16_16816(acc[i][0], al, b0);
          mma_bf16_16816(acc[i][1], ah, b1);
          mma_bf16_16816(acc[i][1], al, b1);
        }
      }
      // rows mb + 16i + g (+8); thread holds two adjacent features per (i, j, half)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int m = m0 + mb + i * 16 + g + hrow * 8;
          if (m >= p.M) continue;
          const bool lo = m < p.split;
          TO* rowp = lo ? reinterpret_cast<TO*>(p.out0) + (long long)m * p.o0_sm
                        : reinterpret_cast<TO*>(p.out1) + (long long)(m - p.split) * p.o1_sm;
          const long long sf = lo ? p.o0_sf : p.o1_sf;
          const bool vec = pair && sf == 1 && (reinterpret_cast<uintptr_t>(rowp) & (2 * sizeof(TO) - 1)) == 0;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float v0 = acc[i][j][hrow * 2], v1 = acc[i][j][hrow * 2 + 1];
            if (k0 == 0) {
              v0 += bias0[j];
              v1 += bias1[j];
            } else if (pair) {   // later K passes add into the (fp32) output
              v0 += Elem<TO>::to_f(rowp[goff[j] * sf]);
              v1 += Elem<TO>::to_f(rowp[(goff[j] + 1) * sf]);
            }
            if (vec) {
              wm_store2<TO>(rowp + goff[j], v0, v1);     // goff is even: aligned pair
            } else if (pair) {
              rowp[goff[j] * sf] = Elem<TO>::from_f(v0);
              rowp[(goff[j] + 1) * sf] = Elem<TO>::from_f(v1);
            } else {
              wm_scatter2<TO>(rowp, sf, goff[j], p.permC, p.permP, v0, v1);
            }
          }
        }
      }
    }
  }
}

// bf16 Bm, every tile inside F, and either a single K pass or an fp32 output that later passes can add into
template <typename TO>
static bool wide_f_mma_ok(const FcWide& p) {
  if (p.F % WF_FT) return false;
  if (p.K <= WF_KMAX) return true;
  const bool pair = p.permC == 0 || (p.permP % 16 == 0 && p.permC % 8 == 0);
  return sizeof(TO) == 4 && pair;
}

template <typename TO>
static int wide_f_mma_launch(const FcWide& p, cudaStream_t st) {
  int ktlog = 4;
  while ((1 << ktlog) < p.K && (1 << ktlog) < WF_KMAX) ++ktlog;
  const int kt = 1 << ktlog;
  int mrlog = 5;
  while ((1 << mrlog) < p.M && mrlog < 7) ++mrlog;
  const int mr = 1 << mrlog;
  const size_t smem = ((size_t)kt * WM_LDB + 2 * (size_t)mr * (kt + 8)) * sizeof(__nv_bfloat16);
  auto kern = fc_wide_f_mma_kernel<TO>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "fc_wide_f_mma: cudaFuncSetAttribute");
  }
  dim3 grid(p.F / WF_FT, cdiv(p.M, mr));
  kern<<<grid, 256, smem, st>>>(p, ktlog, mrlog);
  BVAE_LAUNCH_CHECK("fc_wide_f_mma");
  return BVAE_OK;
}

template <typename TB, typename TO>
static int wide_f_launch(const FcWide& p, cudaStream_t st) {
  if (sizeof(TB) == 2 && wide_f_mma_ok<TO>(p)) return wide_f_mma_launch<TO>(p, st);

  const int kt = p.K < WF_KMAX ? p.K : WF_KMAX;
  const bool big = p.M > 32;
  const int MT = big ? 64 : 32;
  size_t smem = (size_t)kt * ((MT + 4) + (WF_FT + 4)) * sizeof(float);
  const size_t cs = (size_t)WF_FT * (MT + 1) * sizeof(float);
  if (p.o0_sm == 1 && cs > smem) smem = cs;
  dim3 grid(cdiv(p.F, WF_FT), cdiv(p.M, MT));
  auto kern = big ? fc_wide_f_kernel<TB, TO, 8> : fc_wide_f_kernel<TB, TO, 4>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(WF_KMAX * (68 + 132) * sizeof(float)));
    if (e != cudaSuccess) return cuda_fail(e, "fc_wide_f: cudaFuncSetAttribute");
  }
  kern<<<grid, 256, smem, st>>>(p, kt);
  BVAE_LAUNCH_CHECK("fc_wide_f");
  return BVAE_OK;
}

// ------------------------------------------------------------------------------------------
// heads: finalize kernels
// ------------------------------------------------------------------------------------------
// grid = B, block = 2L threads (rounded up to a warp multiple): sum the F-chunks, add bias, clamp,
// reparameterise, per-element KL.
// out[i] = sum_c part[c][i] (fixed order).  32 elements x 8 chunk groups per CTA, four loads in flight per thread;
// `out` may alias chunk 0 of `part` (element i is only touched by the CTA that owns it).
// `z` non-null (fc_dec backward): the sum is dz, zeroed where the latent clamp saturated (|z| > clamp) -- the epilogue
// that used to be its own launch (fcdec_dz_kernel).
__global__ void __launch_bounds__(256) chunk_sum_kernel(const float* part, int nchunks, int n, float* out,
                                                        const float* __restrict__ z = nullptr, float clamp = 0.f) {
  __shared__ float sm[8][32];
  const int e = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + e;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  if (i < n) {
    int c = grp;
    for (; c + 24 < nchunks; c += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] += part[(size_t)(c + 8 * u) * n + i];
    }
    for (; c < nchunks; c += 8) a[0] += part[(size_t)c * n + i];
  }
  sm[grp][e] = (a[0] + a[1]) + (a[2] + a[3]);
  __syncthreads();
  if (grp == 0 && i < n) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) v += sm[g][e];
    if (z != nullptr && clamp > 0.f && (z[i] < -clamp || z[i] > clamp)) v = 0.f;
    out[i] = v;
  }
}

__global__ void heads_finalize_kernel(const float* __restrict__ part, int nchunks, const float* __restrict__ b_mu,
                                      const float* __restrict__ b_lv, const float* __restrict__ eps, float lo,
                                      float hi, int deterministic, float* __restrict__ mu, float* __restrict__ logvar,
                                      float* __restrict__ z, float* __restrict__ kl_elem,
                                      uint8_t* __restrict__ lv_pass, int B, int L) {
  const int b = blockIdx.x;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    float m = b_mu[l], v = b_lv[l];
    for (int c = 0; c < nchunks; ++c) {
      const float* row = part + ((size_t)c * B + b) * 2 * L;
      m += row[l];
      v += row[L + l];
    }
    const bool pass = (v >= lo) && (v <= hi);  // torch.clamp: gradient 1 inside [lo, hi], 0 outside
    const float lv = fminf(fmaxf(v, lo), hi);
    const size_t o = (size_t)b * L + l;
    mu[o] = m;
    logvar[o] = lv;
    lv_pass[o] = pass ? 1 : 0;
    if (deterministic || eps == nullptr) {
      z[o] = m;
      kl_elem[o] = 0.f;
    } else {
      z[o] = fmaf(eps[o], expf(0.5f * lv), m);
      kl_elem[o] = -0.5f * (1.f + lv - m * m - expf(lv));
    }
  }
}

// single CTA: kl_per_dim = mean_b kl_elem; scalars = {kl_mean, kl_effective, objective, latent_reg}
__global__ void __launch_bounds__(256)
    heads_kl_kernel(const float* __restrict__ kl_elem, const float* __restrict__ mu, float beta, int use_capacity,
                    float capacity, float capacity_weight, float free_bits, float reg_lambda, int deterministic,
                    float* __restrict__ kl_per_dim, float* __restrict__ scalars, int B, int L) {
  __shared__ float red[32];
  float klsum = 0.f, kleff = 0.f;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += kl_elem[(size_t)b * L + l];
    const float pd = acc / (float)B;
    kl_per_dim[l] = pd;
    klsum += pd;
    kleff += (free_bits > 0.f && !use_capacity) ? fmaxf(pd, free_bits) : pd;
  }
  float musq = 0.f;
  if (reg_lambda > 0.f)
    for (int i = threadIdx.x; i < B * L; i += blockDim.x) musq = fmaf(mu[i], mu[i], musq);
  klsum = block_sum(klsum, red);
  kleff = block_sum(kleff, red);
  musq = block_sum(musq, red);
  if (threadIdx.x == 0) {
    const float reg = reg_lambda > 0.f ? reg_lambda * musq / ((float)B * (float)L) : 0.f;
    float obj;
    if (deterministic) {
      klsum = 0.f;
      kleff = 0.f;
      obj = reg;
    } else if (use_capacity) {
      obj = capacity_weight * fabsf(klsum - capacity) + reg;
    } else {
      obj = beta * kleff + reg;
    }
    scalars[0] = klsum;
    scalars[1] = kleff;
    scalars[2] = obj;
    scalars[3] = reg;
  }
}

// grid = B: D[b][0..L) = d mu, D[b][L..2L) = d (pre-clamp logvar)
__global__ void heads_bwd_prep_kernel(const float* __restrict__ eps, const float* __restrict__ mu,
                                      const float* __restrict__ logvar, const float* __restrict__ kl_per_dim,
                                      const float* __restrict__ scalars, const uint8_t* __restrict__ lv_pass,
                                      const float* __restrict__ g_mu, const float* __restrict__ g_lv,
                                      const float* __restrict__ g_z, const float* __restrict__ g_sc, float beta,
                                      int use_capacity, float capacity, float capacity_weight, float free_bits,
                                      float reg_lambda, int deterministic, float* __restrict__ D, int B, int L) {
  const int b = blockIdx.x;
  const float g_klmean = g_sc ? g_sc[0] : 0.f;
  const float g_obj = g_sc ? g_sc[2] : 0.f;
  const float invB = 1.f / (float)B;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const size_t o = (size_t)b * L + l;
    const float m = mu[o], lv = logvar[o];
    float dm = g_mu ? g_mu[o] : 0.f;
    float dl = g_lv ? g_lv[o] : 0.f;
    const float gz = g_z ? g_z[o] : 0.f;
    dm += gz;
    if (reg_lambda > 0.f) dm = fmaf(g_obj * reg_lambda * 2.f / ((float)B * (float)L), m, dm);
    if (!deterministic && eps != nullptr) {
      float c;  // d(objective + kl_mean terms) / d kl_elem[b][l]
      if (use_capacity) {
        const float d = scalars[0] - capacity;
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        c = g_obj * capacity_weight * sgn * invB;
      } else if (free_bits > 0.f) {
        c = (kl_per_dim[l] >= free_bits) ? g_obj * beta * invB : 0.f;
      } else {
        c = g_obj * beta * invB;
      }
      c += g_klmean * invB;
      dm = fmaf(c, m, dm);                             // d kl_elem / d mu = mu
      dl += gz * eps[o] * 0.5f * expf(0.5f * lv);      // z = mu + eps*exp(lv/2)
      dl = fmaf(c, 0.5f * (expf(lv) - 1.f), dl);       // d kl_elem / d lv = (e^lv - 1)/2
    }
    if (!lv_pass[o]) dl = 0.f;
    D[(size_t)b * 2 * L + l] = dm;
    D[(size_t)b * 2 * L + L + l] = dl;
  }
}

__global__ void heads_db_kernel(const float* __restrict__ D, float* __restrict__ db_mu, float* __restrict__ db_lv,
                                int B, int L) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= 2 * L) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc += D[(size_t)b * 2 * L + n];
  if (n < L)
    db_mu[n] = acc;
  else
    db_lv[n - L] = acc;
}

struct HeadsWs {
  float* part;     // [nchunks][B][2L]
  float* kl_elem;  // [B][L]
  float* D;        // [B][2L]
  size_t bytes;
};
static HeadsWs heads_ws_carve(void* base, int B, int L, int F) {
  HeadsWs w;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.part = reinterpret_cast<float*>(take((size_t)cdiv(F, RF_CHUNK) * B * 2 * L * sizeof(float)));
  w.kl_elem = reinterpret_cast<float*>(take((size_t)B * L * sizeof(float)));
  w.D = reinterpret_cast<float*>(take((size_t)B * 2 * L * sizeof(float)));
  w.bytes = off;
  return w;
}

struct FcDecWs {
  float* part;  // [nchunks][B][L]
  size_t bytes;
};
static FcDecWs fcdec_ws_carve(void* base, int B, int L, int F) {
  FcDecWs w;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.part = reinterpret_cast<float*>(take((size_t)cdiv(F, RF_CHUNK) * B * L * sizeof(float)));
  w.bytes = off;
  return w;
}

static inline int warp_round(int n) {
  int t = (n + 31) / 32 * 32;
  return t > 256 ? 256 : (t < 32 ? 32 : t);
}

}  // namespace bvae

using namespace bvae;

extern "C" size_t bvae_heads_workspace(int B, int L, int F) {
  if (B <= 0 || L <= 0 || F <= 0) return 0;
  return heads_ws_carve(nullptr, B, L, F).bytes;
}

extern "C" int bvae_heads_reparam_kl_fwd(const void* h, const void* wh, const float* b_mu, const float* b_logvar,
                                         const float* eps, float lv_lo, float lv_hi, float beta, int use_capacity,
                                         float capacity, float capacity_weight, float free_bits, float reg_lambda,
                                         int deterministic, float* mu, float* logvar, float* z, float* kl_per_dim,
                                         float* scalars, uint8_t* lv_pass, void* ws, size_t ws_bytes, int B, int L,
                                         int F, int dtype, bvae_stream_t stream) {
  BVAE_REQUIRE(h && wh && b_mu && b_logvar && mu && logvar && z && kl_per_dim && scalars && lv_pass && ws,
               BVAE_ERR_ARG, "heads_fwd: null pointer");
  BVAE_REQUIRE(B > 0 && L > 0 && F > 0, BVAE_ERR_ARG, "heads_fwd: empty shape");
  BVAE_REQUIRE(ws_bytes >= bvae_heads_workspace(B, L, F), BVAE_ERR_WORKSPACE, "heads_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  HeadsWs w = heads_ws_carve(ws, B, L, F);
  int rc;
  if (dtype == BVAE_F32)
    rc = reduce_f_launch<float>(h, wh, w.part, B, 2 * L, F, st);
  else if (dtype == BVAE_BF16)
    rc = reduce_f_launch<__nv_bfloat16>(h, wh, w.part, B, 2 * L, F, st);
  else
    BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "heads_fwd: dtype %d", dtype);
  if (rc != BVAE_OK) return rc;
  chunk_sum_kernel<<<cdiv(B * 2 * L, 32), 256, 0, st>>>(w.part, cdiv(F, RF_CHUNK), B * 2 * L, w.part);
  BVAE_LAUNCH_CHECK("heads_chunk_sum");
  heads_finalize_kernel<<<B, warp_round(L), 0, st>>>(w.part, 1, b_mu, b_logvar, eps, lv_lo, lv_hi,
                                                    deterministic, mu, logvar, z, w.kl_elem, lv_pass, B, L);
  BVAE_LAUNCH_CHECK("heads_finalize");
  heads_kl_kernel<<<1, 256, 0, st>>>(w.kl_elem, mu, beta, use_capacity, capacity, capacity_weight, free_bits,
                                     reg_lambda, (deterministic || eps == nullptr) ? 1 : 0, kl_per_dim, scalars, B, L);
  BVAE_LAUNCH_CHECK("heads_kl");
  return BVAE_OK;
}

extern "C" int bvae_heads_reparam_kl_bwd(const void* h, const void* wh, const float* eps, const float* mu,
                                         const float* logvar, const float* kl_per_dim, const float* scalars,
                                         const uint8_t* lv_pass, const float* g_mu, const float* g_logvar,
                                         const float* g_z, const float* g_scalars, float beta, int use_capacity,
                                         float capacity, float capacity_weight, float free_bits, float reg_lambda,
                                         int deterministic, void* dh, float* dw_mu, float* dw_logvar, float* db_mu,
                                         float* db_logvar, void* ws, size_t ws_bytes, int B, int L, int C, int P,
                                         int dtype, bvae_stream_t stream) {
  BVAE_REQUIRE(h && wh && mu && logvar && kl_per_dim && scalars && lv_pass && dw_mu && dw_logvar && db_mu &&
                   db_logvar && ws,
               BVAE_ERR_ARG, "heads_bwd: null pointer");
  BVAE_REQUIRE(B > 0 && L > 0 && C > 0 && P > 0, BVAE_ERR_ARG, "heads_bwd: empty shape");
  const int F = C * P;
  BVAE_REQUIRE(ws_bytes >= bvae_heads_workspace(B, L, F), BVAE_ERR_WORKSPACE, "heads_bwd: workspace too small");
  BVAE_REQUIRE(dtype == BVAE_F32 || dtype == BVAE_BF16, BVAE_ERR_UNSUPPORTED, "heads_bwd: dtype %d", dtype);
  cudaStream_t st = as_stream(stream);
  HeadsWs w = heads_ws_carve(ws, B, L, F);
  heads_bwd_prep_kernel<<<B, warp_round(L), 0, st>>>(eps, mu, logvar, kl_per_dim, scalars, lv_pass, g_mu, g_logvar,
                                                    g_z, g_scalars, beta, use_capacity, capacity, capacity_weight,
                                                    free_bits, reg_lambda, (deterministic || eps == nullptr) ? 1 : 0,
                                                    w.D, B, L);
  BVAE_LAUNCH_CHECK("heads_bwd_prep");
  int rc = BVAE_OK;
  if (dh) {  // dh[b][f'] = sum_n D[b][n] * wh[n][f']
    FcWide p;
    p.A = w.D; p.a_sm = 2 * L; p.a_sk = 1; p.M = B; p.K = 2 * L; p.ones_row = -1; p.a_clamp = 0.f;
    p.Bm = wh; p.F = F; p.bias = nullptr;
    p.out0 = dh; p.o0_sm = F; p.o0_sf = 1; p.out1 = dh; p.o1_sm = F; p.o1_sf = 1; p.split = B;
    p.permC = 0; p.permP = 0;
    rc = (dtype == BVAE_F32) ? wide_f_launch<float, float>(p, st) : wide_f_launch<__nv_bfloat16, __nv_bfloat16>(p, st);
    if (rc != BVAE_OK) return rc;
  }
  {  // dW[n][g(f')] = sum_b D[b][n] * h[b][f']
    FcWide p;
    p.A = w.D; p.a_sm = 1; p.a_sk = 2 * L; p.M = 2 * L; p.K = B; p.ones_row = -1; p.a_clamp = 0.f;
    p.Bm = h; p.F = F; p.bias = nullptr;
    p.out0 = dw_mu; p.o0_sm = F; p.o0_sf = 1; p.out1 = dw_logvar; p.o1_sm = F; p.o1_sf = 1; p.split = L;
    p.permC = C; p.permP = P;
    rc = (dtype == BVAE_F32) ? wide_f_launch<float, float>(p, st) : wide_f_launch<__nv_bfloat16, float>(p, st);
    if (rc != BVAE_OK) return rc;
  }
  heads_db_kernel<<<cdiv(2 * L, 128), 128, 0, st>>>(w.D, db_mu, db_logvar, B, L);
  BVAE_LAUNCH_CHECK("heads_db");
  return BVAE_OK;
}

extern "C" size_t bvae_fc_dec_workspace(int B, int L, int F) {
  if (B <= 0 || L <= 0 || F <= 0) return 0;
  return fcdec_ws_carve(nullptr, B, L, F).bytes;
}

extern "C" int bvae_fc_dec_fwd(const float* z, const void* wd, const float* bd, void* h0, float z_clamp, int B, int L,
                               int F, int dtype, bvae_stream_t stream) {
  BVAE_REQUIRE(z && wd && bd && h0, BVAE_ERR_ARG, "fc_dec_fwd: null pointer");
  BVAE_REQUIRE(B > 0 && L > 0 && F > 0, BVAE_ERR_ARG, "fc_dec_fwd: empty shape");
  BVAE_REQUIRE(dtype == BVAE_F32 || dtype == BVAE_BF16, BVAE_ERR_UNSUPPORTED, "fc_dec_fwd: dtype %d", dtype);
  cudaStream_t st = as_stream(stream);
  FcWide p;
  p.A = z; p.a_sm = L; p.a_sk = 1; p.M = B; p.K = L; p.ones_row = -1; p.a_clamp = z_clamp;
  p.Bm = wd; p.F = F; p.bias = bd;
  p.out0 = h0; p.o0_sm = F; p.o0_sf = 1; p.out1 = h0; p.o1_sm = F; p.o1_sf = 1; p.split = B;
  p.permC = 0; p.permP = 0;
  return (dtype == BVAE_F32) ? wide_f_launch<float, float>(p, st) : wide_f_launch<__nv_bfloat16, __nv_bfloat16>(p, st);
}

extern "C" int bvae_fc_dec_bwd(const void* dh0, const float* z, const void* wd, float z_clamp, float* dz, float* dw,
                               float* db, void* ws, size_t ws_bytes, int B, int L, int C, int P, int dtype,
                               bvae_stream_t stream) {
  BVAE_REQUIRE(dh0 && z && wd && dz && dw && db && ws, BVAE_ERR_ARG, "fc_dec_bwd: null pointer");
  BVAE_REQUIRE(B > 0 && L > 0 && C > 0 && P > 0, BVAE_ERR_ARG, "fc_dec_bwd: empty shape");
  const int F = C * P;
  BVAE_REQUIRE(ws_bytes >= bvae_fc_dec_workspace(B, L, F), BVAE_ERR_WORKSPACE, "fc_dec_bwd: workspace too small");
  BVAE_REQUIRE(dtype == BVAE_F32 || dtype == BVAE_BF16, BVAE_ERR_UNSUPPORTED, "fc_dec_bwd: dtype %d", dtype);
  cudaStream_t st = as_stream(stream);
  FcDecWs w = fcdec_ws_carve(ws, B, L, F);
  // dz = dh0 . wd^T (reduce over F), then mask by the latent clamp
  int rc = (dtype == BVAE_F32) ? reduce_f_launch<float>(dh0, wd, w.part, B, L, F, st)
                               : reduce_f_launch<__nv_bfloat16>(dh0, wd, w.part, B, L, F, st);
  if (rc != BVAE_OK) return rc;
  const int n = B * L;
  chunk_sum_kernel<<<cdiv(n, 32), 256, 0, st>>>(w.part, cdiv(F, RF_CHUNK), n, dz, z, z_clamp);
  BVAE_LAUNCH_CHECK("fcdec_chunk_sum_dz");
  // dW[g(f')][l] = sum_b zc[b][l] * dh0[b][f'];  db[g(f')] = sum_b dh0[b][f']  (ones row l == L)
  FcWide p;
  p.A = z; p.a_sm = 1; p.a_sk = L; p.M = L + 1; p.K = B; p.ones_row = L; p.a_clamp = z_clamp;
  p.Bm = dh0; p.F = F; p.bias = nullptr;
  p.out0 = dw; p.o0_sm = 1; p.o0_sf = L; p.out1 = db; p.o1_sm = 0; p.o1_sf = 1; p.split = L;
  p.permC = C; p.permP = P;
  return (dtype == BVAE_F32) ? wide_f_launch<float, float>(p, st) : wide_f_launch<__nv_bfloat16, float>(p, st);
}
