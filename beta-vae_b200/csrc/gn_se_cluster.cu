// GroupNorm(1,C) + ReLU + Squeeze-and-Excitation as ONE kernel per direction: a thread-block cluster
// owns a sample and keeps it resident in (distributed) shared memory.
// (reference ops: native_group_norm + relu_ + adaptive_avg_pool2d + 2x linear + sigmoid + mul,
//  src/models/beta_vae.py:41-46, 61-66 and src/models/se_blocks.py:15-19, and their autograd backward.)
//
// GroupNorm(1,C) statistics and the SE squeeze both span the whole sample (up to 2 MB in bf16), so the
// multi-kernel formulation (gn_se.cu) reads the tensor three times.  Here CS CTAs (CS = 1..16, one cluster)
// split the sample's pixels; each CTA pulls its contiguous slice into shared memory with bulk async copies,
// and the three sample-wide quantities -- (sum, sum of squares), the per-channel positive-pixel sums, the
// SE hidden / gate vectors -- are all-gathered across the cluster through distributed shared memory
// (remote stores + cluster barrier; every CTA then reduces the CS partials in the same fixed order, so
// results are deterministic and identical in every CTA).  HBM traffic is the algorithmic minimum:
// forward one read + one write, backward two reads + one write.
//
// Thread mapping as in gn_se.cu: thread t owns the 16-byte channel vector  t % cv  and walks pixel rows
// t / cv, + threads / cv, ...  (cv = C / 8 for bf16 must divide the block size).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "gn_cluster.cuh"

namespace cg = cooperative_groups;

namespace bvae {

static constexpr int GC_THREADS = 512;
static constexpr float kGcEps = 1e-5f;  // nn.GroupNorm default

// dev builds (-DBVAE_GC_TIMING): cycle stamps of CTA (0, 0) at the phase boundaries, read by tools/gc_phase_times.py
#ifdef BVAE_GC_TIMING
__device__ long long g_gc_ts[32];
#define GC_TS(i)                                                                      \
  do {                                                                                \
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_gc_ts[i] = clock64(); \
  } while (0)
#else
#define GC_TS(i)
#endif

__device__ __forceinline__ uint32_t gc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void gc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "GC_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra GC_WAIT_DONE;\n\t"
      "bra GC_WAIT_LOOP;\n\t"
      "GC_WAIT_DONE:\n\t"
      "}" ::"r"(gc_smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> this CTA's shared memory, completion on an mbarrier
__device__ __forceinline__ void gc_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   gc_smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(gc_smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void gc_load_slice(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  // one thread: arm the barrier with the total and issue <= 32 KB pieces
  gc_mbar_expect_tx(bar, bytes);
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
    gc_bulk_load(reinterpret_cast<char*>(dst) + off, reinterpret_cast<const char*>(src) + off, n, bar);
  }
}

// Per-channel sums over this CTA's pixels of two per-thread accumulator sets -> out0[c], out1[c] (fixed order).
// Thread t holds channels (t % cv) * V .. + V of pixel rows t / cv, + rows, ...  Rows that share a warp (cv < 32) are
// combined first with an xor butterfly in registers; the per-warp (or per-row) partials then go to shared memory as
// 16-byte stores -- the former scalar stores at a 32-byte lane stride were 8-way bank conflicts, ~2000 cycles per
// call -- and 2*C threads add the <= 16 partials of one output each, in double.  fp32 activations (V = 4) run the
// butterfly in double as well: the fp32 path is held to 1e-5 of the fp32 reference, the bf16 path is not.
template <int V>
__device__ __forceinline__ void gc_channel_reduce(float (&a0)[V], float (&a1)[V], float* scratch, int C, int cv, int rows,
                                                  float* out0, float* out1) {
  constexpr bool kPrecise = V == 4;      // fp32 activations: every combine in double (parity with the fp32 reference at 1e-5)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = threadIdx.x % cv;
  int nparts;
  if (cv < 32) {
    nparts = (int)blockDim.x >> 5;
    if (kPrecise) {
      double d0[V], d1[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        d0[i] = (double)a0[i];
        d1[i] = (double)a1[i];
      }
      for (int o = cv; o < 32; o <<= 1) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          d0[i] += __shfl_xor_sync(0xffffffffu, d0[i], o);
          d1[i] += __shfl_xor_sync(0xffffffffu, d1[i], o);
        }
      }
      if (lane < cv) {
        double* p0 = reinterpret_cast<double*>(scratch) + (size_t)warp * 2 * C + col * V;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          p0[i] = d0[i];
          p0[C + i] = d1[i];
        }
      }
    } else {
      for (int o = cv; o < 32; o <<= 1) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          a0[i] += __shfl_xor_sync(0xffffffffu, a0[i], o);
          a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], o);
        }
      }
    }
  } else {
    nparts = rows;
  }
  const bool dbl = kPrecise && cv < 32;
  if (!dbl && (cv >= 32 || lane < cv)) {
    const int part = cv < 32 ? warp : (int)threadIdx.x / cv;
    float* d0 = scratch + (size_t)part * 2 * C + col * V;
    float* d1 = d0 + C;
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      *reinterpret_cast<float4*>(d0 + i) = make_float4(a0[i], a0[i + 1], a0[i + 2], a0[i + 3]);
      *reinterpret_cast<float4*>(d1 + i) = make_float4(a1[i], a1[i + 1], a1[i + 2], a1[i + 3]);
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 2 * C; o += blockDim.x) {
    double acc = 0.0;
    if (dbl) {
      for (int r = 0; r < nparts; ++r) acc += reinterpret_cast<const double*>(scratch)[(size_t)r * 2 * C + o];
    } else {
      for (int r = 0; r < nparts; ++r) acc += (double)scratch[(size_t)r * 2 * C + o];
    }
    if (o < C) out0[o] = (float)acc;
    else out1[o - C] = (float)acc;
  }
  __syncthreads();
}

// two block-wide double sums with one set of barriers; `red` holds 64 doubles
__device__ __forceinline__ void gc_block_sum2(double& a, double& b, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (int)blockDim.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if (lane == 0) {
    red[w] = a;
    red[32 + w] = b;
  }
  __syncthreads();
  if (w == 0) {
    double ra = lane < nw ? red[lane] : 0.0, rb = lane < nw ? red[32 + lane] : 0.0;
    ra = warp_sum(ra);
    rb = warp_sum(rb);
    if (lane == 0) {
      red[0] = ra;
      red[32] = rb;
    }
  }
  __syncthreads();
  a = red[0];
  b = red[32];
}

struct GcSmem {
  // byte offsets into dynamic shared memory (host computes, kernel mirrors)
  uint32_t slice0, slice1, scratch, gather, vecs, misc;
};

__host__ __device__ inline GcSmem gc_layout(uint32_t slice_bytes, int nslices, int C, int CS) {
  GcSmem L;
  uint32_t off = 0;
  L.slice0 = off; off += (slice_bytes + 127u) & ~127u;
  L.slice1 = off; if (nslices > 1) off += (slice_bytes + 127u) & ~127u;
  L.scratch = off; off += GC_THREADS * 8 * 4 * 2;           // [rows][C] x 2 floats  (rows * C = threads * V <= threads * 8)
  L.gather = off; off += (uint32_t)CS * 3 * C * 4;          // [CS][2][C] floats + [CS][R <= C] (backward dz1 partials)
  L.vecs = off; off += 8u * (uint32_t)C * 4;                // eight [C] float vectors
  L.misc = off; off += 1024;                                // stats gather [CS][2] doubles (<= 256 B), mbarriers, scalars
  return L;
}

// ------------------------------------------------------------------------------------------ forward
template <typename T>
__global__ void __launch_bounds__(GC_THREADS, 1)
    gn_se_cluster_fwd_kernel(const T* __restrict__ y, T* __restrict__ out, const float* __restrict__ gamma,
                             const float* __restrict__ beta, const float* __restrict__ w1, const float* __restrict__ b1,
                             const float* __restrict__ w2, const float* __restrict__ b2, int P, int C, int R,
                             float* __restrict__ stats, float* __restrict__ cpos_g, float* __restrict__ spos_g,
                             float* __restrict__ hidden_g, float* __restrict__ gate_g) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ __align__(128) uint8_t gsm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const bool has_se = w1 != nullptr;
  const int cv = C / V, rows = GC_THREADS / cv;
  const int pix = P / CS;                                   // pixels of this CTA
  const uint32_t slice_bytes = (uint32_t)pix * C * sizeof(T);
  const GcSmem L = gc_layout(slice_bytes, 1, C, CS);
  T* ys = reinterpret_cast<T*>(gsm + L.slice0);
  float* scratch = reinterpret_cast<float*>(gsm + L.scratch);
  float* gather = reinterpret_cast<float*>(gsm + L.gather);
  float* v_cpos = reinterpret_cast<float*>(gsm + L.vecs);
  float* v_spos = v_cpos + C;
  float* v_pool = v_spos + C;
  float* v_gate = v_pool + C;
  float* v_hid = v_gate + C;
  float* v_p0 = v_hid + C;
  float* v_p1 = v_p0 + C;
  double* st_gather = reinterpret_cast<double*>(gsm + L.misc);          // [CS][2]
  uint64_t* bar = reinterpret_cast<uint64_t*>(gsm + L.misc + 512);
  __shared__ double red[64];

  GC_TS(0);
  cluster.barrier_arrive();   // "this CTA is running"; waited on right before the first DSMEM access, free by then
  const T* src = y + ((size_t)b * P + (size_t)rank * pix) * C;
  if (threadIdx.x == 0) {
    gc_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    gc_load_slice(ys, src, slice_bytes, bar);
  }
  const int col = threadIdx.x % cv, row = threadIdx.x / cv;
  float gm[V], bt[V];      // loaded now: their latency hides behind the slice's bulk copy
#pragma unroll
  for (int i = 0; i < V; ++i) {
    gm[i] = gamma[col * V + i];
    bt[i] = beta[col * V + i];
  }
  __syncthreads();         // barrier initialised before anyone waits on it
  gc_mbar_wait(bar, 0);
  GC_TS(1);

  // ---- 1. statistics
  float s = 0.f, ss = 0.f;
  for (int p = row; p < pix; p += rows) {
    Vec16<T> v;
    v.load(ys + (size_t)p * C + col * V);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s += v.v[i];
      ss = fmaf(v.v[i], v.v[i], ss);
    }
  }
  double ds = (double)s, dss = (double)ss;
  gc_block_sum2(ds, dss, red);
  GC_TS(2);
  cluster.barrier_wait();  // every CTA of the cluster is running before the first DSMEM access
  if (threadIdx.x < CS) {  // all-gather: my partial into slot `rank` of every CTA
    double* dst = cluster.map_shared_rank(st_gather, threadIdx.x);
    dst[2 * rank] = ds;
    dst[2 * rank + 1] = dss;
  }
  cluster.sync();
  GC_TS(3);
  double ts = 0.0, tss = 0.0;
  for (int r = 0; r < CS; ++r) {
    ts += st_gather[2 * r];
    tss += st_gather[2 * r + 1];
  }
  const double count = (double)P * C;
  const double dmean = ts / count;
  double var = tss / count - dmean * dmean;
  if (var < 0.0) var = 0.0;
  const float mean = (float)dmean, rstd = (float)(1.0 / sqrt(var + (double)kGcEps));

  // z = (y - mean) * (rstd * gamma) + beta  is THE pre-activation in every pass (forward and backward, both kernel
  // families): one FADD + one FFMA per element, and the relu mask z > 0 is identical everywhere.
  float za[V], sc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    za[i] = rstd * gm[i];
    sc[i] = 1.f;
  }
  if (has_se) {
    // ---- 2. pooled sums over positive pixels
    float cnt[V], sx[V];
#pragma unroll
    for (int i = 0; i < V; ++i) cnt[i] = sx[i] = 0.f;
    for (int p = row; p < pix; p += rows) {
      Vec16<T> v;
      v.load(ys + (size_t)p * C + col * V);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float vm = v.v[i] - mean;
        const bool pos = fmaf(vm, za[i], bt[i]) > 0.f;
        cnt[i] += pos ? 1.f : 0.f;
        sx[i] += pos ? vm : 0.f;           // sum of (y - mean); scaled by rstd once below
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) sx[i] *= rstd;
    GC_TS(4);
    gc_channel_reduce<V>(cnt, sx, scratch, C, cv, rows, v_p0, v_p1);
    GC_TS(5);
    for (int r = 0; r < CS; ++r) {  // all-gather my [2][C] partials
      float* dst = cluster.map_shared_rank(gather, r) + (size_t)rank * 2 * C;
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        dst[c] = v_p0[c];
        dst[C + c] = v_p1[c];
      }
    }
    cluster.sync();
    GC_TS(6);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double a = 0.0, q = 0.0;
      for (int r = 0; r < CS; ++r) {
        a += (double)gather[(size_t)r * 2 * C + c];
        q += (double)gather[(size_t)r * 2 * C + C + c];
      }
      const float af = (float)a, qf = (float)q;
      v_cpos[c] = af;
      v_spos[c] = qf;
      v_pool[c] = (gamma[c] * qf + beta[c] * af) / (float)P;   // mean over pixels of relu(gn(y))
    }
    __syncthreads();
    GC_TS(7);
    // ---- 3. excitation.  Small SE blocks (C*R <= 4096 weights): every CTA computes all of it, no all-gather;
    //         otherwise hidden[j] for j = rank, rank + CS, ... and gate[c] for this CTA's channel slice, all-gathered.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const bool split = CS > 1 && (long)C * R > 4096;
    const int jstep = split ? CS : 1, j0 = split ? rank : 0;
    for (int j = j0 + warp * jstep; j < R; j += nwarp * jstep) {
      float acc = 0.f;
      const float* wr = w1 + (size_t)j * C;
#pragma unroll 8
      for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(wr + c), v_pool[c], acc);
      acc = warp_sum(acc);
      const float h = fmaxf(acc + b1[j], 0.f);
      if (split) {
        if (lane < CS) cluster.map_shared_rank(v_hid, lane)[j] = h;
      } else if (lane == 0) {
        v_hid[j] = h;
      }
    }
    if (split) cluster.sync(); else __syncthreads();
    GC_TS(8);
    const int cpr = split ? (C + CS - 1) / CS : C;
    const int c_lo = split ? rank * cpr : 0, c_hi = split ? min(C, (rank + 1) * cpr) : C;
    for (int c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
      float acc = b2[c];
      const float* wr = w2 + (size_t)c * R;
      if ((R & 3) == 0) {   // 16-byte loads of the thread's own W2 row, four independent chains
        float a4[4] = {0.f, 0.f, 0.f, 0.f};
        const float4* wq = reinterpret_cast<const float4*>(wr);
        const float4* hq = reinterpret_cast<const float4*>(v_hid);
#pragma unroll 16
        for (int j = 0; j < R / 4; ++j) {
          const float4 w4 = __ldg(wq + j), h4 = hq[j];
          a4[0] = fmaf(w4.x, h4.x, a4[0]);
          a4[1] = fmaf(w4.y, h4.y, a4[1]);
          a4[2] = fmaf(w4.z, h4.z, a4[2]);
          a4[3] = fmaf(w4.w, h4.w, a4[3]);
        }
        acc += (a4[0] + a4[1]) + (a4[2] + a4[3]);
      } else {
        for (int j = 0; j < R; ++j) acc = fmaf(wr[j], v_hid[j], acc);
      }
      const float gt = 1.0f / (1.0f + expf(-acc));
      if (split) {
        for (int r = 0; r < CS; ++r) cluster.map_shared_rank(v_gate, r)[c] = gt;
      } else {
        v_gate[c] = gt;
      }
    }
    if (split) cluster.sync(); else __syncthreads();
    GC_TS(9);
#pragma unroll
    for (int i = 0; i < V; ++i) sc[i] = v_gate[col * V + i];
  }
  // ---- 4. normalise + relu + rescale, one write
  T* dst = out + ((size_t)b * P + (size_t)rank * pix) * C;
  float zs[V], bs_[V];   // gate folded in: relu(z) * s = max(z * s, 0) for the sigmoid gate s > 0
#pragma unroll
  for (int i = 0; i < V; ++i) {
    zs[i] = za[i] * sc[i];
    bs_[i] = bt[i] * sc[i];
  }
  for (int p = row; p < pix; p += rows) {
    Vec16<T> v;
    v.load(ys + (size_t)p * C + col * V);
#pragma unroll
    for (int i = 0; i < V; ++i) v.v[i] = fmaxf(fmaf(v.v[i] - mean, zs[i], bs_[i]), 0.f);
    v.store(dst + (size_t)p * C + col * V);
  }
  GC_TS(10);
  // ---- 5. saved tensors
  if (rank == 0) {
    if (threadIdx.x == 0) {
      stats[2 * b] = mean;
      stats[2 * b + 1] = rstd;
    }
    if (has_se) {
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        cpos_g[(size_t)b * C + c] = v_cpos[c];
        spos_g[(size_t)b * C + c] = v_spos[c];
        gate_g[(size_t)b * C + c] = v_gate[c];
      }
      for (int j = threadIdx.x; j < R; j += blockDim.x) hidden_g[(size_t)b * R + j] = v_hid[j];
    }
  }
}

// ------------------------------------------------------------------------------------------ backward
// Notation as gn_se.cu::gn_bwd_mid_kernel.  DOUT_RESIDENT: the dout slice also lives in shared memory; otherwise it is
// streamed from global twice (second time from L2).
template <typename T, bool DOUT_RESIDENT>
__global__ void __launch_bounds__(GC_THREADS, 1)
    gn_se_cluster_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ y, T* __restrict__ dy,
                             const float* __restrict__ gamma, const float* __restrict__ beta,
                             const float* __restrict__ w1, const float* __restrict__ w2, int P, int C, int R,
                             const float* __restrict__ stats, const float* __restrict__ cpos_g,
                             const float* __restrict__ spos_g, const float* __restrict__ hidden_g,
                             const float* __restrict__ gate_g, float* __restrict__ G1o, float* __restrict__ G2o,
                             float* __restrict__ dz2o, float* __restrict__ dz1o) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ __align__(128) uint8_t gsm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const bool has_se = w1 != nullptr;
  const int cv = C / V, rows = GC_THREADS / cv;
  const int pix = P / CS;
  const uint32_t slice_bytes = (uint32_t)pix * C * sizeof(T);
  const GcSmem L = gc_layout(slice_bytes, DOUT_RESIDENT ? 2 : 1, C, CS);
  T* ys = reinterpret_cast<T*>(gsm + L.slice0);
  T* ds_ = reinterpret_cast<T*>(gsm + L.slice1);
  float* scratch = reinterpret_cast<float*>(gsm + L.scratch);
  float* gather = reinterpret_cast<float*>(gsm + L.gather);
  float* v_A1 = reinterpret_cast<float*>(gsm + L.vecs);
  float* v_A2 = v_A1 + C;
  float* v_dz2 = v_A2 + C;
  float* v_dz1 = v_dz2 + C;
  float* v_dp = v_dz1 + C;
  float* v_p0 = v_dp + C;
  float* v_p1 = v_p0 + C;
  float* v_gate = v_p1 + C;
  uint64_t* bar = reinterpret_cast<uint64_t*>(gsm + L.misc + 512);
  __shared__ double red[64];

  GC_TS(16);
  cluster.barrier_arrive();   // see forward
  const size_t goff = ((size_t)b * P + (size_t)rank * pix) * C;
  if (threadIdx.x == 0) {
    gc_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    gc_mbar_expect_tx(bar, DOUT_RESIDENT ? 2 * slice_bytes : slice_bytes);
    for (uint32_t off = 0; off < slice_bytes; off += 32768u) {
      const uint32_t n = slice_bytes - off < 32768u ? slice_bytes - off : 32768u;
      gc_bulk_load(reinterpret_cast<char*>(ys) + off, reinterpret_cast<const char*>(y + goff) + off, n, bar);
      if (DOUT_RESIDENT)
        gc_bulk_load(reinterpret_cast<char*>(ds_) + off, reinterpret_cast<const char*>(dout + goff) + off, n, bar);
    }
  }
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  const int col = threadIdx.x % cv, row = threadIdx.x / cv;
  float g[V], za[V], bt[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    g[i] = gamma[col * V + i];
    za[i] = rstd * g[i];
    bt[i] = beta[col * V + i];
  }
  // this thread's first element of the per-sample vectors phase 2 and 3 read, fetched now (latency under the bulk copy)
  float pre_cpos = 0.f, pre_spos = 0.f, pre_hid = 0.f;
  if (has_se) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) v_gate[c] = gate_g[(size_t)b * C + c];
    if ((int)threadIdx.x < C) {
      pre_cpos = cpos_g[(size_t)b * C + threadIdx.x];
      pre_spos = spos_g[(size_t)b * C + threadIdx.x];
    }
    if ((int)threadIdx.x < R) pre_hid = hidden_g[(size_t)b * R + threadIdx.x];
  }
  __syncthreads();         // barrier initialised, gate staged
  gc_mbar_wait(bar, 0);
  GC_TS(17);

  // ---- 1. A1[c] = sum dout*m, A2[c] = sum dout*m*xhat over this CTA's pixels
  float a1[V], a2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) a1[i] = a2[i] = 0.f;
  const T* dsrc = DOUT_RESIDENT ? ds_ : dout + goff;
  for (int p = row; p < pix; p += rows) {
    Vec16<T> v, d;
    v.load(ys + (size_t)p * C + col * V);
    d.load(dsrc + (size_t)p * C + col * V);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float vm = v.v[i] - mean;
      const float dm = (fmaf(vm, za[i], bt[i]) > 0.f) ? d.v[i] : 0.f;
      a1[i] += dm;
      a2[i] = fmaf(dm, vm, a2[i]);          // sum of dm * (y - mean); scaled by rstd once below
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) a2[i] *= rstd;
  GC_TS(18);
  gc_channel_reduce<V>(a1, a2, scratch, C, cv, rows, v_p0, v_p1);
  GC_TS(19);
  cluster.barrier_wait();  // every CTA of the cluster is running before the first DSMEM access
  for (int r = 0; r < CS; ++r) {
    float* dst = cluster.map_shared_rank(gather, r) + (size_t)rank * 2 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dst[c] = v_p0[c];
      dst[C + c] = v_p1[c];
    }
  }
  cluster.sync();
  GC_TS(20);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int r = 0; r < CS; ++r) {
      a += (double)gather[(size_t)r * 2 * C + c];
      q += (double)gather[(size_t)r * 2 * C + C + c];
    }
    const float af = (float)a, qf = (float)q;
    v_A1[c] = af;
    v_A2[c] = qf;
    if (has_se) {
      const double sg = (double)v_gate[c];
      const double dsv = (double)gamma[c] * q + (double)beta[c] * a;   // the two terms cancel: keep the double sums
      v_dz2[c] = (float)(dsv * sg * (1.0 - sg));
    }
  }
  __syncthreads();
  GC_TS(21);
  // ---- 2. SE backward, split across the cluster
  if (has_se) {
    // dz1[j] = (hidden > 0) * sum_c dz2[c] W2[c][j].  W2 is [C][R]: threads run along j (coalesced rows), channel
    // groups in parallel.  Large SE blocks split the channel range across the cluster and all-gather the partial sums;
    // small ones (C*R <= 4096) are done whole by every CTA.
    const bool split = CS > 1 && (long)C * R > 4096;
    {
      const int cpr = split ? (C + CS - 1) / CS : C;
      const int c_lo = split ? rank * cpr : 0, c_hi = split ? min(C, (rank + 1) * cpr) : C;
      // channel groups that share a warp (R < 32) combine by an xor butterfly, the rest through shared memory: the
      // former one-thread-per-j loop over up to 64 group partials was a serial chain of double adds
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (int)blockDim.x >> 5;
      int RP = 1;
      while (RP < R) RP <<= 1;                               // threads per channel group
      const int ng = max(1, (int)blockDim.x / RP);
      const int jj = threadIdx.x % RP, cgp = threadIdx.x / RP;
      float part = 0.f;
      if (cgp < ng && jj < R) {
        const float* wp = w2 + jj;
#pragma unroll 8
        for (int c = c_lo + cgp; c < c_hi; c += ng) part = fmaf(v_dz2[c], __ldg(wp + (size_t)c * R), part);
      }
      int nparts;
      if (RP < 32) {
        for (int o = RP; o < 32; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        nparts = nwarp;
        if (lane < RP && jj < R) scratch[warp * R + jj] = part;
      } else {
        nparts = ng;
        if (cgp < ng && jj < R) scratch[cgp * R + jj] = part;
      }
      __syncthreads();
      if ((int)threadIdx.x < R) {
        double acc = 0.0;   // signed terms that cancel: combine the partials in double
        for (int q = 0; q < nparts; ++q) acc += (double)scratch[q * R + threadIdx.x];
        if (split) {
          const float af = (float)acc;
          for (int r = 0; r < CS; ++r) cluster.map_shared_rank(gather, r)[(size_t)CS * 2 * C + rank * R + threadIdx.x] = af;
        } else {
          v_dz1[threadIdx.x] = pre_hid > 0.f ? (float)acc : 0.f;
        }
      }
      if (split) {
        cluster.sync();
        if ((int)threadIdx.x < R) {
          double acc = 0.0;
          for (int r = 0; r < CS; ++r) acc += (double)gather[(size_t)CS * 2 * C + r * R + threadIdx.x];
          v_dz1[threadIdx.x] = pre_hid > 0.f ? (float)acc : 0.f;
        }
      }
      __syncthreads();
    }
    GC_TS(22);
    {
      const int cpr = split ? (C + CS - 1) / CS : C;
      const int c_lo = split ? rank * cpr : 0, c_hi = split ? min(C, (rank + 1) * cpr) : C;
      for (int c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
        float d4[4] = {0.f, 0.f, 0.f, 0.f};   // sixteen loads in flight, four chains
        const float* wc = w1 + c;
        int j = 0;
#pragma unroll 4
        for (; j + 3 < R; j += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) d4[u] = fmaf(v_dz1[j + u], __ldg(wc + (size_t)(j + u) * C), d4[u]);
        }
        for (; j < R; ++j) d4[0] = fmaf(v_dz1[j], __ldg(wc + (size_t)j * C), d4[0]);
        float dp = (d4[0] + d4[1]) + (d4[2] + d4[3]);
        dp /= (float)P;
        if (split) {
          for (int r = 0; r < CS; ++r) cluster.map_shared_rank(v_dp, r)[c] = dp;
        } else {
          v_dp[c] = dp;
        }
      }
      if (split) cluster.sync(); else __syncthreads();
    }
  }
  GC_TS(23);
  // ---- 3. G1, G2 per channel, S1, S2 per sample (every CTA computes the same values in the same order)
  double s1 = 0.0, s2 = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float g1, g2;
    if (has_se) {
      const float sg = v_gate[c], dp = v_dp[c];
      const bool first = c == (int)threadIdx.x;
      g1 = fmaf(dp, first ? pre_cpos : cpos_g[(size_t)b * C + c], sg * v_A1[c]);
      g2 = fmaf(dp, first ? pre_spos : spos_g[(size_t)b * C + c], sg * v_A2[c]);
    } else {
      g1 = v_A1[c];
      g2 = v_A2[c];
    }
    if (rank == 0) {
      G1o[(size_t)b * C + c] = g1;
      G2o[(size_t)b * C + c] = g2;
      if (has_se) dz2o[(size_t)b * C + c] = v_dz2[c];
    }
    s1 += (double)gamma[c] * (double)g1;
    s2 += (double)gamma[c] * (double)g2;
  }
  if (rank == 0 && has_se)
    for (int j = threadIdx.x; j < R; j += blockDim.x) dz1o[(size_t)b * R + j] = v_dz1[j];
  gc_block_sum2(s1, s2, red);
  const float invM = 1.0f / ((float)P * (float)C);
  const float S1 = (float)s1 * invM, S2 = (float)s2 * invM;
  GC_TS(24);
  // ---- 4. dy
  // dy = rstd * (gg * gamma - S1 - xhat * S2),  gg = mask ? dout * gate + dpool : 0
  //    = (mask ? dout * c1 + c2 : 0) + (y - mean) * k2 + k1      with the per-channel / per-sample constants below
  float c1[V], c2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float rg = rstd * g[i];
    c1[i] = rg * (has_se ? v_gate[col * V + i] : 1.f);
    c2[i] = rg * (has_se ? v_dp[col * V + i] : 0.f);
  }
  const float k1 = -rstd * S1, k2 = -rstd * rstd * S2;
  T* dstp = dy + goff;
  for (int p = row; p < pix; p += rows) {
    Vec16<T> v, d;
    v.load(ys + (size_t)p * C + col * V);
    d.load(dsrc + (size_t)p * C + col * V);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float vm = v.v[i] - mean;
      const float t = (fmaf(vm, za[i], bt[i]) > 0.f) ? fmaf(d.v[i], c1[i], c2[i]) : 0.f;
      d.v[i] = t + fmaf(vm, k2, k1);
    }
    d.store(dstp + (size_t)p * C + col * V);
  }
  GC_TS(25);
}

// ------------------------------------------------------------------------------------------ host
static constexpr size_t GC_SMEM_MAX = 226 * 1024;  // 227 KB minus the static reduction scratch

static bool gc_env_off() { return getenv("BVAE_GN_NO_CLUSTER") != nullptr; }

// cluster size for a sample with `nslices` resident slices per CTA; 0 = not eligible
static int gc_pick_cs(int B, int P, int C, int vec, size_t elem, int nslices, size_t* smem_out) {
  if (gc_env_off()) return 0;
  if (C % vec) return 0;
  const int cv = C / vec;
  if (cv > GC_THREADS || GC_THREADS % cv || C > 2048) return 0;
  auto fits = [&](int cs, size_t* smem) {
    if (P % cs) return false;
    const size_t slice = (size_t)(P / cs) * C * elem;
    if (slice % 16 || slice * nslices >= (1u << 20)) return false;   // bulk-copy granularity, mbarrier tx-count limit
    const GcSmem L = gc_layout((uint32_t)slice, nslices, C, cs);
    *smem = L.misc + 1024;
    return *smem <= GC_SMEM_MAX;
  };
  int cs = 1;
  size_t smem = 0;
  // 16-CTA clusters (2 MB samples) measured slower than the multi-pass kernels: one CTA per SM with its
  // load / compute / store phases in lock-step leaves HBM idle half the time, and few 16-SM clusters fit a GPC
  int max_cs = 8;
  if (const char* e = getenv("BVAE_GC_MAXCS")) max_cs = atoi(e);
  while (cs <= max_cs && !fits(cs, &smem)) cs <<= 1;
  if (cs > max_cs) return 0;
  // widen while most SMs would idle and the slices stay large: every doubling adds cluster-barrier latency to five
  // all-gathers, measured best down to 16 KB slices (the SE mat-vecs and per-channel reductions also split across the cluster)
  const int sms = num_sms();
  size_t min_slice = 16384;
  if (const char* e = getenv("BVAE_GC_MIN_SLICE")) min_slice = (size_t)atol(e);
  while ((long)B * cs < (sms * 3) / 4 && cs < 8) {   // (8-CTA clusters at two CTAs per SM measured 2x slower)
    size_t s2 = 0;
    if (!fits(cs * 2, &s2) || (size_t)(P / (cs * 2)) * C * elem < min_slice) break;
    cs *= 2;
    smem = s2;
  }
  *smem_out = smem;
  return cs;
}

template <typename K, typename... Args>
static int gc_launch(K kern, int CS, int B, size_t smem, cudaStream_t st, Args... args) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GC_SMEM_MAX);
  if (e != cudaSuccess) return cuda_fail(e, "gn_se_cluster: cudaFuncSetAttribute(smem)");
  if (CS > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return cuda_fail(e, "gn_se_cluster: cudaFuncSetAttribute(cluster 16)");
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(CS, B, 1);
  cfg.blockDim = dim3(GC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) return cuda_fail(e, "gn_se_cluster: launch");
  ++g_launches;
  return BVAE_OK;
}

bool gn_cluster_eligible(int B, int P, int C, int R, int elem_bytes) {
  if (R > C) return false;
  size_t smem = 0;
  return gc_pick_cs(B, P, C, 16 / elem_bytes, (size_t)elem_bytes, 1, &smem) != 0;
}

template <typename T>
int gn_cluster_fwd(const T* y, T* out, const float* gamma, const float* beta, const float* w1, const float* b1,
                   const float* w2, const float* b2, int B, int P, int C, int R, float* stats, float* cpos, float* spos,
                   float* hidden, float* gate, cudaStream_t st, bool* handled) {
  *handled = false;
  size_t smem = 0;
  if (w1 && R > C) return BVAE_OK;
  const int CS = gc_pick_cs(B, P, C, Vec16<T>::N, sizeof(T), 1, &smem);
  if (!CS) return BVAE_OK;
  *handled = true;
  return gc_launch(gn_se_cluster_fwd_kernel<T>, CS, B, smem, st, y, out, gamma, beta, w1, b1, w2, b2, P, C, R, stats, cpos,
                   spos, hidden, gate);
}

template <typename T>
int gn_cluster_bwd(const T* dout, const T* y, T* dy, const float* gamma, const float* beta, const float* w1,
                   const float* w2, int B, int P, int C, int R, const float* stats, const float* cpos, const float* spos,
                   const float* hidden, const float* gate, float* G1, float* G2, float* dz2, float* dz1, cudaStream_t st,
                   bool* handled) {
  *handled = false;
  if (w1 && R > C) return BVAE_OK;
  size_t smem2 = 0, smem1 = 0;
  const int cs2 = gc_pick_cs(B, P, C, Vec16<T>::N, sizeof(T), 2, &smem2);
  const int cs1 = gc_pick_cs(B, P, C, Vec16<T>::N, sizeof(T), 1, &smem1);
  // both slices resident only when two CTAs still fit an SM (one loads while the other computes)
  const bool prefer2 = cs2 && smem2 <= 113 * 1024 && (!cs1 || cs2 <= cs1);
  if (prefer2 || (cs2 && !cs1)) {
    *handled = true;
    return gc_launch(gn_se_cluster_bwd_kernel<T, true>, cs2, B, smem2, st, dout, y, dy, gamma, beta, w1, w2, P, C, R, stats,
                     cpos, spos, hidden, gate, G1, G2, dz2, dz1);
  }
  if (cs1) {
    *handled = true;
    return gc_launch(gn_se_cluster_bwd_kernel<T, false>, cs1, B, smem1, st, dout, y, dy, gamma, beta, w1, w2, P, C, R,
                     stats, cpos, spos, hidden, gate, G1, G2, dz2, dz1);
  }
  return BVAE_OK;
}

template int gn_cluster_fwd<float>(const float*, float*, const float*, const float*, const float*, const float*,
                                   const float*, const float*, int, int, int, int, float*, float*, float*, float*, float*,
                                   cudaStream_t, bool*);
template int gn_cluster_fwd<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, const float*, const float*, const float*,
                                           const float*, const float*, const float*, int, int, int, int, float*, float*,
                                           float*, float*, float*, cudaStream_t, bool*);
template int gn_cluster_bwd<float>(const float*, const float*, float*, const float*, const float*, const float*,
                                   const float*, int, int, int, int, const float*, const float*, const float*,
                                   const float*, const float*, float*, float*, float*, float*, cudaStream_t, bool*);
template int gn_cluster_bwd<__nv_bfloat16>(const __nv_bfloat16*, const __nv_bfloat16*, __nv_bfloat16*, const float*,
                                           const float*, const float*, const float*, int, int, int, int, const float*,
                                           const float*, const float*, const float*, const float*, float*, float*, float*,
                                           float*, cudaStream_t, bool*);

}  // namespace bvae

#ifdef BVAE_GC_TIMING
extern "C" int bvae_gc_timestamps(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, bvae::g_gc_ts, sizeof(long long) * 32);
}
#endif
