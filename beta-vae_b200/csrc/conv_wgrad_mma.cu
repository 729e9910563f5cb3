// Weight gradient of the 3x3 convolutions for SMALL channel counts (16 / 32 / 64 in, 16 / 32 / 64 out, not both 64):
//   dW[t][ci][co] = sum over output pixels p of  x[p*stride + off_t][ci] * dy[p][co]
// (what convolution_backward computes for nn.Conv2d at src/models/beta_vae.py:37,42,54 on the narrow layers of
//  configs/overfit_capacity.yaml: base_channels 16, 256x256 images).
//
// The tcgen05 wgrad (conv_tc_wgrad.cu) wants 64-channel MN-major boxes on both operands, and the CUDA-core tile kernel
// (conv_simt.cu) computes a 64 x 64 output tile per CTA, of which a 16 x 16 layer uses 1/16: the five narrow wgrads
// were 72 % of that config's step (490 us each).  These shapes are far below the tensor-core roofline -- 2304 MACs per
// pixel against 64 bytes of bf16 input at 16 x 16 -- so the warp-level MMA path is enough to make them memory bound,
// and it reads its operands from arbitrary shared-memory rows, which is what the shifted tap windows need:
//   * a CTA walks (16 x 8)-pixel output tiles; per tile the x halo ((18 x 10) pixels at stride 1, (33 x 17) at stride 2)
//     and the dy tile are staged in shared memory once (16-byte loads, zero-filled outside the image = the padding);
//   * K runs over the 128 pixels of the tile (eight k16 steps of two tile rows); A = x^T and B = dy both have the
//     contraction index (the pixel) as their slow dimension, so both fragments come from ldmatrix.trans, whose eight
//     row addresses per matrix are simply eight consecutive output pixels of a tile row, shifted by the tap (and
//     two input pixels apart at stride 2);
//   * the 9 * Cin/16 (tap, 16-channel block) work items are dealt round-robin to the eight warps, each item owning
//     all Cout/8 accumulator tiles; one warp also multiplies an all-ones A fragment: the bias gradient;
//   * split-K over tiles: every CTA writes one fp32 partial [9][Cin][Cout] (+ [Cout]) that the deterministic reduce of
//     conv_simt.cu sums.
#include "common.cuh"
#include "conv_tc.cuh"

namespace bvae {

static constexpr int WS_TH = 16, WS_TW = 8;   // output tile

__device__ __forceinline__ void wsm_ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void wsm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct WsParams {
  const __nv_bfloat16* x;    // [B][H][W][Cin]
  const __nv_bfloat16* dy;   // [B][Ho][Wo][Cout]
  float* part;               // [split][9][Cin][Cout]
  float* bpart;              // [split][Cout] or null
  int B, H, W, Cin, Ho, Wo, Cout;
  int tiles_w, tiles_h, tiles, tiles_per_cta;
};

// ITEMS = ceil(9 * (Cin/16) / 8) work items per warp, NT = Cout / 8 accumulator tiles per item
template <int ITEMS, int NT, int STRIDE>
__global__ void __launch_bounds__(256) conv_wgrad_mma_small_kernel(const WsParams p) {
  extern __shared__ __align__(16) uint8_t ws_raw[];
  constexpr int HH = STRIDE == 1 ? WS_TH + 2 : 2 * WS_TH + 1;   // halo rows / columns (input pixels)
  constexpr int HW = STRIDE == 1 ? WS_TW + 2 : 2 * WS_TW + 1;
  const int xp = p.Cin * 2 + 16, dp = p.Cout * 2 + 16;          // padded pixel pitches: conflict-free ldmatrix rows
  uint8_t* xs = ws_raw;                                         // [HH * HW][xp]
  uint8_t* ds = ws_raw + (size_t)HH * HW * xp;                  // [128][dp]
  const uint32_t xs_s = (uint32_t)__cvta_generic_to_shared(xs), ds_s = (uint32_t)__cvta_generic_to_shared(ds);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int MT = p.Cin >> 4;
  const int nitems = 9 * MT;
  const int cvx = p.Cin >> 3, cvd = p.Cout >> 3;                // 16-byte vectors per pixel

  float acc[ITEMS][NT][4];
  float accb[NT][4];
#pragma unroll
  for (int u = 0; u < ITEMS; ++u)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[u][j][e] = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) accb[j][e] = 0.f;
  const bool bias_warp = (warp == 7) && p.bpart != nullptr;     // warp 7 has the fewest items
  const uint32_t ones[4] = {0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u};

  // per-lane ldmatrix row selectors.  A: matrix bit0 -> channel half, bit1 -> pixel (k) half; B: bit0 -> k half, bit1 -> n half
  const int li = lane & 7, lm = lane >> 3;
  const int a_ch = (lm & 1) * 8, a_kh = lm >> 1;
  const int b_kh = lm & 1, b_nh = (lm >> 1) * 8;

  const int t_beg = blockIdx.x * p.tiles_per_cta, t_end = min(p.tiles, t_beg + p.tiles_per_cta);
  for (int tile = t_beg; tile < t_end; ++tile) {
    int tt = tile;
    const int ow0 = (tt % p.tiles_w) * WS_TW;
    tt /= p.tiles_w;
    const int oh0 = (tt % p.tiles_h) * WS_TH;
    const int b = tt / p.tiles_h;
    const int ih0 = oh0 * STRIDE - 1, iw0 = ow0 * STRIDE - 1;   // halo origin in the input
    __syncthreads();                                             // previous tile's fragments are read
    // ---- stage the x halo and the dy tile: batches of four independent 16-byte loads per thread
    const int nx = HH * HW * cvx;
    for (int base = 0; base < nx; base += 256 * 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 256 + tid;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (i < nx) {
          const int px = i / cvx, cq = i - px * cvx;
          const int hy = px / HW, hx = px - hy * HW;
          const int ih = ih0 + hy, iw = iw0 + hx;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
            v[u] = *reinterpret_cast<const uint4*>(p.x + (((size_t)b * p.H + ih) * p.W + iw) * p.Cin + cq * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 256 + tid;
        if (i < nx) {
          const int px = i / cvx, cq = i - px * cvx;
          *reinterpret_cast<uint4*>(xs + (size_t)px * xp + cq * 16) = v[u];
        }
      }
    }
    const int nd = 128 * cvd;
    for (int base = 0; base < nd; base += 256 * 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 256 + tid;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (i < nd) {
          const int px = i / cvd, cq = i - px * cvd;
          const int oy = px >> 3, ox = px & 7;
          v[u] = *reinterpret_cast<const uint4*>(p.dy + (((size_t)b * p.Ho + oh0 + oy) * p.Wo + ow0 + ox) * p.Cout + cq * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 256 + tid;
        if (i < nd) {
          const int px = i / cvd, cq = i - px * cvd;
          *reinterpret_cast<uint4*>(ds + (size_t)px * dp + cq * 16) = v[u];
        }
      }
    }
    __syncthreads();
    // ---- eight k16 steps: output rows 2ks, 2ks+1 of the tile
#pragma unroll 1
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t bf[NT][2];
#pragma unroll
      for (int nb = 0; nb < NT / 2; ++nb) {
        uint32_t r[4];
        wsm_ldmatrix_x4_trans(r, ds_s + (uint32_t)(((2 * ks + b_kh) * 8 + li) * dp + (nb * 16 + b_nh) * 2));
        bf[2 * nb][0] = r[0]; bf[2 * nb][1] = r[1]; bf[2 * nb + 1][0] = r[2]; bf[2 * nb + 1][1] = r[3];
      }
#pragma unroll
      for (int u = 0; u < ITEMS; ++u) {
        const int q = warp + 8 * u;
        if (q < nitems) {
          const int tap = q / MT, mt = q - tap * MT;
          const int r3 = tap / 3, s3 = tap - r3 * 3;
          const int oy = 2 * ks + a_kh;
          const int hidx = (oy * STRIDE + r3) * HW + li * STRIDE + s3;
          uint32_t a[4];
          wsm_ldmatrix_x4_trans(a, xs_s + (uint32_t)(hidx * xp + (mt * 16 + a_ch) * 2));
#pragma unroll
          for (int j = 0; j < NT; ++j) wsm_mma(acc[u][j], a, bf[j][0], bf[j][1]);
        }
      }
      if (bias_warp) {
#pragma unroll
        for (int j = 0; j < NT; ++j) wsm_mma(accb[j], ones, bf[j][0], bf[j][1]);
      }
    }
  }
  // ---- partials: C fragment = (ci = 16 mt + g [+8], co = 8 j + 2 t4, +1)
  const int g = lane >> 2, t4 = lane & 3;
  float* o = p.part + (size_t)blockIdx.x * 9 * p.Cin * p.Cout;
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    const int q = warp + 8 * u;
    if (q >= nitems) continue;
    const int tap = q / MT, mt = q - tap * MT;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      float* row0 = o + ((size_t)tap * p.Cin + mt * 16 + g) * p.Cout + j * 8 + 2 * t4;
      *reinterpret_cast<float2*>(row0) = make_float2(acc[u][j][0], acc[u][j][1]);
      *reinterpret_cast<float2*>(row0 + (size_t)8 * p.Cout) = make_float2(acc[u][j][2], acc[u][j][3]);
    }
  }
  if (bias_warp && g == 0) {
#pragma unroll
    for (int j = 0; j < NT; ++j)
      *reinterpret_cast<float2*>(p.bpart + (size_t)blockIdx.x * p.Cout + j * 8 + 2 * t4) = make_float2(accb[j][0], accb[j][1]);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
bool conv_wgrad_mma_small_supported(int H, int W, int Cin, int Cout, int stride) {
  if (getenv("BVAE_CONV_NO_HALO")) return false;
  if (stride != 1 && stride != 2) return false;
  if (Cin % 16 || Cout % 16 || Cin > 64 || Cout > 64 || (Cin == 64 && Cout == 64) || Cin == 48 || Cout == 48) return false;
  if (stride == 2 && ((H | W) & 1)) return false;
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  return Ho % WS_TH == 0 && Wo % WS_TW == 0;
}

WgradTcPlan conv_wgrad_mma_small_plan(int B, int H, int W, int Cin, int Cout, int stride) {
  WgradTcPlan pl;
  memset(&pl, 0, sizeof(pl));
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  pl.tiles_total = B * (Ho / WS_TH) * (Wo / WS_TW);
  int ctas = 2 * num_sms();
  if (ctas > pl.tiles_total) ctas = pl.tiles_total;
  pl.tiles_per_split = (pl.tiles_total + ctas - 1) / ctas;
  pl.splits = (pl.tiles_total + pl.tiles_per_split - 1) / pl.tiles_per_split;
  pl.bias_chunks = pl.splits;
  pl.part_bytes = align_up((size_t)pl.splits * 9 * Cin * Cout * sizeof(float), 256);
  pl.bpart_bytes = align_up((size_t)pl.splits * Cout * sizeof(float), 256);
  return pl;
}

template <int ITEMS, int NT>
static int ws_launch(const WsParams& p, int stride, int grid, size_t smem, cudaStream_t st) {
  auto kern = stride == 1 ? conv_wgrad_mma_small_kernel<ITEMS, NT, 1> : conv_wgrad_mma_small_kernel<ITEMS, NT, 2>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "conv_wgrad_mma_small: cudaFuncSetAttribute");
  }
  kern<<<grid, 256, smem, st>>>(p);
  BVAE_LAUNCH_CHECK("conv_wgrad_mma_small");
  return BVAE_OK;
}

int conv_wgrad_mma_small(const void* x, const void* dy, float* part, float* bpart, const WgradTcPlan& pl, int B, int H, int W,
                         int Cin, int Cout, int stride, cudaStream_t st) {
  WsParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy);
  p.part = part;
  p.bpart = bpart;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.Ho = (H - 1) / stride + 1;
  p.Wo = (W - 1) / stride + 1;
  p.tiles_w = p.Wo / WS_TW;
  p.tiles_h = p.Ho / WS_TH;
  p.tiles = pl.tiles_total;
  p.tiles_per_cta = pl.tiles_per_split;
  const int hh = stride == 1 ? WS_TH + 2 : 2 * WS_TH + 1, hw = stride == 1 ? WS_TW + 2 : 2 * WS_TW + 1;
  const size_t smem = (size_t)hh * hw * (Cin * 2 + 16) + (size_t)128 * (Cout * 2 + 16);
  const int mt = Cin / 16, nt = Cout / 8;
  const int grid = pl.splits;
#define WS_CASE(M, N, I)                                   \
  if (mt == (M) && nt == (N)) return ws_launch<I, N>(p, stride, grid, smem, st);
  WS_CASE(1, 2, 2) WS_CASE(1, 4, 2) WS_CASE(1, 8, 2)
  WS_CASE(2, 2, 3) WS_CASE(2, 4, 3) WS_CASE(2, 8, 3)
  WS_CASE(4, 2, 5) WS_CASE(4, 4, 5)
#undef WS_CASE
  BVAE_REQUIRE(false, BVAE_ERR_UNSUPPORTED, "conv_wgrad_mma_small: %d -> %d channels", Cin, Cout);
}

}  // namespace bvae
