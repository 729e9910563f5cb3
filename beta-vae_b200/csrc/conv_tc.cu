// 3x3 convolution on the 5th-generation tensor cores: implicit GEMM with tcgen05.mma,
// accumulators in TMEM, operands staged in shared memory by TMA (bf16 in, fp32 accumulate).
// Covers fprop (stride 1 | 2) and dgrad (stride 1 | 2) of nn.Conv2d(3x3, pad 1)
// (reference: src/models/beta_vae.py:37,42,54 -> cuDNN fprop/dgrad behind ATen conv2d).
//
// One "tap GEMM" serves all four cases (same formulation as conv_simt.cu):
//   out[pixel, n] = bias[n] + sum_t sum_k SRC_{map_t}[pixel + (di_t, dj_t), k] * W[widx_t][n][k]
//   * GEMM rows M = 128 output pixels of one tile (tw x th x tn box of the NHWC tensor),
//     columns N = BN output channels, reduction = taps x channel chunks of KC (<= 64) elements.
//   * A operand: one TMA box per (tap, chunk): the tile's pixels shifted by the tap offset;
//     out-of-range coordinates (the conv's zero padding, including w = -1 / h = -1) are zero-filled by
//     the TMA unit, so there is no im2col buffer and no boundary code in the kernel.
//     Stride-2 fprop reads through four "parity" tensor maps (even/odd rows x even/odd columns of
//     x, each a plain strided view), stride-2 dgrad runs as four dense sub-convolutions (one per
//     output parity class) whose epilogue scatters to (2i+po, 2j+pw).
//   * B operand: weights pre-packed as [tap][n][k] bf16 (K-major), one TMA box per (tap, chunk).
//   * both operands land in the 128B/64B/32B-swizzled K-major layout tcgen05 reads directly.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM -> registers -> +bias -> bf16 -> global).  A ring of STAGES
// shared-memory slots is handed back and forth with mbarriers (full: TMA complete_tx;
// empty: tcgen05.commit); the epilogue is released by a final tcgen05.commit.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace bvae {

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
static constexpr int TC_THREADS = 192;
static constexpr int TC_BM = 128;

__global__ void __launch_bounds__(TC_THREADS, 1)
    conv_tap_gemm_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned operand ring (swizzle-128B atoms), barriers behind it
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = TC_BM * p.kc * 2, b_bytes = p.bn * p.kc * 2;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023) & ~1023u;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* accum_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  // tile coordinates
  int tile = blockIdx.x;
  const int tw_i = tile % p.tiles_w;
  tile /= p.tiles_w;
  const int th_i = tile % p.tiles_h;
  const int tn_i = tile / p.tiles_h;
  const int w0 = tw_i * p.tw, h0 = th_i * p.th, b0 = tn_i * p.tn;
  const int n0 = blockIdx.y * p.bn;
  const int cls = blockIdx.z;                      // stride-2 dgrad: output parity class (po, pw) = (cls / 2, cls % 2)
  const int iters = p.ntaps[cls] * p.kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (all lanes run the loop, the elected lane issues; see elect_one() in tc_ptx.cuh) =====
    const bool leader = elect_one();
    if (leader) tma_prefetch_desc(&maps.w);
    int s = 0, t = 0, kc = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty_bar[s], ph ^ 1u);
      const TcTap tap = p.taps[cls][t];
      uint8_t* a_dst = smem + (size_t)s * stage_bytes;
      uint8_t* b_dst = a_dst + a_bytes;
      if (leader) {
        mbar_expect_tx(&full_bar[s], a_bytes + b_bytes);
        tma_load_4d(&maps.src[tap.map], &full_bar[s], a_dst, kc * p.kc, w0 + tap.dj, h0 + tap.di, b0);
        tma_load_3d(&maps.w, &full_bar[s], b_dst, kc * p.kc, n0, tap.widx);
      }
      __syncwarp();
      if (++kc == p.kchunks) { kc = 0; ++t; }
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (all lanes run the loop, the elected lane issues) =====
    const bool leader = elect_one();
    // instruction descriptor: D fp32, A/B bf16, both K-major, N = bn, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t row_bytes = p.kc * 2;
    int s = 0;
    uint32_t ph = 0;
    const uint32_t smem0 = smem_u32(smem);
    const uint64_t desc_hi = kmajor_desc(0, row_bytes);     // everything but the address field
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t a_addr = smem0 + (uint32_t)s * stage_bytes;
      const uint64_t adesc = desc_hi + (uint64_t)(a_addr >> 4), bdesc = desc_hi + (uint64_t)((a_addr + a_bytes) >> 4);
      if (leader) {
        // advance 16 elements (32 bytes) along K inside the swizzle span: +2 in 16-byte units
        umma_bf16(tmem_base, adesc, bdesc, idesc, it != 0 ? 1u : 0u);
        if (p.kc >= 32) umma_bf16(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
        if (p.kc >= 64) {
          umma_bf16(tmem_base, adesc + 4, bdesc + 4, idesc, 1u);
          umma_bf16(tmem_base, adesc + 6, bdesc + 6, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);  // slot reusable once these MMAs have read it
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
    if (leader) umma_commit(accum_bar);        // accumulator complete
    __syncwarp();
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;  // GEMM row = pixel index inside the tile
    const int wl = r % p.tw, hl = (r / p.tw) % p.th, nl = r / (p.tw * p.th);
    const int b = b0 + nl, oi = h0 + hl, oj = w0 + wl;
    const bool valid = b < p.B;
    const int po = p.ncls == 4 ? (cls >> 1) : 0, pw = p.ncls == 4 ? (cls & 1) : 0;
    __nv_bfloat16* out = p.dst + (((size_t)b * p.DH + (size_t)(oi * p.ds + po)) * p.DW + (size_t)(oj * p.ds + pw)) * p.N + n0;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c = 0; c < p.bn; c += 32) {
      uint32_t v[32];
      const int width = (p.bn - c) >= 32 ? 32 : 16;  // bn is a multiple of 16
      if (width == 32)
        tmem_ld32(taddr + (uint32_t)c, v);
      else
        tmem_ld16(taddr + (uint32_t)c, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g * 8 < width) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float f0 = __uint_as_float(v[g * 8 + 2 * e]), f1 = __uint_as_float(v[g * 8 + 2 * e + 1]);
              if (p.bias) {
                f0 += __ldg(p.bias + n0 + c + g * 8 + 2 * e);
                f1 += __ldg(p.bias + n0 + c + g * 8 + 2 * e + 1);
              }
              __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
              pk[e] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(out + c + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int pick_kc(int K) { return (K % 64 == 0) ? 64 : (K == 32 ? 32 : (K == 16 ? 16 : 0)); }
static int pick_bn(int N) {
  for (int bn = 256; bn >= 16; bn >>= 1)
    if (N % bn == 0) return bn;
  return 0;
}

bool conv_tc_supported(int H, int W, int Cin, int Cout, int stride, int kind) {
  // kind 0: fprop (K = Cin, N = Cout, grid = output);  kind 1: dgrad (K = Cout, N = Cin, grid = input)
  const int K = kind == 0 ? Cin : Cout, N = kind == 0 ? Cout : Cin;
  if (pick_kc(K) == 0 || pick_bn(N) == 0) return false;
  if (stride == 2 && ((H | W) & 1)) return false;
  int OW, OH;
  if (kind == 0) {
    OW = (W - 1) / stride + 1;
    OH = (H - 1) / stride + 1;
  } else {
    OW = W / stride;  // stride 2: each parity class is an (H/2 x W/2) grid
    OH = H / stride;
  }
  int tw, th, tn;
  return tile_geom(OW, OH, &tw, &th, &tn);
}

static int launch_tc(const TcMaps& maps, TcParams& p, int m_tiles, cudaStream_t st) {
  const uint32_t a_bytes = TC_BM * p.kc * 2, b_bytes = p.bn * p.kc * 2;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023) & ~1023u;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > 6) stages = 6;
  if (stages < 2) stages = 2;
  p.stages = stages;
  int cols = 32;
  while (cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tap_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "conv_tc: cudaFuncSetAttribute");
    configured = smem;
  }
  dim3 grid(m_tiles, p.N / p.bn, p.ncls);
  conv_tap_gemm_tc_kernel<<<grid, TC_THREADS, smem, st>>>(maps, p);
  BVAE_LAUNCH_CHECK("conv_tap_gemm_tc");
  return BVAE_OK;
}

// weights packed [9][N][K] bf16
static int make_weight_map(CUtensorMap* m, const void* wp, int K, int N, int kc, int bn) {
  const uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, 9};
  const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)K * N * 2};
  const uint32_t box[3] = {(uint32_t)kc, (uint32_t)bn, 1};
  return make_map(m, wp, 3, dims, strides, box, kc * 2);
}

int conv_tc_fprop(const void* x, const void* wp, const float* bias, void* y, int B, int H, int W, int Cin, int Cout,
                  int stride, cudaStream_t st, double* stat_part) {
  if (stride == 1 && conv_halo_supported(H, W, Cin, Cout))
    return conv_halo_launch(x, wp, bias, y, B, H, W, Cin, Cout, 0, st, stat_part);
  BVAE_REQUIRE(stat_part == nullptr, BVAE_ERR_UNSUPPORTED, "conv_tc_fprop: statistics output needs the stride-1 halo kernel");
  TcMaps maps;
  TcParams p;
  memset(&p, 0, sizeof(p));
  const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
  p.kc = pick_kc(Cin);
  p.bn = pick_bn(Cout);
  p.kchunks = Cin / p.kc;
  BVAE_REQUIRE(tile_geom(OW, OH, &p.tw, &p.th, &p.tn), BVAE_ERR_UNSUPPORTED, "conv_tc_fprop: tile geometry");
  p.tiles_w = OW / p.tw;
  p.tiles_h = OH / p.th;
  const int tiles_n = cdiv(B, p.tn);
  for (long ctas = (long)p.tiles_w * p.tiles_h * tiles_n * (Cout / p.bn); ctas < 100 && p.bn > 64; ctas *= 2) p.bn >>= 1;
  p.B = B; p.N = Cout; p.DH = OH; p.DW = OW; p.ds = 1; p.po = 0; p.pw = 0;
  p.dst = reinterpret_cast<__nv_bfloat16*>(y);
  p.bias = bias;
  p.ncls = 1;
  p.ntaps[0] = 9;
  const uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
  int rc;
  if (stride == 1) {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    rc = make_map(&maps.src[0], x, 4, dims, strides, box, p.kc * 2);
    if (rc) return rc;
    for (int t = 0; t < 9; ++t) p.taps[0][t] = TcTap{0, (int8_t)(t / 3 - 1), (int8_t)(t % 3 - 1), (int8_t)t};
  } else {
    // parity views: x[b, 2i+ph, 2j+pw, c]
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)(W / 2), (uint64_t)(H / 2), (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)2 * Cin * 2, (uint64_t)2 * W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const char* base = reinterpret_cast<const char*>(x) + ((size_t)ph * W + pw) * Cin * 2;
        rc = make_map(&maps.src[ph * 2 + pw], base, 4, dims, strides, box, p.kc * 2);
        if (rc) return rc;
      }
    for (int t = 0; t < 9; ++t) {
      const int r = t / 3, s = t % 3;
      // input row 2*oh + r - 1: r=0 -> odd row, index oh-1; r=1 -> even row, index oh; r=2 -> odd row, index oh
      const int ph = (r == 1) ? 0 : 1, di = (r == 0) ? -1 : 0;
      const int pw = (s == 1) ? 0 : 1, dj = (s == 0) ? -1 : 0;
      p.taps[0][t] = TcTap{(int8_t)(ph * 2 + pw), (int8_t)di, (int8_t)dj, (int8_t)t};
    }
  }
  rc = make_weight_map(&maps.w, wp, Cin, Cout, p.kc, p.bn);
  if (rc) return rc;
  return launch_tc(maps, p, p.tiles_w * p.tiles_h * tiles_n, st);
}

int conv_tc_dgrad(const void* dy, const void* wpt, void* dx, int B, int H, int W, int Cin, int Cout, int stride,
                  cudaStream_t st) {
  if (stride == 1 && conv_halo_supported(H, W, Cout, Cin)) return conv_halo_launch(dy, wpt, nullptr, dx, B, H, W, Cout, Cin, 1, st);
  if (stride == 2 && conv_s2_dgrad_halo_supported(H, W, Cin, Cout)) return conv_s2_dgrad_halo(dy, wpt, dx, B, H, W, Cin, Cout, st);
  TcMaps maps;
  TcParams p;
  memset(&p, 0, sizeof(p));
  const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;  // dy spatial size
  p.kc = pick_kc(Cout);
  p.bn = pick_bn(Cin);
  p.kchunks = Cout / p.kc;
  const int GH = H / stride, GW = W / stride;  // iterated output grid (per parity class when stride 2)
  BVAE_REQUIRE(tile_geom(GW, GH, &p.tw, &p.th, &p.tn), BVAE_ERR_UNSUPPORTED, "conv_tc_dgrad: tile geometry");
  p.tiles_w = GW / p.tw;
  p.tiles_h = GH / p.th;
  const int tiles_n = cdiv(B, p.tn);
  for (long ctas = (long)p.tiles_w * p.tiles_h * tiles_n * (Cin / p.bn) * (stride == 2 ? 4 : 1); ctas < 100 && p.bn > 64;
       ctas *= 2)
    p.bn >>= 1;
  p.B = B; p.N = Cin; p.DH = H; p.DW = W;
  p.dst = reinterpret_cast<__nv_bfloat16*>(dx);
  p.bias = nullptr;
  const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)OW, (uint64_t)OH, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)OW * Cout * 2, (uint64_t)OH * OW * Cout * 2};
  const uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
  int rc = make_map(&maps.src[0], dy, 4, dims, strides, box, p.kc * 2);
  if (rc) return rc;
  rc = make_weight_map(&maps.w, wpt, Cout, Cin, p.kc, p.bn);
  if (rc) return rc;
  if (stride == 1) {
    p.ds = 1; p.po = 0; p.pw = 0;
    p.ncls = 1;
    p.ntaps[0] = 9;
    for (int t = 0; t < 9; ++t) p.taps[0][t] = TcTap{0, (int8_t)(1 - t / 3), (int8_t)(1 - t % 3), (int8_t)t};
    return launch_tc(maps, p, p.tiles_w * p.tiles_h * tiles_n, st);
  }
  // stride 2: the four output parity classes are dense sub-convolutions; one launch, class = blockIdx.z
  p.ds = 2; p.po = 0; p.pw = 0;
  p.ncls = 4;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const int cls = ph * 2 + pw;
      int n = 0;
      for (int r = 0; r < 3; ++r) {
        if ((ph + 1 - r) % 2 != 0) continue;
        for (int s = 0; s < 3; ++s) {
          if ((pw + 1 - s) % 2 != 0) continue;
          p.taps[cls][n++] = TcTap{0, (int8_t)((ph + 1 - r) / 2), (int8_t)((pw + 1 - s) / 2), (int8_t)(r * 3 + s)};
        }
      }
      p.ntaps[cls] = n;
    }
  return launch_tc(maps, p, p.tiles_w * p.tiles_h * tiles_n, st);
}

}  // namespace bvae
