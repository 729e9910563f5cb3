// Development probe (not part of the library): do tcgen05 shared-memory descriptors accept
// start addresses / strides that are multiples of 128 B but not of the 1024-B swizzle atom?
// One TMA load brings an (18 x 10 pixel) x 64-channel halo of a 16x8 tile into SW128 shared
// memory; the nine 3x3 taps are then read as shifted windows of that one copy.
//   mode K : fprop-style, A = pixels x channels (K-major), B = weights [tap][n][k]
//   mode M : wgrad-style, A = channels x pixels (MN-major), B = dy tile (MN-major)
// argv[1] = base-offset policy: 0 -> field left 0, 1 -> (start >> 7) & 7
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <vector>
#include <cmath>
#include "../tc_ptx.cuh"

namespace bvae {
void set_error(const char* fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); fputc('\n', stderr); }
int cuda_fail(cudaError_t e, const char* what) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -3; }
unsigned long long g_launches = 0;
int num_sms() { return 148; }
}
using namespace bvae;

static constexpr int TH = 16, TW = 8, HW_ = TW + 2, HH = TH + 2;
static constexpr uint32_t HALO_BYTES = HH * HW_ * 128;  // 23040

__device__ __forceinline__ uint64_t with_base(uint64_t d, uint32_t saddr, int policy) {
  if (policy) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap bmap, float* out, int mode, int policy) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* halo = smem;                    // 23040 -> padded to 24576
  uint8_t* bsm = smem + 24576;             // mode K: 9 x 8192 weights; mode M: 16384 dy tile
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 24576 + 9 * 8192);
  uint64_t* done = bar + 1;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tslot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tslot;
  if (threadIdx.x == 0) {
    if (mode == 0) {
      mbar_expect_tx(bar, HALO_BYTES + 9 * 8192);
      tma_load_4d(&xmap, bar, halo, 0, -1, -1, 0);
      for (int t = 0; t < 9; ++t) tma_load_3d(&bmap, bar, bsm + t * 8192, 0, 0, t);
    } else {
      mbar_expect_tx(bar, HALO_BYTES + 16384);
      tma_load_4d(&xmap, bar, halo, 0, -1, -1, 0);
      tma_load_4d(&bmap, bar, bsm, 0, 0, 0, 0);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (mode == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int first = 1;
      for (int t = 0; t < 9; ++t) {
        const int di = t / 3 - 1, dj = t % 3 - 1;
        const uint32_t a_addr = smem_u32(halo) + (uint32_t)(((di + 1) * HW_ + (dj + 1)) * 128);
        const uint32_t b_addr = smem_u32(bsm) + t * 8192;
        for (int k = 0; k < 4; ++k) {
          // K-major SW128: 8-pixel groups (one image row of the tile) are HW_*128 = 1280 B apart
          uint64_t ad = (uint64_t)(((a_addr + k * 32) & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((HW_ * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
          ad = with_base(ad, a_addr, policy);
          const uint64_t bd = kmajor_desc(b_addr, 128) + (uint64_t)(k * 2);
          umma_bf16(tmem, ad, bd, idesc, first ? 0u : 1u);
          first = 0;
        }
      }
    } else {
      // D[128 = two taps x 64 ci][64 co]; slot s covers taps 2s, 2s+1 (tap 9 = padding, skipped by host check)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int s = 0; s < 5; ++s) {
        const int t0 = 2 * s, t1 = (2 * s + 1 < 9) ? 2 * s + 1 : 8;
        const uint32_t a0 = smem_u32(halo) + (uint32_t)(((t0 / 3) * HW_ + (t0 % 3)) * 128);
        const uint32_t a1 = smem_u32(halo) + (uint32_t)(((t1 / 3) * HW_ + (t1 % 3)) * 128);
        const uint32_t lbo = (a1 > a0) ? (a1 - a0) : 128;   // s = 4: second half is junk
        for (int k = 0; k < 8; ++k) {  // 16 pixels = 2 image rows of the tile per K step
          const uint32_t as = a0 + (uint32_t)(k * 2 * HW_ * 128);
          uint64_t ad = mnmajor_desc(as, lbo, HW_ * 128);
          ad = with_base(ad, as, policy);
          const uint64_t bd = mnmajor_desc(smem_u32(bsm) + k * 2048u, 8192, 1024u);
          umma_bf16(tmem + s * 64, ad, bd, idesc, k ? 1u : 0u);
        }
      }
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  const int nslots = mode == 0 ? 1 : 5;
  for (int s = 0; s < nslots; ++s)
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + s * 64 + c, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[((size_t)s * 128 + warp * 32 + lane) * 64 + c + i] = __uint_as_float(v[i]);
    }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

int main(int argc, char** argv) {
  const int policy = argc > 1 ? atoi(argv[1]) : 1;
  const int H = TH, W = TW, C = 64, N = 64;
  std::vector<float> x(H * W * C), w(9 * N * C), dy(H * W * N);
  srand(1);
  for (auto& v : x) v = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : w) v = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : dy) v = bf((rand() % 2001 - 1000) / 1000.f);
  std::vector<__nv_bfloat16> xb(x.size()), wb(w.size()), dyb(dy.size());
  for (size_t i = 0; i < x.size(); ++i) xb[i] = __float2bfloat16_rn(x[i]);
  for (size_t i = 0; i < w.size(); ++i) wb[i] = __float2bfloat16_rn(w[i]);
  for (size_t i = 0; i < dy.size(); ++i) dyb[i] = __float2bfloat16_rn(dy[i]);
  __nv_bfloat16 *dx, *dw, *ddy; float* dout;
  cudaMalloc(&dx, xb.size() * 2); cudaMalloc(&dw, wb.size() * 2); cudaMalloc(&ddy, dyb.size() * 2); cudaMalloc(&dout, 5 * 128 * 64 * 4);
  cudaMemcpy(dx, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(ddy, dyb.data(), dyb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap xmap, wmap, dymap;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, 1};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {64, (uint32_t)HW_, (uint32_t)HH, 1};
    if (make_map(&xmap, dx, 4, dims, strides, box, 128)) return 1;
    const uint32_t box2[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
    const uint64_t dims2[4] = {(uint64_t)N, (uint64_t)W, (uint64_t)H, 1};
    const uint64_t strides2[3] = {(uint64_t)N * 2, (uint64_t)W * N * 2, (uint64_t)H * W * N * 2};
    if (make_map(&dymap, ddy, 4, dims2, strides2, box2, 128)) return 1;
  }
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)N, 9};
    const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)C * N * 2};
    const uint32_t box[3] = {64, 64, 1};
    if (make_map(&wmap, dw, 3, dims, strides, box, 128)) return 1;
  }
  const size_t smem = 1024 + 24576 + 9 * 8192 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> out(5 * 128 * 64);
  int rc = 0;
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dout, 0, out.size() * 4);
    probe_kernel<<<1, 128, smem>>>(xmap, mode == 0 ? wmap : dymap, dout, mode, policy);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    if (mode == 0) {
      for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int t = 0; t < 9; ++t) {
          const int ii = i + t / 3 - 1, jj = j + t % 3 - 1;
          if (ii < 0 || ii >= H || jj < 0 || jj >= W) continue;
          for (int c = 0; c < C; ++c) acc += (double)x[(ii * W + jj) * C + c] * w[(t * N + n) * C + c];
        }
        const double got = out[(i * W + j) * 64 + n];
        maxerr = fmax(maxerr, fabs(got - acc)); maxref = fmax(maxref, fabs(acc));
      }
    } else {
      for (int t = 0; t < 9; ++t) for (int c = 0; c < C; ++c) for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) {
          const int ii = i + t / 3 - 1, jj = j + t % 3 - 1;
          if (ii < 0 || ii >= H || jj < 0 || jj >= W) continue;
          acc += (double)x[(ii * W + jj) * C + c] * dy[(i * W + j) * N + n];
        }
        const int s = t / 2, half = t % 2;
        const double got = out[((size_t)s * 128 + half * 64 + c) * 64 + n];
        maxerr = fmax(maxerr, fabs(got - acc)); maxref = fmax(maxref, fabs(acc));
      }
    }
    printf("policy %d mode %c: max|err| = %.4g  (max|ref| = %.4g)  %s\n", policy, mode == 0 ? 'K' : 'M', maxerr, maxref,
           maxerr <= 2e-2 * maxref ? "OK" : "MISMATCH");
    if (maxerr > 2e-2 * maxref) rc = 1;
  }
  return rc;
}
