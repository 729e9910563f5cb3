// Development probe: tcgen05 B operand in MN-major layout with 32-byte rows (N = 16, SWIZZLE_32B) -- the layout of a
// [pixels][16 taps] bf16 tile delivered by TMA -- multiplied with an MN-major SW128 A tile (pixels x 64 channels):
//   D[c][t] = sum_p h[p][c] * d[p][t]      (the final conv's weight gradient)
// argv[1]: SBO bytes for the B descriptor (default 256 = 8 rows x 32 B); argv[2]: LBO bytes (default 16)
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

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap hmap, const __grid_constant__ CUtensorMap dmap, float* out, uint32_t sbo, uint32_t lbo) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* hs = smem;            // 128 px x 128 B
  uint8_t* ds = smem + 16384;    // 128 px x 32 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 4096);
  uint64_t* done = bar + 1;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tslot, 32);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tslot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 16384 + 4096);
    tma_load_4d(&hmap, bar, hs, 0, 0, 0, 0);
    tma_load_4d(&dmap, bar, ds, 0, 0, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < 8; ++k) {
      const uint64_t ad = mnmajor_desc(smem_u32(hs) + k * 2048u, 128u, 1024u);      // second 64-row half: junk (next pixel)
      uint64_t bd = (uint64_t)(((smem_u32(ds) + k * 512u) & 0x3FFFFu) >> 4);
      bd |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
      bd |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
      bd |= 1ull << 46;
      bd |= 6ull << 61;  // SWIZZLE_32B
      umma_bf16(tmem, ad, bd, idesc, k ? 1u : 0u);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 16 + i] = __uint_as_float(v[i]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

int main(int argc, char** argv) {
  const uint32_t sbo = argc > 1 ? atoi(argv[1]) : 256, lbo = argc > 2 ? atoi(argv[2]) : 16;
  const int P = 128, C = 64, T = 16;
  std::vector<float> h(P * C), d(P * T);
  srand(3);
  for (auto& v : h) v = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : d) v = bf((rand() % 2001 - 1000) / 1000.f);
  std::vector<__nv_bfloat16> hb(h.size()), db(d.size());
  for (size_t i = 0; i < h.size(); ++i) hb[i] = __float2bfloat16_rn(h[i]);
  for (size_t i = 0; i < d.size(); ++i) db[i] = __float2bfloat16_rn(d[i]);
  __nv_bfloat16 *dh, *dd; float* dout;
  cudaMalloc(&dh, hb.size() * 2); cudaMalloc(&dd, db.size() * 2); cudaMalloc(&dout, 128 * 16 * 4);
  cudaMemcpy(dh, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dd, db.data(), db.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap hmap, dmap;
  {
    const uint64_t dims[4] = {(uint64_t)C, 8, 16, 1};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)8 * C * 2, (uint64_t)P * C * 2};
    const uint32_t box[4] = {64, 8, 16, 1};
    if (make_map(&hmap, dh, 4, dims, strides, box, 128)) return 1;
  }
  {
    const uint64_t dims[4] = {(uint64_t)T, 8, 16, 1};
    const uint64_t strides[3] = {(uint64_t)T * 2, (uint64_t)8 * T * 2, (uint64_t)P * T * 2};
    const uint32_t box[4] = {16, 8, 16, 1};
    if (make_map(&dmap, dd, 4, dims, strides, box, 32)) return 1;
  }
  const size_t smem = 1024 + 16384 + 4096 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(hmap, dmap, dout, sbo, lbo);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<float> out(128 * 16);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int c = 0; c < C; ++c) for (int t = 0; t < T; ++t) {
    double acc = 0;
    for (int p = 0; p < P; ++p) acc += (double)h[p * C + c] * d[p * T + t];
    maxerr = fmax(maxerr, fabs(out[c * 16 + t] - acc)); maxref = fmax(maxref, fabs(acc));
  }
  printf("sbo %u lbo %u: max|err| = %.4g (max|ref| = %.4g) %s\n", sbo, lbo, maxerr, maxref, maxerr <= 2e-2 * maxref ? "OK" : "MISMATCH");
  return maxerr <= 2e-2 * maxref ? 0 : 1;
}
