// Shared device/host helpers for the bvae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bvae_b200.h"

namespace bvae {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define BVAE_REQUIRE(cond, code, ...)       \
  do {                                      \
    if (!(cond)) {                          \
      ::bvae::set_error(__VA_ARGS__);       \
      return (code);                        \
    }                                       \
  } while (0)

extern unsigned long long g_launches;  // kernels launched by this library (bvae_launch_count)

#define BVAE_LAUNCH_CHECK(name)                                  \
  do {                                                           \
    ++::bvae::g_launches;                                        \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ::bvae::cuda_fail(e__, name); \
  } while (0)

static inline cudaStream_t as_stream(bvae_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t dtype_size(int dt) { return dt == BVAE_BF16 ? 2 : 4; }

int num_sms();

// ---------------------------------------------------------------------------------------------
// element types
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int kVec = 4;  // elements per 16-byte vector
  __device__ static __forceinline__ float to_f(float v) { return v; }
  __device__ static __forceinline__ float from_f(float v) { return v; }
  // two neighbours with one store; `p` must be 8-byte aligned
  __device__ static __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kVec = 8;
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  // two neighbours with one store; `p` must be 4-byte aligned
  __device__ static __forceinline__ void store2(__nv_bfloat16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
};

// 16-byte vector of T, unpacked to / packed from fp32 registers.
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  // raw 16-byte load / unpack, split so a loop can put several loads in flight before it touches the data
  __device__ static __forceinline__ uint4 load_raw(const float* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(const uint4& t) {
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { unpack(*reinterpret_cast<const uint4*>(p)); }
  __device__ static __forceinline__ uint4 load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(const uint4& t) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the whole block; result valid in thread 0 (and broadcast to all). `red` >= 32 elems.
template <typename A>
__device__ __forceinline__ A block_sum(A v, A* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  A r = (threadIdx.x < nw) ? red[threadIdx.x] : A(0);
  if (w == 0) {
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

}  // namespace bvae
