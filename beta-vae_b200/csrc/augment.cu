// Train-time augmentation of a uint8 image batch on the device (input pipeline).
// (reference: src/data_processing/augmentations.py:5-18 -- torchvision RandomHorizontalFlip -> RandomRotation(d) ->
//  ColorJitter(brightness) applied per sample to PIL "L" / "RGB" images inside the DataLoader workers, then ToTensor.)
//
// Byte work, bit-exact by construction: the host (bvae_rt/augment.py) draws the per-sample parameters and turns the
// rotation into the six 16.16 fixed-point coefficients Pillow's nearest-neighbour affine walk uses; the kernel does
// what Pillow's C loops do per destination pixel,
//     source column = (a2 + y*a1 + x*a0) >> 16,   source row = (a5 + y*a4 + x*a3) >> 16,   outside -> 0,
// reads the (optionally mirrored) source byte, applies Pillow's blend-with-black brightness in fp32
//     f in [0, 1]: (uint8)(f * v);   otherwise clip(f * v, 0, 255) truncated;   f == 1: v,
// and writes the byte and/or ToTensor's float (v / 255, IEEE division).  The exact transposes Pillow substitutes for
// 0 / 90 / 180 / 270 degrees are affine walks with integer coefficients, also prepared by the host.
// One thread = four consecutive destination pixels of one row (one 4-byte store, one 16-byte float store); the batch
// is 0.5-6 MB, so the kernel is a few microseconds on the copy stream, under the previous step.
#include "common.cuh"

namespace bvae {

static constexpr int AUG_FLIP = 1 << 4, AUG_BRIGHT = 1 << 5;

struct AugSample {        // 32 bytes, int32[8] on the wire
  int flags;              // AUG_FLIP | AUG_BRIGHT
  int a0, a1, a2, a3, a4, a5;
  float factor;
};

__device__ __forceinline__ uint32_t aug_pixel(const uint8_t* __restrict__ img, const AugSample& s, int x, int y, int H, int W) {
  const long long xx = (long long)s.a2 + (long long)y * s.a1 + (long long)x * s.a0;
  const long long yy = (long long)s.a5 + (long long)y * s.a4 + (long long)x * s.a3;
  const long long xin = xx >> 16, yin = yy >> 16;
  if (xin < 0 || xin >= W || yin < 0 || yin >= H) return 0u;
  const int xs = (s.flags & AUG_FLIP) ? W - 1 - (int)xin : (int)xin;
  uint32_t v = img[(size_t)yin * W + xs];
  if ((s.flags & AUG_BRIGHT) && s.factor != 1.0f) {
    float t = __fmul_rn(s.factor, (float)v);
    if (!(s.factor >= 0.0f && s.factor <= 1.0f)) t = fminf(fmaxf(t, 0.0f), 255.0f);
    v = (uint32_t)t;      // truncation, as the C cast
  }
  return v;
}

__global__ void __launch_bounds__(256)
    augment_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, float* __restrict__ dst_unit,
                      const AugSample* __restrict__ params, int C, int H, int W) {
  const int b = blockIdx.z, c = blockIdx.y;
  const AugSample s = params[b];
  const size_t plane = ((size_t)b * C + c) * (size_t)H * W;
  const uint8_t* img = src + plane;
  const int wq = W / 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * wq; i += gridDim.x * blockDim.x) {
    const int y = i / wq, x0 = (i % wq) * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = aug_pixel(img, s, x0 + k, y, H, W);
    const size_t o = plane + (size_t)y * W + x0;
    if (dst) *reinterpret_cast<uint32_t*>(dst + o) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    if (dst_unit)
      *reinterpret_cast<float4*>(dst_unit + o) = make_float4(__fdiv_rn((float)v[0], 255.f), __fdiv_rn((float)v[1], 255.f),
                                                            __fdiv_rn((float)v[2], 255.f), __fdiv_rn((float)v[3], 255.f));
  }
}

}  // namespace bvae

using namespace bvae;

extern "C" int bvae_augment_u8(const void* src, void* dst_u8, float* dst_unit, const void* params, int B, int C, int H,
                               int W, bvae_stream_t stream) {
  BVAE_REQUIRE(src && params && (dst_u8 || dst_unit), BVAE_ERR_ARG, "augment_u8: null pointer");
  BVAE_REQUIRE(src != dst_u8, BVAE_ERR_ARG, "augment_u8: the rotation gathers, it cannot run in place");
  BVAE_REQUIRE(B > 0 && B <= 65535 && C > 0 && C <= 65535 && H > 0 && W > 0, BVAE_ERR_ARG, "augment_u8: empty shape");
  // Pillow leaves its fixed-point walk for images whose transformed corners exceed 32768; far beyond any input here
  BVAE_REQUIRE(W % 4 == 0 && H <= 8192 && W <= 8192, BVAE_ERR_UNSUPPORTED, "augment_u8: %d x %d (W %% 4 == 0, <= 8192)", H, W);
  BVAE_REQUIRE(((reinterpret_cast<uintptr_t>(dst_u8) & 3) | (reinterpret_cast<uintptr_t>(dst_unit) & 15) |
                (reinterpret_cast<uintptr_t>(params) & 3)) == 0,
               BVAE_ERR_ARG, "augment_u8: misaligned pointer");
  const int per = H * (W / 4);
  int gx = (per + 255) / 256;
  if (gx > 64) gx = 64;
  augment_u8_kernel<<<dim3(gx, C, B), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint8_t*>(src),
                                                                  reinterpret_cast<uint8_t*>(dst_u8), dst_unit,
                                                                  reinterpret_cast<const AugSample*>(params), C, H, W);
  BVAE_LAUNCH_CHECK("augment_u8");
  return BVAE_OK;
}
