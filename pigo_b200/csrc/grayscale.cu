// grayscale.cu -- RgbToGrayscale (core/grayscale.go:8-23) for NRGBA pixels: the stage right before the hot path
// (SURVEY.md section 8f, row N2).  Pure streaming: 4 bytes in, 1 byte out per pixel => HBM-bound (5 B/pixel);
// 16 pixels per thread (4 x 16-byte loads, one 16-byte store), float64 arithmetic in the reference's operation order
// (the library is built with -fmad=false).
#include "common.cuh"
#include "host.h"

namespace pigo {

// color.NRGBA.RGBA(): c = v | v<<8; c = c*A/0xff  (image/color)
__device__ __forceinline__ uint32_t luma(uint32_t px) {
  const uint32_t a = px >> 24;
  const uint32_t r = ((px & 0xffu) * 0x101u) * a / 0xffu;
  const uint32_t g = (((px >> 8) & 0xffu) * 0x101u) * a / 0xffu;
  const uint32_t b = (((px >> 16) & 0xffu) * 0x101u) * a / 0xffu;
  const double y = __dadd_rn(__dadd_rn(__dmul_rn(0.299, (double)r), __dmul_rn(0.587, (double)g)), __dmul_rn(0.114, (double)b));
  return (uint32_t)(int)(y / 256.0) & 0xffu;   // uint8(float64) truncation; y/256 < 256
}

__global__ void __launch_bounds__(256) gray_kernel(const uint32_t* __restrict__ rgba, size_t npix, uint8_t* __restrict__ gray, int vec_ok) {
  const size_t nvec = vec_ok ? npix / 16 : 0;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(rgba) + v * 4;
    uint4 q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = __ldg(src + k);
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = luma(q[k].x) | (luma(q[k].y) << 8) | (luma(q[k].z) << 16) | (luma(q[k].w) << 24);
    reinterpret_cast<uint4*>(gray)[v] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (size_t i = nvec * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x)
    gray[i] = (uint8_t)luma(__ldg(rgba + i));
}

void launch_gray(const uint8_t* rgba, size_t npix, uint8_t* gray, int grid, cudaStream_t st) {
  const int vec_ok = (((uintptr_t)rgba) % 16 == 0) && (((uintptr_t)gray) % 16 == 0);
  gray_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(rgba), npix, gray, vec_ok);
}

}  // namespace pigo
