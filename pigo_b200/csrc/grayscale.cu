// grayscale.cu -- RgbToGrayscale (core/grayscale.go:8-23) for NRGBA pixels: the stage right before the hot path
// (SURVEY.md section 8f, row N2).  Pure streaming: 4 bytes in, 1 byte out per pixel => HBM-bound (5 B/pixel);
// 16 pixels per thread (4 x 16-byte loads, one 16-byte store), float64 arithmetic in the reference's operation order
// (the library is built with -fmad=false).
#include <algorithm>

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

// ---- ImgToNRGBA for *image.YCbCr (core/image.go:60-76): section 8f row N3 ---------------------------------------------------
// The reference calls color.YCbCrToRGB (Go standard library, image/color/ycbcr.go, go1.22 per go.mod:3 -- not under the
// reference tree), restated here: JFIF conversion in 16.16 fixed point,
//   yy1 = y * 0x10101;  r = yy1 + 91881*cr1;  g = yy1 - 22554*cb1 - 46802*cr1;  b = yy1 + 116130*cb1   (cb1 = cb-128, cr1 = cr-128)
// each then clamped: a value whose bits 24..31 are all zero is shifted right by 16, anything else saturates to 0 (negative)
// or 255.  The constants are the ones the reference's own test restates (core/image_test.go:118-138, which rounds instead
// and therefore allows +-1); alpha is 0xff.  Chroma sample of pixel (x, y): image.YCbCr.COffset for the six subsample ratios,
// relative to the rectangle origin (min_x, min_y) like image.go:66-67.
// `gray` (optional) fuses RgbToGrayscale (core/grayscale.go:8-23) of the converted pixel, so the NRGBA image never has to
// make the round trip through HBM when only the detector's input is wanted.
__device__ __forceinline__ uint32_t sat16(int v) { return (uint32_t)v & 0xff000000u ? (uint32_t)(~(v >> 31)) & 0xffu : (uint32_t)(v >> 16); }

// One thread converts 4 consecutive pixels of a row (32-bit index arithmetic, one 16-byte store for the NRGBA output).
__global__ void __launch_bounds__(256) ycbcr_kernel(const uint8_t* __restrict__ yp, const uint8_t* __restrict__ cbp, const uint8_t* __restrict__ crp,
                                                    int y_stride, int c_stride, int sub, int min_x, int min_y, int width, int height,
                                                    uint32_t* __restrict__ nrgba, uint8_t* __restrict__ gray, int vec_ok) {
  const int xs = (sub == 1 || sub == 2) ? 1 : ((sub == 4 || sub == 5) ? 2 : 0);   // chroma x shift: 422/420 halve, 411/410 quarter
  const int ys = (sub == 2 || sub == 3 || sub == 5) ? 1 : 0;                       // chroma y shift: 420/440/410 halve
  const unsigned chunks = (unsigned)(width + 3) >> 2;
  const unsigned total = chunks * (unsigned)height;                                // < 2^31: width * height < 2^31 is checked by the host
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    const int dy = (int)(q / chunks), dx0 = (int)(q - (unsigned)dy * chunks) * 4;
    const int sy = min_y + dy;                                                     // image.go:66-67
    const uint8_t* yrow = yp + (size_t)dy * y_stride;                             // YOffset: (y-Rect.Min.Y)*YStride + (x-Rect.Min.X)
    const size_t crow = (size_t)((sy >> ys) - (min_y >> ys)) * c_stride;          // COffset row
    uint32_t px[4], lum = 0;
    const int n = min(4, width - dx0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      px[k] = 0;
      if (k < n) {
        const int dx = dx0 + k;
        const size_t sic = crow + (size_t)(((min_x + dx) >> xs) - (min_x >> xs));
        const int yy1 = (int)__ldg(yrow + dx) * 0x10101, cb1 = (int)__ldg(cbp + sic) - 128, cr1 = (int)__ldg(crp + sic) - 128;
        const uint32_t r = sat16(yy1 + 91881 * cr1), gg = sat16(yy1 - 22554 * cb1 - 46802 * cr1), b = sat16(yy1 + 116130 * cb1);
        px[k] = r | (gg << 8) | (b << 16) | 0xff000000u;
        if (gray) lum |= luma(px[k]) << (8 * k);
      }
    }
    const size_t o = (size_t)dy * width + dx0;
    if (nrgba) {
      if (vec_ok && n == 4) *reinterpret_cast<uint4*>(nrgba + o) = make_uint4(px[0], px[1], px[2], px[3]);
      else for (int k = 0; k < n; ++k) nrgba[o + k] = px[k];
    }
    if (gray) {
      if (vec_ok && n == 4) *reinterpret_cast<uint32_t*>(gray + o) = lum;
      else for (int k = 0; k < n; ++k) gray[o + k] = (uint8_t)(lum >> (8 * k));
    }
  }
}

void launch_ycbcr(const uint8_t* y, const uint8_t* cb, const uint8_t* cr, int y_stride, int c_stride, int subsample, int min_x, int min_y,
                  int width, int height, uint8_t* nrgba, uint8_t* gray, int grid, cudaStream_t st) {
  const size_t want = ((size_t)((width + 3) / 4) * height + 255) / 256;
  // vector stores need width % 4 == 0 (every row chunk starts at a multiple of 4 pixels) and aligned outputs
  const int vec_ok = (width % 4 == 0) && (((uintptr_t)nrgba) % 16 == 0) && (((uintptr_t)gray) % 4 == 0);
  ycbcr_kernel<<<(int)std::max<size_t>(1, std::min<size_t>(want, (size_t)grid)), 256, 0, st>>>(y, cb, cr, y_stride, c_stride, subsample, min_x, min_y, width,
                                                                                             height, reinterpret_cast<uint32_t*>(nrgba), gray, vec_ok);
}

}  // namespace pigo
