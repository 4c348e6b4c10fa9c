// Per-pixel arithmetic of the image-preprocessing kernels, shared by the two CUDA kernels in
// sv_preprocess.cu.  It computes, for one image, what the reference's `ImageTrainProcessor` (reference
// starvector/data/util.py:40-66) and `SimpleStarVectorProcessor` (starvector_arch.py:39-45) obtain from Pillow:
// alpha paste on white (Pillow Paste.c paste_mask_L / MULDIV255) or alpha drop, white padding to a square, and the two
// passes of the 8-bit bicubic resample (Resample.c ImagingResampleHorizontal_8bpc / Vertical_8bpc: int32 accumulators,
// 22 fractional bits, rounding constant 1<<21, clip to [0,255], 8-bit intermediate image).
// The functions are `__host__ __device__` only so that tests/test_preprocess_emul.py can run the very same index
// arithmetic under g++ against Pillow on a machine without a GPU; the library itself never executes them on the host.
#ifndef SV_PREPROCESS_CORE_H
#define SV_PREPROCESS_CORE_H

#include <stdint.h>

#if defined(__CUDACC__)
#define SV_HD __host__ __device__ __forceinline__
#else
#define SV_HD inline
#endif

namespace svpre {

constexpr int kPrecisionBits = 32 - 8 - 2;        // Resample.c PRECISION_BITS for 8-bit channels

// One image of a batch.  All offsets index arenas shared by the batch.
struct ImageMeta {
  int64_t src_off;      // byte offset of pixel (0,0) in the input arena
  int64_t tmp_off;      // pixel (uint32) offset of this image's horizontal-pass output [in_h][out_w]
  int32_t width, height, channels, row_stride;   // the stored image (uint8 HWC, 3 or 4 channels)
  int32_t in_w, in_h;   // the image the resample sees: (S,S) with S=max(w,h) when padded, else (w,h)
  int32_t pad_left, pad_top;
  int32_t alpha_white;  // 1: RGBA is pasted on white with alpha as mask; 0: alpha is dropped
  int32_t kx_off, ky_off, ksize_x, ksize_y;      // int32 offsets into the coefficient arena: bounds[out][2] then taps[ksize][out]
                                                 // (tap-major: the lanes of a warp, adjacent outputs, read adjacent words)
};

SV_HD int muldiv255(int a, int b) {
  int tmp = a * b + 128;
  return ((tmp >> 8) + tmp) >> 8;
}

// Pixel (y,x) of the padded RGB image as r | g<<8 | b<<16.
SV_HD uint32_t fetch_rgb(const uint8_t* __restrict__ arena, const ImageMeta& im, int y, int x) {
  const int sy = y - im.pad_top, sx = x - im.pad_left;
  if (sy < 0 || sy >= im.height || sx < 0 || sx >= im.width) return 0x00FFFFFFu;
  const uint8_t* p = arena + im.src_off + (int64_t)sy * im.row_stride + (int64_t)sx * im.channels;
  if (im.channels == 4) {
    // one 32-bit load per RGBA pixel: the arena base is cudaMalloc-aligned, src_off a multiple of 16 and rows are tight
    // (row_stride = 4*width), so every pixel is 4-byte aligned; bytes are little-endian r,g,b,a
    const uint32_t px = *reinterpret_cast<const uint32_t*>(p);
    if (!im.alpha_white) return px & 0x00FFFFFFu;
    const int m = (int)(px >> 24);
    const int keep = muldiv255(255, 255 - m);
    const int r = keep + muldiv255((int)(px & 255u), m);
    const int g = keep + muldiv255((int)((px >> 8) & 255u), m);
    const int b = keep + muldiv255((int)((px >> 16) & 255u), m);
    return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
  }
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
}

SV_HD int clip8(int acc) {
  const int v = acc >> kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Horizontal pass: pixel (y, xx) of the [in_h][out_w] intermediate image.
SV_HD uint32_t horizontal_pixel(const uint8_t* __restrict__ arena, const int32_t* __restrict__ coeffs, const ImageMeta& im,
                                int out_w, int y, int xx) {
  const int32_t* bounds = coeffs + im.kx_off;
  const int32_t* taps = bounds + 2 * out_w + xx;
  const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
  int r = 1 << (kPrecisionBits - 1), g = r, b = r;
  for (int t = 0; t < n; ++t) {
    const uint32_t px = fetch_rgb(arena, im, y, x0 + t);
    const int k = taps[(int64_t)t * out_w];
    r += (int)(px & 255u) * k;
    g += (int)((px >> 8) & 255u) * k;
    b += (int)((px >> 16) & 255u) * k;
  }
  return (uint32_t)clip8(r) | ((uint32_t)clip8(g) << 8) | ((uint32_t)clip8(b) << 16);
}

// Vertical pass: resized bytes (r,g,b) of output pixel (yy, xx).
SV_HD void vertical_pixel(const uint32_t* __restrict__ tmp, const int32_t* __restrict__ coeffs, const ImageMeta& im, int out_w,
                          int out_h, int yy, int xx, int rgb[3]) {
  const int32_t* bounds = coeffs + im.ky_off;
  const int32_t* taps = bounds + 2 * out_h + yy;
  const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
  const uint32_t* col = tmp + im.tmp_off + (int64_t)y0 * out_w + xx;
  int r = 1 << (kPrecisionBits - 1), g = r, b = r;
  for (int t = 0; t < n; ++t) {
    const uint32_t px = col[(int64_t)t * out_w];
    const int k = taps[(int64_t)t * out_h];
    r += (int)(px & 255u) * k;
    g += (int)((px >> 8) & 255u) * k;
    b += (int)((px >> 16) & 255u) * k;
  }
  rgb[0] = clip8(r);
  rgb[1] = clip8(g);
  rgb[2] = clip8(b);
}

}  // namespace svpre
#endif
