// Test harness (not part of the product): runs the per-pixel functions of the preprocessing kernels
// (starvector_b200/csrc/sv_preprocess_core.h) over the batch plan produced by sv_preproc_plan_host, with the same
// (image, flat index) -> (row, column) mapping as resize_h_kernel / resize_v_kernel, so that the kernels' integer and index
// arithmetic can be compared with Pillow on a machine without a GPU.  Built on the fly by tests/test_preprocess_emul.py.
#include <cstdint>

#include "../../starvector_b200/csrc/sv_preprocess_core.h"

extern "C" int emul_preprocess(const uint8_t* arena, const uint8_t* blob, int64_t meta_bytes, int n, int out_size, uint32_t* tmp,
                               uint8_t* out_rgb /* [n][S][S][3] */) {
  const svpre::ImageMeta* metas = reinterpret_cast<const svpre::ImageMeta*>(blob);
  const int32_t* coeffs = reinterpret_cast<const int32_t*>(blob + meta_bytes);
  const int S = out_size;
  for (int i = 0; i < n; ++i) {
    const svpre::ImageMeta im = metas[i];
    for (int idx = 0; idx < im.in_h * S; ++idx) {
      const int y = idx / S, xx = idx - y * S;
      tmp[im.tmp_off + idx] = svpre::horizontal_pixel(arena, coeffs, im, S, y, xx);
    }
    for (int idx = 0; idx < S * S; ++idx) {
      const int yy = idx / S, xx = idx - yy * S;
      int rgb[3];
      svpre::vertical_pixel(tmp, coeffs, im, S, S, yy, xx, rgb);
      for (int c = 0; c < 3; ++c) out_rgb[((int64_t)i * S * S + idx) * 3 + c] = (uint8_t)rgb[c];
    }
  }
  return (int)sizeof(svpre::ImageMeta);
}
