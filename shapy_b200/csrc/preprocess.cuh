// Input stage arithmetic shared by the CUDA kernel (preprocess.cu) and the host-compiled copy the tests build from the
// same source (oracle/preprocess_host.cpp): crop window + cv2-style bilinear resize + normalise, one output pixel.
//
// Reference (regressor/human_shape): utils/img_utils.py:57-61 (uint8 -> float32 / 255, clip), utils/transf_utils.py:51-96
// (crop: window [ul, br), zero padding, cv2.resize INTER_LINEAR), data/transforms/transforms.py:710-733 (clamp, (x - mean) / std).
// cv2.resize INTER_LINEAR for float images (imgproc/resize.cpp, resizeGeneric_): source coordinate
// f = (float)((d + 0.5) * scale - 0.5) with scale = 1 / ((double)dst / src); s = floor(f); f -= s; s < 0 -> (0, 0);
// s >= src - 1 -> (src - 1, 0); value = (S[y0][x0] (1 - fx) + S[y0][x1] fx) (1 - fy) + (S[y1][x0] (1 - fx) + S[y1][x1] fx) fy.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef SHAPY_HD
#define SHAPY_HD __host__ __device__ __forceinline__
#endif

namespace shapy {

// source index and weight of the second tap along one axis
SHAPY_HD void resize_coord(int d, int dst, int src, int &s, float &w) {
  const double inv_scale = (double)dst / (double)src;
  const double scale = 1.0 / inv_scale;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  // floor for values that can be slightly negative
  int i = (int)f;
  if ((float)i > f) --i;
  f -= (float)i;
  if (i < 0) { i = 0; f = 0.f; }
  if (i >= src - 1) { i = src - 1; f = 0.f; }
  s = i;
  w = f;
}

// one channel of the zero-padded window, window coordinates (wy, wx)
SHAPY_HD float window_fetch(const uint8_t *img, int H, int W, int ul_x, int ul_y, int wy, int wx, int c) {
  const int oy = ul_y + wy, ox = ul_x + wx;
  if (oy < 0 || oy >= H || ox < 0 || ox >= W) return 0.f;
  return (float)img[((size_t)oy * W + ox) * 3 + c] / 255.0f;
}

// out[c] = normalised value of output pixel (y, x) of the size x size crop
SHAPY_HD void preprocess_pixel(const uint8_t *img, int H, int W, int ul_x, int ul_y, int br_x, int br_y, int size, int y, int x,
                               const float *mean, const float *stdv, float *out) {
  const int src_w = br_x - ul_x, src_h = br_y - ul_y;
  int sx, sy;
  float fx, fy;
  resize_coord(x, size, src_w, sx, fx);
  resize_coord(y, size, src_h, sy, fy);
  const int sx1 = sx + 1 < src_w ? sx + 1 : src_w - 1, sy1 = sy + 1 < src_h ? sy + 1 : src_h - 1;
  const float a0 = 1.f - fx, b0 = 1.f - fy;
  for (int c = 0; c < 3; ++c) {
    const float s00 = window_fetch(img, H, W, ul_x, ul_y, sy, sx, c), s01 = window_fetch(img, H, W, ul_x, ul_y, sy, sx1, c);
    const float s10 = window_fetch(img, H, W, ul_x, ul_y, sy1, sx, c), s11 = window_fetch(img, H, W, ul_x, ul_y, sy1, sx1, c);
    const float r0 = s00 * a0 + s01 * fx, r1 = s10 * a0 + s11 * fx;
    float v = r0 * b0 + r1 * fy;
    v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
    out[c] = (v - mean[c]) / stdv[c];
  }
}

}  // namespace shapy
