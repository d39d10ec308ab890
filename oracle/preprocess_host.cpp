// TEST INFRASTRUCTURE ONLY.  The per-pixel function of the product kernel (shapy_b200/csrc/preprocess.cuh), compiled for
// the host from the SAME source, looped over a crop on the CPU: lets the CPU test suite check the exact code the GPU
// runs against the numpy restatement (oracle/preprocess_oracle.py) and the reference's golden crops, so that only the
// thread indexing of preprocess.cu is left to the GPU tests.  Built by oracle/build_oracle.py into oracle/_build/.
#define SHAPY_HD inline
#include "../shapy_b200/csrc/preprocess.cuh"

extern "C" void preprocess_host(const unsigned char *img, int H, int W, int ul_x, int ul_y, int br_x, int br_y, int size,
                                const float *mean, const float *stdv, float *out /* (3, size, size) */) {
  for (int y = 0; y < size; ++y)
    for (int x = 0; x < size; ++x) {
      float v[3];
      shapy::preprocess_pixel(img, H, W, ul_x, ul_y, br_x, br_y, size, y, x, mean, stdv, v);
      for (int c = 0; c < 3; ++c) out[((size_t)c * size + y) * size + x] = v[c];
    }
}
