// TEST INFRASTRUCTURE ONLY.  shapy_b200/csrc/metrics.cuh (the per-point functions of the P2P metric kernels) compiled for the
// host from the same source and looped over a batch, for the CPU tests.  Built by oracle/build_oracle.py.
#define SHAPY_HD inline
#include "../shapy_b200/csrc/metrics.cuh"
#include <vector>

extern "C" void p2p_host(const int *rp_in, const int *col_in, const float *val_in, const int *rp_tg, const int *col_tg,
                         const float *val_tg, const float *v_in, const float *v_tg, int B, int P, int V1, int V2, int align,
                         float *error, float *mean) {
  std::vector<float> d((size_t)P * 3);
  for (int b = 0; b < B; ++b) {
    double s[3] = {0, 0, 0};
    for (int p = 0; p < P; ++p) {
      float a[3], c[3];
      shapy::csr_row_point(rp_in, col_in, val_in, v_in + (size_t)b * V1 * 3, p, a);
      shapy::csr_row_point(rp_tg, col_tg, val_tg, v_tg + (size_t)b * V2 * 3, p, c);
      for (int k = 0; k < 3; ++k) { d[(size_t)p * 3 + k] = a[k] - c[k]; s[k] += d[(size_t)p * 3 + k]; }
    }
    float t[3] = {0, 0, 0};
    if (align) for (int k = 0; k < 3; ++k) t[k] = (float)(-s[k] / P);
    double m = 0;
    for (int p = 0; p < P; ++p) {
      const float e = shapy::aligned_error(&d[(size_t)p * 3], t);
      error[(size_t)b * P + p] = e;
      m += e;
    }
    mean[b] = (float)(m / P);
  }
}
