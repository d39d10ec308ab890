// TEST INFRASTRUCTURE ONLY.  shapy_b200/csrc/attributes.cuh (the per-output function of the B2A kernel) compiled for the
// host from the same source and looped over a batch, for the CPU tests.  Built by oracle/build_oracle.py.
#define SHAPY_HD inline
#include "../shapy_b200/csrc/attributes.cuh"

extern "C" void b2a_host(const float *betas, const int *gender, const float *Wm, const float *bm, const float *Wf,
                         const float *bf, int B, int n, int n_out, float *out) {
  for (int b = 0; b < B; ++b)
    for (int o = 0; o < n_out; ++o) {
      float v = 0.f;
      if (gender[b] == 0) v = shapy::b2a_output(betas + (size_t)b * n, n, Wm, bm, o);
      else if (gender[b] == 1) v = shapy::b2a_output(betas + (size_t)b * n, n, Wf, bf, o);
      out[(size_t)b * n_out + o] = v;
    }
}
