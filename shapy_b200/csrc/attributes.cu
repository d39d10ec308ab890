// B2A: shape coefficients -> linguistic attribute ratings, per gender.
// Replaces regressor/human_shape/models/common/iterative_regressor.py:761-776 (two index_selects, two Polynomial
// forwards = 2 x (2 gathers + prod + cat + addmm), two scatters into a zeros tensor) by one launch.
// Latency-bound: B x n_out threads, 65 features each.
#include "common.cuh"
#include "attributes.cuh"

namespace shapy {

__global__ void b2a_kernel(const float *__restrict__ betas, const int *__restrict__ gender, const float *__restrict__ Wm,
                           const float *__restrict__ bm, const float *__restrict__ Wf, const float *__restrict__ bf, int B, int n,
                           int n_out, float *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n_out) return;
  const int b = i / n_out, o = i - b * n_out;
  const int g = gender[b];                     // 0 male, 1 female, anything else: no attributes (zeros)
  float v = 0.f;
  if (g == 0) v = b2a_output(betas + (size_t)b * n, n, Wm, bm, o);
  else if (g == 1) v = b2a_output(betas + (size_t)b * n, n, Wf, bf, o);
  out[i] = v;
}

}  // namespace shapy

using namespace shapy;

extern "C" int shapy_b2a_forward(const float *betas, const int *gender, const float *w_male, const float *b_male,
                                 const float *w_female, const float *b_female, int B, int num_betas, int num_outputs, float *out,
                                 void *stream) {
  SHAPY_REQUIRE(betas && gender && w_male && b_male && w_female && b_female && out, "shapy_b2a_forward: null argument");
  SHAPY_REQUIRE(B > 0 && num_betas > 0 && num_betas <= 64 && num_outputs > 0, "shapy_b2a_forward: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int total = B * num_outputs;
  b2a_kernel<<<ceil_div(total, 128), 128, 0, st>>>(betas, gender, w_male, b_male, w_female, b_female, B, num_betas, num_outputs, out);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}
