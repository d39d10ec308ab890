// B2A: shape coefficients -> linguistic attribute ratings, per gender.
// Replaces regressor/human_shape/models/common/iterative_regressor.py:761-776 (two index_selects, two Polynomial
// forwards = 2 x (2 gathers + prod + cat + addmm), two scatters into a zeros tensor) by one launch.
// Latency-bound: B x n_out threads, 65 features each.
#include "common.cuh"
#include "attributes.cuh"

namespace shapy {

// x_m / x_f: the input rows seen by the male / female regressor (B2A: the same betas; A2B: one feature vector per gender,
// iterative_regressor.py:808-836).  linear != 0: a plain Linear(n, n_out) (W is (n_out, n)), no quadratic monomials.
__global__ void b2a_kernel(const float *__restrict__ x_m, const float *__restrict__ x_f, const int *__restrict__ gender,
                           const float *__restrict__ Wm, const float *__restrict__ bm, const float *__restrict__ Wf,
                           const float *__restrict__ bf, int B, int n, int n_out, int linear, float *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n_out) return;
  const int b = i / n_out, o = i - b * n_out;
  const int g = gender[b];                     // 0 male, 1 female, anything else: no output (zeros)
  float v = 0.f;
  if (g == 0 || g == 1) {
    const float *x = (g == 0 ? x_m : x_f) + (size_t)b * n;
    const float *W = g == 0 ? Wm : Wf, *bias = g == 0 ? bm : bf;
    if (linear) {
      float acc = 0.f;
      for (int k = 0; k < n; ++k) acc += x[k] * W[(size_t)o * n + k];
      v = acc + bias[o];
    } else {
      v = b2a_output(x, n, W, bias, o);
    }
  }
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
  b2a_kernel<<<ceil_div(total, 128), 128, 0, st>>>(betas, betas, gender, w_male, b_male, w_female, b_female, B, num_betas,
                                                   num_outputs, 0, out);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}

extern "C" int shapy_a2b_forward(const float *feat_male, const float *feat_female, const int *gender, const float *w_male,
                                 const float *b_male, const float *w_female, const float *b_female, int B, int num_features,
                                 int num_betas, int linear, float *out, void *stream) {
  SHAPY_REQUIRE(feat_male && feat_female && gender && w_male && b_male && w_female && b_female && out,
                "shapy_a2b_forward: null argument");
  SHAPY_REQUIRE(B > 0 && num_features > 0 && num_features <= 64 && num_betas > 0, "shapy_a2b_forward: bad sizes");
  const int total = B * num_betas;
  b2a_kernel<<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(feat_male, feat_female, gender, w_male, b_male, w_female,
                                                                     b_female, B, num_features, num_betas, linear, out);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}
