// Evaluation metric on the GPU: point-to-point error between two meshes of different topology
// through sparse point regressors, translation aligned.  Replaces regressor/human_shape/utils/metrics.py:368-456, which
// the reference runs in float64 on the CPU after a device->host copy of every predicted mesh (evaluation.py:227-265).
// HBM / latency bound: (V1 + V2) * 12 B read per body, P * 4 B written.  All reductions are deterministic (fixed order).
#include "common.cuh"
#include "metrics.cuh"

namespace shapy {

// d[b][p] = input point - target point
__global__ void __launch_bounds__(256) p2p_diff_kernel(const int *__restrict__ rp_in, const int *__restrict__ col_in,
                                                       const float *__restrict__ val_in, const int *__restrict__ rp_tg,
                                                       const int *__restrict__ col_tg, const float *__restrict__ val_tg,
                                                       const float *__restrict__ v_in, const float *__restrict__ v_tg, int B, int P,
                                                       int V1, int V2, float *__restrict__ d) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (p >= P || b >= B) return;
  float a[3], c[3];
  csr_row_point(rp_in, col_in, val_in, v_in + (size_t)b * V1 * 3, p, a);
  csr_row_point(rp_tg, col_tg, val_tg, v_tg + (size_t)b * V2 * 3, p, c);
  float *o = d + ((size_t)b * P + p) * 3;
  o[0] = a[0] - c[0]; o[1] = a[1] - c[1]; o[2] = a[2] - c[2];
}

// same-topology variant (PointError with TranslationAlignment / NoAlignment, metrics.py:232-277, 335-366): d = est - gt
__global__ void __launch_bounds__(256) v2v_diff_kernel(const float *__restrict__ v_in, const float *__restrict__ v_tg, size_t n,
                                                       float *__restrict__ d) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = v_in[i] - v_tg[i];
}

// one block per body: t = -mean_p d  (= mean target - mean input), or 0 without alignment
__global__ void __launch_bounds__(256) p2p_translation_kernel(const float *__restrict__ d, int P, int align, float *__restrict__ t) {
  __shared__ float s[3][256];
  const int b = blockIdx.x;
  float x = 0.f, y = 0.f, z = 0.f;
  if (align)
    for (int p = threadIdx.x; p < P; p += 256) {
      const float *q = d + ((size_t)b * P + p) * 3;
      x += q[0]; y += q[1]; z += q[2];
    }
  s[0][threadIdx.x] = x; s[1][threadIdx.x] = y; s[2][threadIdx.x] = z;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w)
      for (int c = 0; c < 3; ++c) s[c][threadIdx.x] += s[c][threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x < 3) t[b * 3 + threadIdx.x] = -s[threadIdx.x][0] / (float)P;
}

// one block per body: error[b][p] = |d + t|, mean[b]
__global__ void __launch_bounds__(256) p2p_error_kernel(const float *__restrict__ d, const float *__restrict__ t, int P,
                                                        float *__restrict__ error, float *__restrict__ mean) {
  __shared__ float s[256];
  const int b = blockIdx.x;
  const float tb[3] = {t[b * 3], t[b * 3 + 1], t[b * 3 + 2]};
  float acc = 0.f;
  for (int p = threadIdx.x; p < P; p += 256) {
    const float e = aligned_error(d + ((size_t)b * P + p) * 3, tb);
    error[(size_t)b * P + p] = e;
    acc += e;
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) mean[b] = s[0] / (float)P;
}

}  // namespace shapy

using namespace shapy;

extern "C" size_t shapy_p2p_workspace_bytes(int B, int P) { return (size_t)B * P * 3 * sizeof(float) + (size_t)B * 3 * sizeof(float) + 256; }

extern "C" int shapy_p2p_error(const int *in_row_ptr, const int *in_col, const float *in_val, const int *tg_row_ptr, const int *tg_col,
                               const float *tg_val, const float *input_vertices, const float *target_vertices, int B, int P, int V1,
                               int V2, int align, float *error, float *mean_error, void *workspace, size_t workspace_bytes,
                               void *stream) {
  SHAPY_REQUIRE(in_row_ptr && in_col && in_val && tg_row_ptr && tg_col && tg_val && input_vertices && target_vertices && error &&
                    mean_error && workspace,
                "shapy_p2p_error: null argument");
  SHAPY_REQUIRE(B > 0 && B <= 65535 && P > 0 && V1 > 0 && V2 > 0, "shapy_p2p_error: bad sizes");
  SHAPY_REQUIRE(workspace_bytes >= shapy_p2p_workspace_bytes(B, P), "shapy_p2p_error: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float *d = (float *)workspace;
  float *t = d + (size_t)B * P * 3;
  p2p_diff_kernel<<<dim3(ceil_div(P, 256), B), 256, 0, st>>>(in_row_ptr, in_col, in_val, tg_row_ptr, tg_col, tg_val, input_vertices,
                                                             target_vertices, B, P, V1, V2, d);
  SHAPY_LAUNCH_CHECK();
  p2p_translation_kernel<<<B, 256, 0, st>>>(d, P, align, t);
  SHAPY_LAUNCH_CHECK();
  p2p_error_kernel<<<B, 256, 0, st>>>(d, t, P, error, mean_error);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}

extern "C" int shapy_v2v_error(const float *input_vertices, const float *target_vertices, int B, int V, int align, float *error,
                               float *mean_error, void *workspace, size_t workspace_bytes, void *stream) {
  SHAPY_REQUIRE(input_vertices && target_vertices && error && mean_error && workspace, "shapy_v2v_error: null argument");
  SHAPY_REQUIRE(B > 0 && B <= 65535 && V > 0, "shapy_v2v_error: bad sizes");
  SHAPY_REQUIRE(workspace_bytes >= shapy_p2p_workspace_bytes(B, V), "shapy_v2v_error: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float *d = (float *)workspace;
  float *t = d + (size_t)B * V * 3;
  const size_t n = (size_t)B * V * 3;
  v2v_diff_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(input_vertices, target_vertices, n, d);
  SHAPY_LAUNCH_CHECK();
  p2p_translation_kernel<<<B, 256, 0, st>>>(d, V, align, t);
  SHAPY_LAUNCH_CHECK();
  p2p_error_kernel<<<B, 256, 0, st>>>(d, t, V, error, mean_error);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}
