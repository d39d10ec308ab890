// Iterative regression head: p_0 = mean; p_{k+1} = p_k + MLP(cat[f, p_k]), MLP = 3 x Linear, no
// activation (reference regressor/human_shape/models/common/networks.py:536-592, MLP 387-396).
// The feature part of the first layer, f . W0[:, :F]^T + b0, is identical in every stage and is
// computed once; per stage only the 145-wide parameter part is added (exact re-association).
// fp32 SIMT skinny GEMMs (M = batch <= a few hundred): 20.6 MFLOP / image, launch-latency bound.
#include "common.cuh"

namespace shapy {

constexpr int HB = 32;  // tile M, N, K

// Y[m][n] = sum_k X[m][k] W[n][k] + (bias ? bias[n] : 0) + (add ? add[m*ldadd + n] : 0)
__global__ void __launch_bounds__(256) skinny_gemm_nt(const float *__restrict__ X, int ldx, const float *__restrict__ W,
                                                      int ldw, const float *__restrict__ bias,
                                                      const float *__restrict__ add, int ldadd, float *__restrict__ Y,
                                                      int ldy, int M, int N, int K) {
  __shared__ float Xs[HB][HB + 1];
  __shared__ float Ws[HB][HB + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * HB, n0 = blockIdx.x * HB;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < K; k0 += HB) {
    for (int i = threadIdx.x; i < HB * HB; i += 256) {
      int r = i / HB, c = i % HB;  // c along K (contiguous)
      Xs[c][r] = (m0 + r < M && k0 + c < K) ? X[(size_t)(m0 + r) * ldx + k0 + c] : 0.f;
      Ws[c][r] = (n0 + r < N && k0 + c < K) ? W[(size_t)(n0 + r) * ldw + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < HB; ++k) {
      float x0 = Xs[k][ty * 2], x1 = Xs[k][ty * 2 + 1], w0 = Ws[k][tx * 2], w1 = Ws[k][tx * 2 + 1];
      acc[0][0] += x0 * w0; acc[0][1] += x0 * w1; acc[1][0] += x1 * w0; acc[1][1] += x1 * w1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int m = m0 + ty * 2 + i, n = n0 + tx * 2 + j;
      if (m < M && n < N) {
        float v = acc[i][j];
        if (bias) v += bias[n];
        if (add) v += add[(size_t)m * ldadd + n];
        Y[(size_t)m * ldy + n] = v;
      }
    }
}

static int gemm(cudaStream_t st, const float *X, int ldx, const float *W, int ldw, const float *bias, const float *add,
                int ldadd, float *Y, int ldy, int M, int N, int K) {
  dim3 grid(ceil_div(N, HB), ceil_div(M, HB));
  skinny_gemm_nt<<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, add, ldadd, Y, ldy, M, N, K);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}

}  // namespace shapy

using namespace shapy;

extern "C" size_t shapy_head_workspace_bytes(int B, int F, int P, int h0, int h1) {
  (void)F; (void)P;
  return align_up((size_t)B * h0 * 4, 256) * 2 + align_up((size_t)B * h1 * 4, 256);
}

extern "C" int shapy_head_forward(const float *feats, int B, int F, int P, int h0, int h1, const float *W0,
                                  const float *b0, const float *W1, const float *b1, const float *W2, const float *b2,
                                  const float *mean, int num_stages, float *params_out, void *workspace,
                                  size_t workspace_bytes, void *stream) {
  SHAPY_REQUIRE(feats && W0 && b0 && W1 && b1 && W2 && b2 && mean && params_out, "shapy_head_forward: null argument");
  SHAPY_REQUIRE(B > 0 && num_stages >= 1, "shapy_head_forward: bad sizes");
  SHAPY_REQUIRE(workspace && workspace_bytes >= shapy_head_workspace_bytes(B, F, P, h0, h1), "head workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char *w = (char *)workspace;
  float *base0 = (float *)w; w += align_up((size_t)B * h0 * 4, 256);
  float *a0 = (float *)w; w += align_up((size_t)B * h0 * 4, 256);
  float *a1 = (float *)w;
  int rc;
  if ((rc = gemm(st, feats, F, W0, F + P, b0, nullptr, 0, base0, h0, B, h0, F))) return rc;
  for (int s = 0; s < num_stages; ++s) {
    const float *p = s == 0 ? mean : params_out + (size_t)(s - 1) * B * P;
    const int ldp = s == 0 ? 0 : P;  // stage 0: the mean row is broadcast
    float *pn = params_out + (size_t)s * B * P;
    if ((rc = gemm(st, p, ldp, W0 + F, F + P, nullptr, base0, h0, a0, h0, B, h0, P))) return rc;
    if ((rc = gemm(st, a0, h0, W1, h0, b1, nullptr, 0, a1, h1, B, h1, h0))) return rc;
    if ((rc = gemm(st, a1, h1, W2, h1, b2, p, ldp, pn, P, B, P, h1))) return rc;
  }
  return SHAPY_OK;
}
