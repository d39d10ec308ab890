// Iterative regression head: p_0 = mean; p_{k+1} = p_k + MLP(cat[f, p_k]), MLP = 3 x Linear, no
// activation (reference regressor/human_shape/models/common/networks.py:536-592, MLP 387-396).
// The feature part of the first layer, f . W0[:, :F]^T + b0, is identical in every stage and is
// computed once; per stage only the 145-wide parameter part is added (exact re-association).
// fp32 SIMT skinny GEMMs (M = batch <= a few hundred): 20.6 MFLOP / image, launch-latency bound.
#include "common.cuh"

namespace shapy {

constexpr int HM = 64, HN = 32, HK = 32;  // CTA tile: all (<= 64) rows x 32 columns, K chunks of 32

// Y[m][n] = sum_k X[m][k] W[n][k] + (bias ? bias[n] : 0) + (add ? add[m*ldadd + n] : 0)
// 256 threads, each a 4 (rows) x 2 (columns) register tile; operands staged k-major in shared memory so the
// inner loop is one LDS.128 + one LDS.64 per 8 FMAs.
__global__ void __launch_bounds__(256) skinny_gemm_nt(const float *__restrict__ X, int ldx, const float *__restrict__ W,
                                                      int ldw, const float *__restrict__ bias,
                                                      const float *__restrict__ add, int ldadd, float *__restrict__ Y,
                                                      int ldy, int M, int N, int K) {
  __shared__ __align__(16) float Xs[HK][HM + 4];
  __shared__ __align__(16) float Ws[HK][HN + 2];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;  // tx -> 2 columns, ty -> 4 rows
  const int m0 = blockIdx.y * HM, n0 = blockIdx.x * HN;
  float acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
  for (int k0 = 0; k0 < K; k0 += HK) {
    for (int i = t; i < HM * HK; i += 256) {
      int r = i / HK, c = i % HK;  // c along K (contiguous in memory)
      Xs[c][r] = (m0 + r < M && k0 + c < K) ? X[(size_t)(m0 + r) * ldx + k0 + c] : 0.f;
    }
    for (int i = t; i < HN * HK; i += 256) {
      int r = i / HK, c = i % HK;
      Ws[c][r] = (n0 + r < N && k0 + c < K) ? __ldg(W + (size_t)(n0 + r) * ldw + k0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < HK; ++k) {
      const float4 x = *reinterpret_cast<const float4 *>(&Xs[k][ty * 4]);
      const float2 w = *reinterpret_cast<const float2 *>(&Ws[k][tx * 2]);
      acc[0][0] += x.x * w.x; acc[0][1] += x.x * w.y;
      acc[1][0] += x.y * w.x; acc[1][1] += x.y * w.y;
      acc[2][0] += x.z * w.x; acc[2][1] += x.z * w.y;
      acc[3][0] += x.w * w.x; acc[3][1] += x.w * w.y;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int m = m0 + ty * 4 + i, n = n0 + tx * 2 + j;
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
  dim3 grid(ceil_div(N, HN), ceil_div(M, HM));
  skinny_gemm_nt<<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, add, ldadd, Y, ldy, M, N, K);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}

// ---- collapsed head: the 3-layer MLP has no activation (configs/b2a_expose_hrnet_demo.yaml:200-207), so one
// stage is the affine map  p' = p + Mf f + Mp p + c  with Mf = W2 W1 W0[:, :F], Mp = W2 W1 W0[:, F:],
// c = W2 (W1 b0 + b1) + b2, contracted once at load time in fp64 by the host mirror.
// g[m][n] = c[n] + sum_k f[m][k] MfT[k][n]: grid (ceil(P/32), B), 4 warps split K, lanes = columns.
__global__ void __launch_bounds__(512) head_affine_kernel(const float *__restrict__ f, const float *__restrict__ MfT,
                                                          const float *__restrict__ c, int F, int P, float *__restrict__ g) {
  __shared__ float part[16][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;   // 16 warps split K, lanes = columns
  const int n = blockIdx.x * 32 + lane, m = blockIdx.y;
  const int kq = (F + 15) / 16, k0 = warp * kq, k1 = min(F, k0 + kq);
  const float *fr = f + (size_t)m * F;
  float acc = 0.f;
  if (n < P) {
#pragma unroll 8
    for (int k = k0; k < k1; ++k) acc += __ldg(fr + k) * __ldg(MfT + (size_t)k * P + n);
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && n < P) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) s += part[w][lane];
    g[(size_t)m * P + n] = c[n] + s;
  }
}

// all stages of one body: p_{s+1} = p_s + g + Mp p_s   (MpT[j][n] = Mp[n][j], so lanes read consecutive n)
__global__ void __launch_bounds__(256) head_stages_kernel(const float *__restrict__ g, const float *__restrict__ MpT,
                                                          const float *__restrict__ mean, int P, int B, int num_stages,
                                                          float *__restrict__ out) {
  extern __shared__ float ps[];  // [2][P]
  const int m = blockIdx.x, t = threadIdx.x;
  for (int i = t; i < P; i += blockDim.x) ps[i] = mean[i];
  __syncthreads();
  for (int s = 0; s < num_stages; ++s) {
    const float *cur = ps + (s & 1) * P;
    float *nxt = ps + ((s + 1) & 1) * P;
    for (int n = t; n < P; n += blockDim.x) {
      float acc = 0.f;
#pragma unroll 4
      for (int j = 0; j < P; ++j) acc += __ldg(MpT + (size_t)j * P + n) * cur[j];
      float v = cur[n] + (g[(size_t)m * P + n] + acc);
      nxt[n] = v;
      out[((size_t)s * B + m) * P + n] = v;
    }
    __syncthreads();
  }
}

}  // namespace shapy

using namespace shapy;

extern "C" int shapy_head_forward_collapsed(const float *feats, int B, int F, int P, const float *MfT, const float *MpT,
                                            const float *c, const float *mean, int num_stages, float *params_out,
                                            void *workspace, size_t workspace_bytes, void *stream) {
  SHAPY_REQUIRE(feats && MfT && MpT && c && mean && params_out && workspace, "shapy_head_forward_collapsed: null argument");
  SHAPY_REQUIRE(B > 0 && num_stages >= 1 && workspace_bytes >= (size_t)B * P * 4, "shapy_head_forward_collapsed: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  float *g = (float *)workspace;
  head_affine_kernel<<<dim3(ceil_div(P, 32), B), 512, 0, st>>>(feats, MfT, c, F, P, g);
  SHAPY_LAUNCH_CHECK();
  head_stages_kernel<<<B, 256, 2 * P * sizeof(float), st>>>(g, MpT, mean, P, B, num_stages, params_out);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}


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
