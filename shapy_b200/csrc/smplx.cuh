// Shared declarations of the SMPL-X kernels (smplx.cu, smplx_lbs.cu).
#pragma once
#include <cuda.h>

#include <map>
#include <mutex>

#include "common.cuh"

namespace shapy {

constexpr int kMaxJoints = 64;
constexpr int kMaxCoef = 32;

struct SmplxDev {
  int V, J, NB, NE, NC, F, L, D, rows, K, n_chain, n_levels, ell_w_n, n_extra, n_over;
  float *v_template, *shapedirs /* [NC][3V] */, *posedirs /* [(J-1)*9][3V] */;
  float *basis;   /* [NC + (J-1)*9][V3p]: shape, expression and pose blend-shape rows, 16-byte aligned rows */
  int V3p;
  float *J_template /* [3J] */, *J_dirs /* [3J][NC] */;
  int *ell_idx;   /* [W][V] */
  float *ell_w;   /* [W][V] */
  int *parents, *level_joints, *level_off; /* level_off[n_levels+1] */
  int *faces;     /* [F][3] */
  int *lmk_vidx;  /* [L][3] */
  float *lmk_bc;  /* [L][3] */
  int *dyn_vidx;  /* [rows][D][3] */
  float *dyn_bc;  /* [rows][D][3] */
  int *neck;      /* [n_chain] */
  int *ex_ptr, *ex_col;
  float *ex_val;
  int *over_src, *over_tgt;
  // ---- packed operands of the fused tcgen05 LBS kernel (smplx_lbs.cu)
  int Vpad;               /* V rounded up to a multiple of 128 */
  int KPpad;              /* (J-1)*9 rounded up to a multiple of 64 */
  __half *pbasis;         /* [KPpad / 32 k-blocks][3 planes][Vpad][hi 32 | lo 32] fp16: posedirs * kPoseScale as hi + lo / 2048 */
  float *shape_planes;    /* [(NB + 1) * 3][Vpad]: row l*3+c = shapedirs coefficient l, coordinate c; rows NB*3+c = template */
};

}  // namespace shapy

struct shapy_smplx {
  shapy::SmplxDev d;
  std::vector<void *> allocs;
  CUtensorMap basis_map;                    // (64, Vpad, 3, KPpad / 32) fp16, box (64, 128, 3, 1), SWIZZLE_128B
  bool fused_ok = false;                    // the packed operands above exist and the maps encoded
  // blend-GEMM operand of the fused kernel: skinning weights folded onto the rotated joints (joints >= n_rot carry
  // their nearest rotated ancestor's transform), built lazily per n_rot from the host copies below
  std::vector<float> h_lbs_weights;         // (V, J) dense, as given
  std::vector<int> h_parents;
  struct WTiles { __half *dev = nullptr; CUtensorMap map; bool ok = false; };
  mutable std::mutex wmu;
  mutable std::map<int, WTiles> wtiles;
};

namespace shapy {
constexpr float kPoseScale = 1024.0f;       // posedirs entries (~1e-3) are scaled into fp16's normal range

// Fused tcgen05 LBS (smplx_lbs.cu).  Returns SHAPY_ERR_UNSUPPORTED when the configuration is outside what the fused
// kernel handles (the caller then runs the three-kernel path).
int launch_lbs_fused(const shapy_smplx *m, const float *betas, const float *rot, int n_rot, int B, float *vertices,
                     float *v_shaped, float *joints, int *lut, cudaStream_t st);

}  // namespace shapy
