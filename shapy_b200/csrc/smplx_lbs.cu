// Fused SMPL-X linear blend skinning for sm_100a: ONE launch evaluates, for a batch of bodies,
//   joints   J = J_template + J_dirs . beta                      (regressor pre-contracted with the shape basis)
//   pose     relative transforms A_j of the kinematic chain       (lbs.py:242-295 batch_rigid_transform)
//   shape    v_shaped = T + S . beta                              (lbs.py:218-239 blend_shapes)
//   pose blend  v_posed = v_shaped + (R[1:] - I) . P              (lbs.py:157-166)  <- tcgen05, fp32-grade
//   skinning    vertices = sum_j w_vj A_j [v_posed; 1]            (lbs.py:176-190)
// (reference regressor/human_shape/models/body_models/lbs.py:99-196), followed by smplx_joints_kernel (landmarks,
// J14 overwrite, camera) as a programmatic dependent launch.
//
// The pose blend is the only dense contraction of the chain -- (bodies x 9(J-1)) . (9(J-1) x 3V), 30 MFLOP per body --
// and the only part that is not a stream: in fp32 SIMT it alone costs more than the HBM time of the whole kernel.  It
// runs on the tensor cores as a split-fp16 GEMM (hi.hi + hi.lo + lo.hi, fp32 accumulate: ~22 mantissa bits, the same
// scheme as the HRNet convolutions):
//   A operand  pose basis P, re-laid out once at model load as [k-block][coordinate plane][vertex][hi 32 k | lo 32 k]
//              fp16 (K-major, scaled by 2^10 into fp16's normal range): one contiguous 16 KB block per (k-block,
//              plane, vertex tile), fetched by ONE TMA box per k-block into 128-byte SWIZZLE_128B rows
//   B operand  pose features of a group of 32 bodies, written by the CTA itself into the swizzled K-major layout
//   D          TMEM: lane = vertex (128 per tile), column = body; three planes (x, y, z) of [D0 | D1] per tile
// so that in the epilogue a thread owns ONE vertex: its template / shape-basis / skinning-weight constants are loaded
// once (coalesced) and reused for every body of the group, the bodies' A_j live in shared memory, and each warp
// writes 32 consecutive vertices of a body as aligned 128-bit rows (transposed through shared memory).
//
// Work item = (group of 32 bodies, tile of 128 vertices); persistent CTAs own contiguous item ranges so that the
// group prologue (betas, pose features, kinematic chain: ~2 us of SIMT work) is paid once per group and CTA.
// Warp roles: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..17 prologue + epilogue.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "smplx.cuh"
#include "umma.cuh"

namespace shapy {

constexpr int LV = 128;            // vertices per tile = MMA M = TMEM lanes
constexpr int LG = 32;             // bodies per group = columns of D0 (and of D1)
constexpr int LKB = 32;            // k per block: 64-byte rows, SWIZZLE_64B
constexpr int LStages = 2;
constexpr int kLbsEpi = 16;        // prologue / epilogue warps (multiple of 4: TMEM lane-quarter rule)
constexpr int kLbsThreads = 64 + 32 * kLbsEpi;
constexpr int kLbsMaxKB = 6;       // k-blocks of the pose feature that fit next to everything else (n_rot <= 22: SHAPY)
constexpr int kLbsNB = 10;         // shape coefficients held in registers
constexpr int kLbsPairs = (LG * kMaxJoints + 32 * kLbsEpi - 1) / (32 * kLbsEpi);   // (body, joint) pairs per prologue thread
constexpr uint32_t kPlaneBytes = LV * 128;                // 16 KB: one coordinate plane of a k-block: 128 rows x [hi 64 B | lo 64 B]
constexpr uint32_t kStageBytes = 3 * kPlaneBytes;         // 48 KB: one TMA box
constexpr uint32_t kCoefBlkBytes = 2 * LG * LKB * 2;      // 4 KB: [C_hi rows | C_lo rows] of one k-block
constexpr float kLoInvL = 1.0f / 2048.0f;

struct alignas(64) LbsParams {
  CUtensorMap basis;
  SmplxDev m;
  const float *betas, *rot;
  int n_rot, B, Kp, nkb, n_vt, n_items;
  float *vertices, *v_shaped, *joints;
  int *lut;
  uint32_t idesc64, idesc32;
  long long *dbg;   // optional [gridDim.x][32] cycle stamps (SHAPY_LBS_DEBUG=1)
};

// byte offset of element (row r, k kk) inside a [rows][32 k] fp16 block laid out K-major with SWIZZLE_64B
// (Swizzle<2,4,3>: address bits [5:4] ^= bits [8:7]); the block base is 1024-byte aligned
__device__ __forceinline__ uint32_t sw64_off(int r, int kk) {
  return (uint32_t)(r * 64 + ((((kk >> 3) ^ (r >> 1)) & 3) << 4) + (kk & 7) * 2);
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kLbsEpi) : "memory"); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}

__global__ void __launch_bounds__(kLbsThreads, 1) smplx_lbs_kernel(const __grid_constant__ LbsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmplxDev &m = p.m;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: [stages][coef k-blocks][A_j][betas][row staging][barriers]
  const uint32_t stage0 = smem_base;
  const uint32_t coef0 = stage0 + LStages * kStageBytes;
  const uint32_t aj_off = LStages * kStageBytes + kLbsMaxKB * kCoefBlkBytes;
  float *Aj = reinterpret_cast<float *>(smem_gen + aj_off);                        // [LG][J][12]
  const uint32_t betas_off = aj_off + (uint32_t)(LG * m.J * 12 * 4);
  float *betas_s = reinterpret_cast<float *>(smem_gen + betas_off);                // [LG][12]
  const uint32_t rowst_off = betas_off + LG * 12 * 4;
  float *rowst = reinterpret_cast<float *>(smem_gen + rowst_off);                  // [kLbsEpi][2][100]
  const uint32_t jtab_off = rowst_off + kLbsEpi * 2 * 100 * 4;
  float *jtab = reinterpret_cast<float *>(smem_gen + jtab_off);                    // [J * 3][12]: J_dirs[:, :NB] | J_template
  const uint32_t bar_base = smem_base + ((jtab_off + (uint32_t)(m.J * 3 * 12 * 4) + 15u) & ~15u);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (LStages + s); };
  const uint32_t acc_full0 = bar_base + 16u * LStages;   // [2]
  const uint32_t acc_empty0 = acc_full0 + 16u;           // [2]
  const uint32_t coef_full = acc_empty0 + 16u;
  const uint32_t tmem_slot = coef_full + 8u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // kinematic-tree tables (a few hundred bytes): read from shared memory inside the level loop
  __shared__ int s_parents[kMaxJoints], s_level_joints[kMaxJoints], s_level_off[kMaxJoints + 1];
  for (int i = threadIdx.x; i < m.J; i += blockDim.x) { s_parents[i] = m.parents[i]; s_level_joints[i] = m.level_joints[i]; }
  for (int i = threadIdx.x; i <= m.n_levels; i += blockDim.x) s_level_off[i] = m.level_off[i];
  for (int i = threadIdx.x; i < m.J * 3 * 12; i += blockDim.x) {
    const int r = i / 12, l = i % 12;
    jtab[i] = l < m.NB ? m.J_dirs[(size_t)r * m.NC + l] : (l == m.NB ? m.J_template[r] : 0.f);
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < LStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full0 + 8u * i, 1); mbar_init(acc_empty0 + 8u * i, kLbsEpi); }
    mbar_init(coef_full, kLbsEpi);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_launch_dependents();
  pdl_wait();
  const long long c0 = clock64();
  auto stamp = [&](int slot) { if (p.dbg) p.dbg[(size_t)blockIdx.x * 32 + slot] = clock64() - c0; };

  // contiguous item range of this CTA; item id = group * n_vt + vertex tile
  const int it_lo = (int)((long long)p.n_items * blockIdx.x / gridDim.x);
  const int it_hi = (int)((long long)p.n_items * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0) {
    // ===================================================================== TMA producer: pose-basis k-blocks
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = it_lo; it < it_hi; ++it) {
        const int vt = it % p.n_vt;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), kStageBytes);
          const uint32_t sb = stage0 + s * kStageBytes;
          tma_load_4d(sb, &p.basis, full_bar(s), 0, vt * LV, 0, kb);
          if (++s == LStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int s = 0, lt = 0, gprev = -1, gchanges = 0;
    uint32_t ph = 0;
    for (int it = it_lo; it < it_hi; ++it, ++lt) {
      const int g = it / p.n_vt;
      if (g != gprev) {   // the epilogue warps rebuild the pose features of the new group
        mbar_wait(coef_full, (uint32_t)(gchanges & 1));
        ++gchanges;
        gprev = g;
        if (lane == 0 && lt == 0) stamp(16);
      }
      const int buf = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      mbar_wait(acc_empty0 + 8u * buf, aph ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t dhi = desc_hi_swz<LKB>();       // B operand: 64-byte rows, SWIZZLE_64B
        const uint32_t ahi = desc_hi_swz<64>();        // A operand: 128-byte rows [hi | lo], SWIZZLE_128B
        const uint32_t dbase = tmem_base + buf * 256u;
        int s_l = s;
        uint32_t ph_l = ph;
#pragma unroll 1
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(full_bar(s_l), ph_l);
          tc_fence_after();
          const uint32_t a16 = desc_lo_swz(stage0 + s_l * kStageBytes);
          const uint32_t b16 = desc_lo_swz(coef0 + kb * kCoefBlkBytes);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const uint32_t ac = a16 + c * (kPlaneBytes >> 4), d = dbase + c * 64u;
#pragma unroll
            for (int ks = 0; ks < LKB / 16; ++ks) {
              // [D0 | D1] (+)= P_hi . [C_hi | C_lo]^T ; D1 += P_lo . C_hi^T
              umma_f16_lh(d, ac + 2 * ks, ahi, b16 + 2 * ks, dhi, p.idesc64, (kb | ks) ? 1u : 0u);
              umma_f16_lh(d + LG, ac + 4 + 2 * ks, ahi, b16 + 2 * ks, dhi, p.idesc32, 1u);   // lo half of the row: +64 B
            }
          }
          umma_commit(empty_bar(s_l));
          if (++s_l == LStages) { s_l = 0; ph_l ^= 1; }
        }
        umma_commit(acc_full0 + 8u * buf);
        if (lt < 2) stamp(17 + lt);
      }
      __syncwarp();
      for (int kb = 0; kb < p.nkb; ++kb) { if (++s == LStages) { s = 0; ph ^= 1; } }
    }
  } else {
    // ===================================================================== prologue + epilogue warps
    const int ew = warp - 2, q = warp & 3, part = ew >> 2;       // TMEM lane quarter, body sub-range
    const int et = threadIdx.x - 64;                              // 0 .. 511
    const int J = m.J;
    float *myst = rowst + (size_t)ew * 200;                       // two 100-float rows (vertices, v_shaped)
    int lt = 0, gprev = -1;
    for (int it = it_lo; it < it_hi; ++it, ++lt) {
      const int g = it / p.n_vt, vt = it % p.n_vt;
      const int b0 = g * LG, nb = min(LG, p.B - b0);
      if (g != gprev) {
        gprev = g;
        epi_bar();                       // every warp is done with the previous group's A_j / betas
        // ---- betas and rotations of the group -> shared memory, coalesced (the rotations of 32 bodies are one
        // contiguous range; they are staged in the A_j region, which is not written before the level loop below)
        float *rs = Aj;                                  // [nb][n_rot * 9]
        const int rper = p.n_rot * 9;
        for (int i = et; i < LG * 12; i += 32 * kLbsEpi) {
          const int bl = i / 12, l = i % 12;
          betas_s[i] = (bl < nb && l < m.NB) ? __ldg(p.betas + (size_t)(b0 + bl) * m.NB + l) : 0.f;
        }
        {
          const float *src = p.rot + (size_t)b0 * rper;
          for (int i = et; i < nb * rper; i += 32 * kLbsEpi) rs[i] = __ldg(src + i);
        }
        epi_bar();
        // ---- pose features (R[1:] - I) as fp16 hi / lo rows of the B operand
        {
          const int per_body = p.nkb * LKB;
          for (int i = et; i < LG * per_body; i += 32 * kLbsEpi) {
            const int bl = i / per_body, k = i - bl * per_body;
            float f = 0.f;
            if (bl < nb && k < p.Kp) {
              const int e = k % 9;
              f = rs[bl * rper + 9 + k] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
            }
            const __half h = __float2half_rn(f);
            const __half l = __float2half_rn((f - __half2float(h)) * 2048.0f);
            uint8_t *blk = smem_gen + (coef0 - smem_base) + (k / LKB) * kCoefBlkBytes;
            *reinterpret_cast<__half *>(blk + sw64_off(bl, k % LKB)) = h;
            *reinterpret_cast<__half *>(blk + sw64_off(LG + bl, k % LKB)) = l;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(coef_full);
        if (et == 0 && lt == 0) stamp(1);
        // ---- kinematic chain, level by level; only A is kept:
        //   A_c.R = A_p.R R_c ;  A_c.t = A_p.R J_c + A_p.t - A_c.R J_c   (== G_c - [0 | G_c.R J_c] of lbs.py:279-293,
        //   since G_c.t = G_p.R (J_c - J_p) + G_p.t and A_p.t = G_p.t - G_p.R J_p)
        // Thread et owns the (body, joint) pairs et, et + 512, ... in LEVEL order (pair id = level-ordered joint slot
        // * 32 + body), at most kLbsPairs of them.  A pair's rotation and rest joint J = J_template + J_dirs . beta
        // are read from shared memory into registers BEFORE the level loop (which overwrites the staged rotations with
        // A_j).  (A first version read J_dirs and the rotations from global memory inside the level loop: ten levels
        // of serialised L2 latency, 70 us per group; a second one preloaded them from global memory with one sector
        // per lane and was bound by the load unit.)
        float Rp[kLbsPairs][9], Jp[kLbsPairs][3];
#pragma unroll
        for (int k = 0; k < kLbsPairs; ++k) {
          const int id = et + k * 32 * kLbsEpi;
          const int bl = id % LG, slot = id / LG;
          const bool on = slot < J && bl < nb;
          const int j = on ? s_level_joints[slot] : 0;
#pragma unroll
          for (int e = 0; e < 9; ++e)
            Rp[k][e] = (on && j < p.n_rot) ? rs[bl * rper + j * 9 + e] : ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
          const float *bt = betas_s + bl * 12;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float *jd = jtab + (j * 3 + c) * 12;
            float sacc = jd[kLbsNB];
#pragma unroll
            for (int l = 0; l < kLbsNB; ++l) sacc += jd[l] * bt[l];
            Jp[k][c] = sacc;
          }
        }
        // dynamic-contour LUT row (lbs.py:30-41, rotation_utils.py:86-92), once per group, from the staged rotations
        if (vt == 0 && m.D > 0 && p.lut && et < nb) {
          float rel[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
          for (int qn = 0; qn < m.n_chain; ++qn) {
            const int j = m.neck[qn];
            float Rq[9];
            for (int e = 0; e < 9; ++e) Rq[e] = j < p.n_rot ? rs[et * rper + j * 9 + e] : ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
            float o[9];
            for (int r = 0; r < 3; ++r)
              for (int c = 0; c < 3; ++c) o[r * 3 + c] = Rq[r * 3] * rel[c] + Rq[r * 3 + 1] * rel[3 + c] + Rq[r * 3 + 2] * rel[6 + c];
            for (int e = 0; e < 9; ++e) rel[e] = o[e];
          }
          const float sy = sqrtf(rel[0] * rel[0] + rel[3] * rel[3]);
          const float ang = atan2f(-rel[6], sy);
          const float deg = fminf(-ang * 180.0f / 3.14159265358979323846f, 39.0f);
          const int y = (int)rintf(deg);
          const int row = y < 0 ? (y < -39 ? 78 : 39 - y) : y;
          p.lut[b0 + et] = min(max(row, 0), m.rows - 1);
        }
        epi_bar();                       // every thread has its rotations in registers: the A_j region may be written
        if (et == 0 && lt == 0) stamp(2);
        for (int lv = 0; lv < m.n_levels; ++lv) {
          const int off = s_level_off[lv], end = s_level_off[lv + 1];
#pragma unroll
          for (int k = 0; k < kLbsPairs; ++k) {
            const int id = et + k * 32 * kLbsEpi;
            const int bl = id % LG, slot = id / LG;
            if (slot < off || slot >= end || bl >= nb) continue;
            const int j = s_level_joints[slot];
            const float *R = Rp[k], *Jc = Jp[k];
            float *Ao = Aj + ((size_t)bl * J + j) * 12;
            if (lv == 0) {
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                Ao[r * 4 + 0] = R[r * 3]; Ao[r * 4 + 1] = R[r * 3 + 1]; Ao[r * 4 + 2] = R[r * 3 + 2];
                Ao[r * 4 + 3] = Jc[r] - (R[r * 3] * Jc[0] + R[r * 3 + 1] * Jc[1] + R[r * 3 + 2] * Jc[2]);
              }
            } else {
              const float *Ap = Aj + ((size_t)bl * J + s_parents[j]) * 12;
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const float g0 = Ap[r * 4], g1 = Ap[r * 4 + 1], g2 = Ap[r * 4 + 2], gt = Ap[r * 4 + 3];
                const float n0 = g0 * R[0] + g1 * R[3] + g2 * R[6];
                const float n1 = g0 * R[1] + g1 * R[4] + g2 * R[7];
                const float n2 = g0 * R[2] + g1 * R[5] + g2 * R[8];
                Ao[r * 4 + 0] = n0; Ao[r * 4 + 1] = n1; Ao[r * 4 + 2] = n2;
                Ao[r * 4 + 3] = (g0 * Jc[0] + g1 * Jc[1] + g2 * Jc[2]) + gt - (n0 * Jc[0] + n1 * Jc[1] + n2 * Jc[2]);
              }
            }
            // posed joint = G_j.t = A_j.t + A_j.R J_j: written once per group (by the CTA that owns vertex tile 0)
            if (vt == 0 && p.joints) {
              float *jo = p.joints + ((size_t)(b0 + bl) * m.K + j) * 3;
#pragma unroll
              for (int r = 0; r < 3; ++r)
                jo[r] = Ao[r * 4 + 3] + (Ao[r * 4] * Jc[0] + Ao[r * 4 + 1] * Jc[1] + Ao[r * 4 + 2] * Jc[2]);
            }
          }
          epi_bar();
        }
        if (et == 0 && lt == 0) stamp(3);
      }
      // ================================================================= epilogue of this item
      const int v0w = vt * LV + q * 32;            // first vertex of this warp's lane quarter
      const int v = v0w + lane;
      const bool vok = v < m.V;
      const int vc = vok ? v : m.V - 1;
      // per-vertex constants (coalesced across the warp), reused by every body of the group
      float T3[3], S[kLbsNB][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) T3[c] = m.shape_planes[(size_t)(kLbsNB * 3 + c) * m.Vpad + vc];
#pragma unroll
      for (int l = 0; l < kLbsNB; ++l)
#pragma unroll
        for (int c = 0; c < 3; ++c) S[l][c] = m.shape_planes[(size_t)(l * 3 + c) * m.Vpad + vc];
      // skinning weights of this vertex: the first four ELL slots live in registers for the whole group
      float ew[4];
      int ej[4];
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const bool on = vok && sl < m.ell_w_n;
        ew[sl] = on ? __ldg(m.ell_w + (size_t)sl * m.V + v) : 0.f;
        ej[sl] = on ? __ldg(m.ell_idx + (size_t)sl * m.V + v) : 0;
      }
      const int buf = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      if (et == 0 && lt < 2) stamp(4 + 3 * lt);
      mbar_wait(acc_full0 + 8u * buf, aph);
      if (et == 0 && lt < 2) stamp(5 + 3 * lt);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16) + part * 8;
      const int rows_here = min(32, m.V - v0w);      // > 0 for every tile (V > (n_vt - 1) * 128 + 96 is NOT assumed)
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t d0[3][4], d1[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          tmem_ld4(lane_addr + c * 64 + half * 4, d0[c]);
          tmem_ld4(lane_addr + c * 64 + LG + half * 4, d1[c]);
        }
        tmem_ld_wait();
        if (half == 1) {   // last TMEM read of this item: hand the accumulator buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty0 + 8u * buf);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int bl = part * 8 + half * 4 + i;
          if (bl >= nb) break;                                   // warp-uniform
          const int b = b0 + bl;
          const float4 q0 = *reinterpret_cast<const float4 *>(betas_s + bl * 12);
          const float4 q1 = *reinterpret_cast<const float4 *>(betas_s + bl * 12 + 4);
          const float4 q2 = *reinterpret_cast<const float4 *>(betas_s + bl * 12 + 8);
          const float bb[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
          float vs[3] = {T3[0], T3[1], T3[2]};
#pragma unroll
          for (int l = 0; l < kLbsNB; ++l) {
            vs[0] += bb[l] * S[l][0]; vs[1] += bb[l] * S[l][1]; vs[2] += bb[l] * S[l][2];
          }
          float vp[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            vp[c] = vs[c] + (__uint_as_float(d0[c][i]) + __uint_as_float(d1[c][i]) * kLoInvL) * (1.0f / kPoseScale);
          }
          float o3[3] = {0.f, 0.f, 0.f};
          if (vok) {
            for (int sl = 0; sl < m.ell_w_n; ++sl) {
              const float w = sl < 4 ? (sl == 0 ? ew[0] : (sl == 1 ? ew[1] : (sl == 2 ? ew[2] : ew[3]))) : m.ell_w[(size_t)sl * m.V + v];
              if (w == 0.f) continue;
              const int j = sl < 4 ? (sl == 0 ? ej[0] : (sl == 1 ? ej[1] : (sl == 2 ? ej[2] : ej[3]))) : m.ell_idx[(size_t)sl * m.V + v];
              const float4 *Ab = reinterpret_cast<const float4 *>(Aj + ((size_t)bl * J + j) * 12);
              const float4 r0 = Ab[0], r1 = Ab[1], r2 = Ab[2];
              o3[0] += w * (r0.x * vp[0] + r0.y * vp[1] + r0.z * vp[2] + r0.w);
              o3[1] += w * (r1.x * vp[0] + r1.y * vp[1] + r1.z * vp[2] + r1.w);
              o3[2] += w * (r2.x * vp[0] + r2.y * vp[1] + r2.z * vp[2] + r2.w);
            }
          }
          // ---- transposed row store: the warp's 32 vertices x 3 coordinates are 96 consecutive floats of body b;
          // they are staged at the row's offset inside its 16-byte grid so that lanes 0..24 issue aligned 128-bit
          // stores (head / tail floats of a partially covered chunk are stored one by one)
          const size_t o = ((size_t)b * m.V + v0w) * 3;
          const int mis = (int)(o & 3), nfl = rows_here * 3;
          __syncwarp();
          if (vok) {
            myst[mis + 3 * lane] = o3[0]; myst[mis + 3 * lane + 1] = o3[1]; myst[mis + 3 * lane + 2] = o3[2];
            myst[100 + mis + 3 * lane] = vs[0]; myst[100 + mis + 3 * lane + 1] = vs[1]; myst[100 + mis + 3 * lane + 2] = vs[2];
          }
          __syncwarp();
          const int f0 = 4 * lane;                   // stage index of this lane's chunk
          if (f0 < mis + nfl && f0 + 4 > mis) {
            float *gv = p.vertices + (o - mis) + f0;
            float *gs = p.v_shaped ? p.v_shaped + (o - mis) + f0 : nullptr;
            if (f0 >= mis && f0 + 4 <= mis + nfl) {
              *reinterpret_cast<float4 *>(gv) = *reinterpret_cast<const float4 *>(myst + f0);
              if (gs) *reinterpret_cast<float4 *>(gs) = *reinterpret_cast<const float4 *>(myst + 100 + f0);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (f0 + e >= mis && f0 + e < mis + nfl) {
                  gv[e] = myst[f0 + e];
                  if (gs) gs[e] = myst[100 + f0 + e];
                }
            }
          }
        }
      }
      if (et == 0 && lt < 2) stamp(6 + 3 * lt);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(31);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

size_t lbs_smem_bytes(int J) {
  return 1024 + LStages * kStageBytes + kLbsMaxKB * kCoefBlkBytes + (size_t)LG * J * 12 * 4 + LG * 12 * 4 +
         kLbsEpi * 2 * 100 * 4 + (size_t)J * 3 * 12 * 4 + 16 + 128;   // + 1 KB of static tables (s_parents ...)
}

int launch_lbs_fused(const shapy_smplx *mm, const float *betas, const float *rot, int n_rot, int B, float *vertices,
                     float *v_shaped, float *joints, int *lut, cudaStream_t st) {
  const SmplxDev &d = mm->d;
  const int Kp = (n_rot - 1) * 9, nkb = std::max(1, ceil_div(Kp, LKB));
  if (!mm->fused_ok || !vertices || d.NB != kLbsNB || nkb > kLbsMaxKB || lbs_smem_bytes(d.J) > 226 * 1024)
    return SHAPY_ERR_UNSUPPORTED;
  static const bool off = []() { const char *e = getenv("SHAPY_LBS_FUSED"); return e && e[0] == '0'; }();
  if (off) return SHAPY_ERR_UNSUPPORTED;
  LbsParams p;
  memset(&p, 0, sizeof(p));
  p.basis = mm->basis_map;
  p.m = d;
  p.betas = betas; p.rot = rot; p.n_rot = n_rot; p.B = B; p.Kp = Kp; p.nkb = nkb;
  p.n_vt = d.Vpad / LV;
  p.n_items = ceil_div(B, LG) * p.n_vt;
  p.vertices = vertices; p.v_shaped = v_shaped; p.joints = joints; p.lut = lut;
  p.idesc64 = (1u << 4) | ((uint32_t)((2 * LG) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.idesc32 = (1u << 4) | ((uint32_t)(LG >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static std::atomic<unsigned long long> attr_done{0};
  SHAPY_CUDA_TRY(set_max_dynamic_smem(smplx_lbs_kernel, 226 * 1024, attr_done));   // + 1 KB static <= 227 KB
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(p.n_items, sms);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kLbsThreads);
  cfg.dynamicSmemBytes = lbs_smem_bytes(d.J);
  cfg.stream = st;
  static const bool dbg = getenv("SHAPY_LBS_DEBUG") != nullptr;
  if (dbg) {   // synchronous: per-CTA cycle stamps of the roles (prologue / MMA / epilogue phases)
    long long *d = nullptr;
    cudaMalloc(&d, (size_t)grid * 32 * 8);
    cudaMemset(d, 0, (size_t)grid * 32 * 8);
    p.dbg = d;
    SHAPY_CUDA_TRY(cudaLaunchKernelEx(&cfg, smplx_lbs_kernel, p));
    cudaStreamSynchronize(st);
    std::vector<long long> h((size_t)grid * 32);
    cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    const char *names[32] = {"", "coef_ready", "pairs_loaded", "chain_done", "it0_consts", "it0_acc_full", "it0_epi_done",
                             "it1_consts", "it1_acc_full", "it1_epi_done", "", "", "", "", "", "", "mma_coef_wait",
                             "mma_it0_issued", "mma_it1_issued", "", "", "", "", "", "", "", "", "", "", "", "", "exit"};
    fprintf(stderr, "[lbs] B %d items %d grid %d: cycles since kernel entry (avg / max over CTAs that reached the point)\n", B, p.n_items, grid);
    for (int k = 0; k < 32; ++k) {
      if (!names[k][0]) continue;
      double sum = 0; long long mx = 0; int n = 0;
      for (int c = 0; c < grid; ++c) { long long v = h[(size_t)c * 32 + k]; if (v > 0) { sum += v; mx = std::max(mx, v); ++n; } }
      if (n) fprintf(stderr, "[lbs]   %-16s avg %8.0f max %8lld (n=%d)\n", names[k], sum / n, mx, n);
    }
    count_launch();
    return SHAPY_OK;
  }
  SHAPY_CUDA_TRY(cudaLaunchKernelEx(&cfg, smplx_lbs_kernel, p));
  count_launch();
  return SHAPY_OK;
}

}  // namespace shapy
