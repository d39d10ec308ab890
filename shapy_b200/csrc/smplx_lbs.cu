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
// so that in the epilogue a thread owns ONE vertex: its template / shape-basis constants are loaded once (coalesced) and
// reused for every body of the group.
//
// The skinning sum is linear in A_j, so the per-vertex transform T_v = sum_j w_vj A_j is a second, small GEMM (TV = true,
// the default): A operand = the skinning weights folded onto the rotated joints, [vertex][hi 32 j | lo 32 j] fp16 x 2^10
// (one TMA box per vertex tile), B operand = the group's A_j x 64 as fp16 hi / lo rows (body, component) written by the
// CTA after the chain, D = 12 fp32 columns per (vertex, body), produced in 8 chunks of 4 bodies through two chunk
// accumulators that the epilogue hands back as it reads them.  The epilogue is then three tcgen05.ld and a 3 x 4
// matrix-vector product per vertex-body instead of up to 12 LDS.128 + 48 FMA (TV = false keeps that path: A_j of all
// joints in shared memory, ELL skinning weights).  Each warp writes 32 consecutive vertices of a body as three fully
// coalesced 128-byte stores per output array (transposed through shared memory).
//
// Work item = (group of 32 bodies, tile of 128 vertices); persistent CTAs own contiguous item ranges so that the
// group prologue (betas, pose features, kinematic chain: ~13 us of mostly latency) is paid once per group and CTA.
// Warp roles (20 warps): warp 0 TMA producer, warp 1 TMEM allocator + one-thread MMA issuer polling both GEMMs' queues,
// warps 2-3 idle (they complete the control warpgroup so that setmaxnreg can move its registers), warps 4..19 prologue
// + epilogue with 112 registers each.
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
constexpr int kLbsCtl = 4;         // control warps: TMA producer, MMA issuer, two idle (a warpgroup of its own, see setmaxnreg)
constexpr int kLbsThreads = 32 * (kLbsCtl + kLbsEpi);
// The register file is 16 K registers per SM sub-partition and warps go to sub-partitions round robin: 5 warps per
// sub-partition cap a uniform allocation at 96 registers, where the epilogue rematerialises addresses all over the place
// (ncu r02: 13 % of its instructions).  The control warpgroup hands its registers over instead (setmaxnreg acts on
// whole warpgroups, which is why the two roles that need few registers got a warpgroup to themselves):
// 32 + 4 x 112 = 480 = the 5 x 96 the sub-partition's warps were launched with (setmaxnreg only moves registers inside
// the CTA's launch allocation: asking for more than the control warps released blocks forever).
constexpr int kLbsCtlRegs = 32, kLbsEpiRegs = 112;
// TMEM columns: two GEMM-1 accumulators at [0, 192) and [256, 448), two blend-GEMM chunk accumulators at [192, 240) and
// [448, 496).  (Tried: one GEMM-1 accumulator that the epilogue takes into registers up front plus six chunk
// accumulators, so that chunks are issued far ahead: 857 instead of 777 us at B = 4 096 -- GEMM 1 of the next item then
// starts later, and it is paced by TMA at ~20 B/cycle/SM with the L2 at two thirds of its throughput cap.)
constexpr int kTvBufs = 2;
constexpr int kLbsMaxKB = 6;       // k-blocks of the pose feature that fit next to everything else (n_rot <= 22: SHAPY)
constexpr int kLbsNB = 10;         // shape coefficients held in registers
constexpr int kLbsPairs = (LG * kMaxJoints + 32 * kLbsEpi - 1) / (32 * kLbsEpi);   // (body, joint) pairs per prologue thread
constexpr uint32_t kPlaneBytes = LV * 128;                // 16 KB: one coordinate plane of a k-block: 128 rows x [hi 64 B | lo 64 B]
constexpr uint32_t kStageBytes = 3 * kPlaneBytes;         // 48 KB: one TMA box
constexpr uint32_t kCoefBlkBytes = 2 * LG * LKB * 2;      // 4 KB: [C_hi rows | C_lo rows] of one k-block
constexpr float kLoInvL = 1.0f / 2048.0f;
// TV variant: the skinning transform blend T_v = sum_j w_vj A_j as a second tcgen05 GEMM
constexpr int kTvJ = 32;                                   // joint columns (K) of the blend GEMM: rotated joints, padded
constexpr int kTvChunk = 4;                                // bodies per chunk (one from each epilogue part): N = 48
constexpr int kTvN = kTvChunk * 12;
constexpr uint32_t kTvBlkBytes = kTvN * kTvJ * 2;          // 3 072 B: [48 rows][32 joints] fp16, 64-byte rows (SWIZZLE_64B)
constexpr uint32_t kTvOpBytes = (LG / kTvChunk) * 2 * kTvBlkBytes;   // 8 chunks x (hi, lo) = 49 152 B
constexpr uint32_t kTvWBytes = LV * 128;                   // W tile: 128 rows x [hi 32 | lo 32] fp16 = 16 KB (SWIZZLE_128B)
constexpr float kTvWScale = 1024.0f, kTvAScale = 64.0f;    // operands are scaled so that their fp16 residuals stay normal

struct alignas(64) LbsParams {
  CUtensorMap basis;
  CUtensorMap wmap;           // TV: folded skinning weights (64, Vpad) fp16 [hi 32 | lo 32], box (64, 128), SWIZZLE_128B
  SmplxDev m;
  const float *betas, *rot;
  int n_rot, B, Kp, nkb, n_vt, n_items;
  float *vertices, *v_shaped, *joints;
  int *lut;
  uint32_t idesc64, idesc32, idesc48;
  int dbg_flags;    // SHAPY_LBS_DBGFLAGS experiments: 1 = no output stores, 2 = blocking producer waits
  int dbg_lt;       // first of the two items per CTA whose epilogue is stamped
  long long *dbg;   // optional [gridDim.x][32] cycle stamps (SHAPY_LBS_DEBUG=1)
};

// byte offset of element (row r, k kk) inside a [rows][32 k] fp16 block laid out K-major with SWIZZLE_64B
// (Swizzle<2,4,3>: address bits [5:4] ^= bits [8:7]); the block base is 1024-byte aligned
__device__ __forceinline__ uint32_t sw64_off(int r, int kk) {
  return (uint32_t)(r * 64 + ((((kk >> 3) ^ (r >> 1)) & 3) << 4) + (kk & 7) * 2);
}

// wait of a role with slack (the TMA producer): back off between polls instead of competing for issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(64);
  }
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {   // non-blocking
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kLbsEpi) : "memory"); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}

template <bool TV>
__global__ void __launch_bounds__(kLbsThreads, 1) smplx_lbs_kernel(const __grid_constant__ LbsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmplxDev &m = p.m;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: [stages][coef k-blocks][A_j][betas][row staging][barriers]
  const uint32_t stage0 = smem_base;
  const uint32_t coef0 = stage0 + LStages * kStageBytes;
  const uint32_t aj_off = LStages * kStageBytes + kLbsMaxKB * kCoefBlkBytes;
  // !TV: A_j fp32 [LG][J][12].  TV: the same region first stages the rotations, then holds A_j fp32 of the rotated
  // joints [LG][kTvJ][12] during the chain, and finally the B operand of the blend GEMM (written from registers)
  float *Aj = reinterpret_cast<float *>(smem_gen + aj_off);
  const int AJ = TV ? kTvJ : m.J;                                                  // joint stride of the fp32 A_j array
  const uint32_t aj_bytes = TV ? kTvOpBytes : (uint32_t)(LG * m.J * 12 * 4);
  const uint32_t wt_off = aj_off + aj_bytes;                                       // TV: W tile (1 024-aligned: all sizes above are)
  const uint32_t betas_off = wt_off + (TV ? 2u * kTvWBytes : 0u);
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
  const uint32_t acc2_full0 = coef_full + 8u;            // TV: [kTvBufs] blend-GEMM chunk accumulators
  const uint32_t acc2_empty0 = acc2_full0 + 8u * kTvBufs;   // [kTvBufs]
  const uint32_t aop_full = acc2_empty0 + 8u * kTvBufs;  // TV: A_j operand of the group is written
  const uint32_t w_full0 = aop_full + 8u, w_empty0 = w_full0 + 16u;   // TV: [2] skinning-weight tiles
  const uint32_t tmem_slot = w_empty0 + 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // kinematic-tree tables (a few hundred bytes): read from shared memory inside the level loop
  __shared__ unsigned char s_parents[kMaxJoints], s_level_joints[kMaxJoints], s_level_off[kMaxJoints + 1], s_anc[kMaxJoints];
  __shared__ int s_nlv_rot;   // levels that hold a rotated joint (the deeper ones are identity joints only)
  for (int i = threadIdx.x; i < m.J; i += blockDim.x) { s_parents[i] = (unsigned char)max(m.parents[i], 0); s_level_joints[i] = (unsigned char)m.level_joints[i]; }
  if (threadIdx.x == 0) {   // nearest ancestor with its own rotation (joints >= n_rot are identity: A_j == A_anc(j))
    for (int j = 0; j < m.J; ++j) s_anc[j] = (unsigned char)(j < p.n_rot ? j : s_anc[max(m.parents[j], 0)]);
    int nl = 0;
    for (int lv = 0; lv < m.n_levels; ++lv)
      for (int sl = m.level_off[lv]; sl < m.level_off[lv + 1]; ++sl)
        if (m.level_joints[sl] < p.n_rot) nl = lv + 1;
    s_nlv_rot = nl;
  }
  for (int i = threadIdx.x; i <= m.n_levels; i += blockDim.x) s_level_off[i] = (unsigned char)m.level_off[i];
  for (int i = threadIdx.x; i < m.J * 3 * 12; i += blockDim.x) {
    const int r = i / 12, l = i % 12;
    jtab[i] = l < m.NB ? m.J_dirs[(size_t)r * m.NC + l] : (l == m.NB ? m.J_template[r] : 0.f);
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < LStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full0 + 8u * i, 1); mbar_init(acc_empty0 + 8u * i, kLbsEpi); }
    mbar_init(coef_full, kLbsEpi);
    for (int i = 0; i < kTvBufs; ++i) { mbar_init(acc2_full0 + 8u * i, 1); mbar_init(acc2_empty0 + 8u * i, kLbsEpi); }
    mbar_init(aop_full, kLbsEpi);
    for (int i = 0; i < 2; ++i) { mbar_init(w_full0 + 8u * i, 1); mbar_init(w_empty0 + 8u * i, 1); }
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

  if (warp < kLbsCtl) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kLbsCtlRegs));
  if (warp == 0) {
    // ===================================================================== TMA producer: pose-basis k-blocks
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int it = it_lo; it < it_hi; ++it, ++lt) {
        const int vt = it % p.n_vt;
        if (TV) {   // skinning-weight tile of this vertex tile (two buffers: freed by the item's last blend MMA)
          mbar_wait_relaxed(w_empty0 + 8u * (lt & 1), (uint32_t)(((lt >> 1) & 1) ^ 1));
          mbar_expect_tx(w_full0 + 8u * (lt & 1), kTvWBytes);
          tma_load_2d(smem_base + wt_off + (uint32_t)(lt & 1) * kTvWBytes, &p.wmap, w_full0 + 8u * (lt & 1), 0, vt * LV);
        }
        for (int kb = 0; kb < p.nkb; ++kb) {
          if (p.dbg_flags & 2) mbar_wait(empty_bar(s), ph ^ 1);
          else mbar_wait_relaxed(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), kStageBytes);
          const uint32_t sb = stage0 + s * kStageBytes;
          tma_load_4d(sb, &p.basis, full_bar(s), 0, vt * LV, 0, kb);
          if (++s == LStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread)
    // Two in-order work queues share the tensor pipe: GEMM 1 (pose blend) k-blocks, which become ready as TMA delivers
    // them (two in flight: the ring), and -- TV -- the blend-GEMM chunks, which become ready as the epilogue hands the two
    // chunk accumulators back.  A blocking wait on either would stall the other (measured: the epilogue of item t sat
    // 6 000 cycles behind the k-block waits of item t + 1), so the thread polls both and issues whatever is ready.
    if (lane == 0) {
      const uint32_t dhi = desc_hi_swz<LKB>();       // B operand of GEMM 1: 64-byte rows, SWIZZLE_64B
      const uint32_t ahi = desc_hi_swz<64>();        // A operands: 128-byte rows [hi | lo], SWIZZLE_128B
      const uint32_t ohi = desc_hi_swz<kTvJ>();
      const uint32_t o16 = desc_lo_swz(smem_base + aj_off);
      int s = 0;                                     // ring position: k-blocks are consumed in issue order
      uint32_t ph = 0;
      int g_it = it_lo, g_kb = 0, g_gprev = -1, g_gch = 0;   // GEMM-1 cursor
      bool g_open = false;                                   // accumulator of item g_it acquired
      int c_it = it_lo, c_c = 0, c_gprev = -1, c_gch = 0;    // chunk cursor
      int c_cb = 0;                                          // chunk accumulator (round robin) and its phase
      uint32_t c_cph = 0;
      bool c_open = false;                                   // operands of item c_it are there
      while (g_it < it_hi || (TV && c_it < it_hi)) {
        bool progressed = false;
        if (g_it < it_hi) {
          const int lt = g_it - it_lo, g = g_it / p.n_vt, buf = lt & 1;
          const uint32_t eph = (uint32_t)(((lt >> 1) & 1) ^ 1);
          if (!g_open) {
            bool ok = true;
            if (g != g_gprev) {   // the epilogue warps rebuild the pose features of a new group after its last epilogue
              ok = mbar_test(coef_full, (uint32_t)(g_gch & 1));
              if (ok) { ++g_gch; g_gprev = g; if (lt == 0) stamp(16); }
            }
            if (ok && mbar_test(acc_empty0 + 8u * buf, eph)) {
              g_open = true;
              tc_fence_after();
              if ((unsigned)(lt - p.dbg_lt) < 2u) stamp(13 + lt - p.dbg_lt);
            }
          }
          if (g_open && mbar_test(full_bar(s), ph)) {
            tc_fence_after();
            const uint32_t a16 = desc_lo_swz(stage0 + s * kStageBytes);
            const uint32_t b16 = desc_lo_swz(coef0 + g_kb * kCoefBlkBytes);
            const uint32_t dbase = tmem_base + buf * 256u;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const uint32_t ac = a16 + c * (kPlaneBytes >> 4), d = dbase + c * 64u;
#pragma unroll
              for (int ks = 0; ks < LKB / 16; ++ks) {
                // [D0 | D1] (+)= P_hi . [C_hi | C_lo]^T ; D1 += P_lo . C_hi^T
                umma_f16_lh(d, ac + 2 * ks, ahi, b16 + 2 * ks, dhi, p.idesc64, (g_kb | ks) ? 1u : 0u);
                umma_f16_lh(d + LG, ac + 4 + 2 * ks, ahi, b16 + 2 * ks, dhi, p.idesc32, 1u);   // lo half of the row: +64 B
              }
            }
            umma_commit(empty_bar(s));
            if (++s == LStages) { s = 0; ph ^= 1; }
            if (++g_kb == p.nkb) {
              umma_commit(acc_full0 + 8u * buf);
              if ((unsigned)(lt - p.dbg_lt) < 2u) stamp(17 + lt - p.dbg_lt);
              g_kb = 0; g_open = false; ++g_it;
            }
            progressed = true;
          }
        }
        // blend GEMM chunk c of item c_it: T[v][(part, 12)] = W'[v][j] . A_j[body = part * 8 + c][12]
        if (TV && c_it < g_it) {
          const int lt = c_it - it_lo, g = c_it / p.n_vt;
          if (!c_open) {
            bool ok = true;
            if (g != c_gprev) {
              ok = mbar_test(aop_full, (uint32_t)(c_gch & 1));
              if (ok) { ++c_gch; c_gprev = g; }
            }
            if (ok && mbar_test(w_full0 + 8u * (lt & 1), (uint32_t)((lt >> 1) & 1))) { c_open = true; tc_fence_after(); }
          }
          if (c_open && mbar_test(acc2_empty0 + 8u * c_cb, c_cph ^ 1)) {
            tc_fence_after();
            const uint32_t w16 = desc_lo_swz(smem_base + wt_off + (uint32_t)(lt & 1) * kTvWBytes);
            const uint32_t d = tmem_base + (c_cb ? 448u : 192u);
            const uint32_t bh = o16 + (uint32_t)c_c * ((2 * kTvBlkBytes) >> 4), bl = bh + (kTvBlkBytes >> 4);
#pragma unroll
            for (int ks = 0; ks < kTvJ / 16; ++ks) {
              umma_f16_lh(d, w16 + 2 * ks, ahi, bh + 2 * ks, ohi, p.idesc48, ks ? 1u : 0u);       // W_hi . A_hi
              umma_f16_lh(d, w16 + 2 * ks, ahi, bl + 2 * ks, ohi, p.idesc48, 1u);                 // W_hi . A_lo
              umma_f16_lh(d, w16 + 4 + 2 * ks, ahi, bh + 2 * ks, ohi, p.idesc48, 1u);             // W_lo . A_hi
            }
            umma_commit(acc2_full0 + 8u * c_cb);
            if (++c_cb == kTvBufs) { c_cb = 0; c_cph ^= 1; }
            if (lt == p.dbg_lt && c_c == 0) stamp(12);
            if (++c_c == LG / kTvChunk) {
              if (lt == p.dbg_lt) stamp(11);
              umma_commit(w_empty0 + 8u * (lt & 1));
              c_c = 0; c_open = false; ++c_it;
            }
            progressed = true;
          }
        }
        if (!progressed) __nanosleep(20);
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kLbsEpiRegs));
    if (threadIdx.x == 32 * kLbsCtl) stamp(15);
    // ===================================================================== prologue + epilogue warps
    const int ew = warp - kLbsCtl, q = warp & 3, part = ew >> 2;  // TMEM lane quarter, body sub-range
    const int et = threadIdx.x - 32 * kLbsCtl;                    // 0 .. 511
    const int J = m.J;
    float *myst = rowst + (size_t)ew * 200;                       // two 100-float rows (vertices, v_shaped)
    int lt = 0, gprev = -1;
    int e_cb = 0;                    // TV: chunk accumulator this warp reads next (round robin) and its phase
    uint32_t e_cph = 0;
    for (int it = it_lo; it < it_hi; ++it, ++lt) {
      const int g = it / p.n_vt, vt = it % p.n_vt;
      const int b0 = g * LG, nb = min(LG, p.B - b0);
      if (et == 0 && lt == p.dbg_lt + 1) stamp(10);
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
          const int n = nb * rper;
          if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {   // 128-bit loads, all in flight before the first store
            const int n4 = n >> 2;
#pragma unroll 4
            for (int i = et; i < n4; i += 32 * kLbsEpi)
              reinterpret_cast<float4 *>(rs)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
            if (et < (n & 3)) rs[n4 * 4 + et] = __ldg(src + n4 * 4 + et);
          } else {
            for (int i = et; i < n; i += 32 * kLbsEpi) rs[i] = __ldg(src + i);
          }
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
        // TV: identity joints share the transform of their nearest rotated ancestor, so the levels below the last
        // rotated joint (the three finger levels of SHAPY's 22-rotation pose) are not walked at all
        const int n_lv = TV ? s_nlv_rot : m.n_levels;
        for (int lv = 0; lv < n_lv; ++lv) {
          const int off = s_level_off[lv], end = s_level_off[lv + 1];
#pragma unroll
          for (int k = 0; k < kLbsPairs; ++k) {
            const int id = et + k * 32 * kLbsEpi;
            const int bl = id % LG, slot = id / LG;
            if (slot < off || slot >= end || bl >= nb) continue;
            const int j = s_level_joints[slot];
            const float *R = Rp[k], *Jc = Jp[k];
            // fp32 A_j element (body, joint, e): !TV [body][joint][12] (the epilogue reads rows of it as float4);
            // TV [joint][12][body]: consecutive lanes = consecutive bodies are conflict-free (a [body][32][12] array
            // puts every body of a warp on the same bank: measured 3 200 instead of 900 cycles per level)
            // and the body index is XOR-ed with the joint so that the transposing read of the operand conversion below
            // (lanes = joints of one body) is conflict-free too
            const int es = TV ? LG : 1;
            auto abase = [&](int jq) { return TV ? jq * 12 * LG + (bl ^ (jq & (LG - 1))) : (bl * AJ + jq) * 12; };
            if (TV && j >= p.n_rot) continue;   // identity joint: see below
            float *Ao = Aj + abase(j);
            float an[12];
            if (lv == 0) {
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                an[r * 4 + 0] = R[r * 3]; an[r * 4 + 1] = R[r * 3 + 1]; an[r * 4 + 2] = R[r * 3 + 2];
                an[r * 4 + 3] = Jc[r] - (R[r * 3] * Jc[0] + R[r * 3 + 1] * Jc[1] + R[r * 3 + 2] * Jc[2]);
              }
            } else {
              const int pj = TV ? s_anc[s_parents[j]] : s_parents[j];
              const float *Ap = Aj + abase(pj);
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const float g0 = Ap[(r * 4) * es], g1 = Ap[(r * 4 + 1) * es], g2 = Ap[(r * 4 + 2) * es], gt = Ap[(r * 4 + 3) * es];
                const float n0 = g0 * R[0] + g1 * R[3] + g2 * R[6];
                const float n1 = g0 * R[1] + g1 * R[4] + g2 * R[7];
                const float n2 = g0 * R[2] + g1 * R[5] + g2 * R[8];
                an[r * 4 + 0] = n0; an[r * 4 + 1] = n1; an[r * 4 + 2] = n2;
                an[r * 4 + 3] = (g0 * Jc[0] + g1 * Jc[1] + g2 * Jc[2]) + gt - (n0 * Jc[0] + n1 * Jc[1] + n2 * Jc[2]);
              }
            }
#pragma unroll
            for (int e = 0; e < 12; ++e) Ao[e * es] = an[e];
            // posed joint = G_j.t = A_j.t + A_j.R J_j: written once per group (by the CTA that owns vertex tile 0)
            if (vt == 0 && p.joints) {
              float *jo = p.joints + ((size_t)(b0 + bl) * m.K + j) * 3;
#pragma unroll
              for (int r = 0; r < 3; ++r)
                jo[r] = an[r * 4 + 3] + (an[r * 4] * Jc[0] + an[r * 4 + 1] * Jc[1] + an[r * 4 + 2] * Jc[2]);
            }
          }
          epi_bar();
        }
        if (TV && vt == 0 && p.joints) {
          // posed identity joints (the CTA that owns vertex tile 0 writes the group's joints): A_j == A of the nearest
          // rotated ancestor, so G_j.t = A.t + A.R J_j
#pragma unroll
          for (int k = 0; k < kLbsPairs; ++k) {
            const int id = et + k * 32 * kLbsEpi;
            const int bl = id % LG, slot = id / LG;
            if (slot >= J || bl >= nb) continue;
            const int j = s_level_joints[slot];
            if (j < p.n_rot) continue;
            const int ja = s_anc[j];
            const float *Aa = Aj + ja * 12 * LG + (bl ^ (ja & (LG - 1)));
            const float *Jc = Jp[k];
            float *jo = p.joints + ((size_t)(b0 + bl) * m.K + j) * 3;
#pragma unroll
            for (int r = 0; r < 3; ++r)
              jo[r] = Aa[(r * 4 + 3) * LG] + (Aa[(r * 4) * LG] * Jc[0] + Aa[(r * 4 + 1) * LG] * Jc[1] + Aa[(r * 4 + 2) * LG] * Jc[2]);
          }
        }
        if (TV) {
          // ---- A_j of the rotated joints -> B operand of the blend GEMM: rows (part, component) of chunk c = body % 8,
          // columns = joints, fp16 hi / lo (x 64), K-major SWIZZLE_64B.  The operand overwrites the fp32 array, so
          // every thread first takes its (body, joint) pairs into registers.
          // Item i = operand row (chunk c, part, component e) x joint pair: 16 lanes write one 64-byte row as half2.
          constexpr int kItems = (LG * 12 * kTvJ / 2) / (32 * kLbsEpi);   // 12 per thread
          float av[kItems][2];
#pragma unroll
          for (int k = 0; k < kItems; ++k) {
            const int i = et + k * 32 * kLbsEpi;
            const int row = i >> 4, jj = (i & 15) * 2;
            const int c = row / (4 * 12), prt = (row / 12) & 3, e = row % 12;
            const int bl = prt * (LG / kTvChunk) + c;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const bool on = bl < nb && jj + h < p.n_rot;
              av[k][h] = on ? Aj[((jj + h) * 12 + e) * LG + (bl ^ (jj + h))] * kTvAScale : 0.f;
            }
          }
          epi_bar();
#pragma unroll
          for (int k = 0; k < kItems; ++k) {
            const int i = et + k * 32 * kLbsEpi;
            const int row = i >> 4, jj = (i & 15) * 2;
            const int c = row / (4 * 12), r48 = row % (4 * 12);
            uint8_t *blk = smem_gen + aj_off + (size_t)c * (2 * kTvBlkBytes);
            const __half h0 = __float2half_rn(av[k][0]), h1 = __float2half_rn(av[k][1]);
            const __half l0 = __float2half_rn(av[k][0] - __half2float(h0)), l1 = __float2half_rn(av[k][1] - __half2float(h1));
            const uint32_t o = sw64_off(r48, jj);
            *reinterpret_cast<__half2 *>(blk + o) = __halves2half2(h0, h1);
            *reinterpret_cast<__half2 *>(blk + kTvBlkBytes + o) = __halves2half2(l0, l1);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(aop_full);
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
        const bool on = !TV && vok && sl < m.ell_w_n;
        ew[sl] = on ? __ldg(m.ell_w + (size_t)sl * m.V + v) : 0.f;
        ej[sl] = on ? __ldg(m.ell_idx + (size_t)sl * m.V + v) : 0;
      }
      const int buf = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      if (et == 0 && (unsigned)(lt - p.dbg_lt) < 2u) stamp(4 + 3 * (lt - p.dbg_lt));
      mbar_wait(acc_full0 + 8u * buf, aph);
      if (et == 0 && (unsigned)(lt - p.dbg_lt) < 2u) stamp(5 + 3 * (lt - p.dbg_lt));
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16) + part * 8;
      const int rows_here = min(32, m.V - v0w);      // > 0 for every tile (V > (n_vt - 1) * 128 + 96 is NOT assumed)
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        // pose offsets of four of this warp's bodies
        float pp[4][3];
        {
          uint32_t d0[3][4], d1[3][4];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            tmem_ld4(lane_addr + c * 64 + half * 4, d0[c]);
            tmem_ld4(lane_addr + c * 64 + LG + half * 4, d1[c]);
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c)
              pp[i][c] = (__uint_as_float(d0[c][i]) + __uint_as_float(d1[c][i]) * kLoInvL) * (1.0f / kPoseScale);
        }
        if (half == 1) {   // last TMEM read of this item: hand the accumulator buffer back to the MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty0 + 8u * buf);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int bl = part * 8 + half * 4 + i;
          uint32_t tv[12];
          if (TV) {
            // blend-GEMM chunk c = half * 4 + i holds body `bl` of every part: this warp's 12 columns start at part * 12.
            // The handshake runs for every chunk, also for bodies past the end of the batch.
            mbar_wait(acc2_full0 + 8u * e_cb, e_cph);
            tc_fence_after();
            const uint32_t ta = tmem_base + (e_cb ? 448u : 192u) + ((uint32_t)(q * 32) << 16) + part * 12;
            uint32_t t0[4], t1[4], t2[4];
            tmem_ld4(ta, t0); tmem_ld4(ta + 4, t1); tmem_ld4(ta + 8, t2);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 4; ++e) { tv[e] = t0[e]; tv[4 + e] = t1[e]; tv[8 + e] = t2[e]; }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc2_empty0 + 8u * e_cb);
            if (++e_cb == kTvBufs) { e_cb = 0; e_cph ^= 1; }
            if (bl >= nb) continue;                              // warp-uniform
          } else {
            if (bl >= nb) break;                                 // warp-uniform
          }
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
            vp[c] = vs[c] + pp[i][c];
          }
          float o3[3] = {0.f, 0.f, 0.f};
          if (TV) {
            constexpr float kInv = 1.0f / (kTvWScale * kTvAScale);
#pragma unroll
            for (int r = 0; r < 3; ++r)
              o3[r] = (__uint_as_float(tv[r * 4]) * vp[0] + __uint_as_float(tv[r * 4 + 1]) * vp[1] +
                       __uint_as_float(tv[r * 4 + 2]) * vp[2] + __uint_as_float(tv[r * 4 + 3])) * kInv;
          } else if (vok) {
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
          // staged through shared memory and written as three fully coalesced 128-byte warp stores per array (a
          // version with aligned 128-bit stores needed a scalar path for the row's head / tail, since rows start at
          // any 4-byte offset: more instructions than it saved)
          float *gv = p.vertices + ((size_t)b * m.V + v0w) * 3;
          const int nfl = rows_here * 3;
          __syncwarp();
          if (vok) {
            myst[3 * lane] = o3[0]; myst[3 * lane + 1] = o3[1]; myst[3 * lane + 2] = o3[2];
            myst[100 + 3 * lane] = vs[0]; myst[100 + 3 * lane + 1] = vs[1]; myst[100 + 3 * lane + 2] = vs[2];
          }
          __syncwarp();
          if (p.dbg_flags & 1) continue;
          if (nfl == 96) {
#pragma unroll
            for (int k = 0; k < 3; ++k) gv[k * 32 + lane] = myst[k * 32 + lane];
            if (p.v_shaped) {
              float *gs = p.v_shaped + ((size_t)b * m.V + v0w) * 3;
#pragma unroll
              for (int k = 0; k < 3; ++k) gs[k * 32 + lane] = myst[100 + k * 32 + lane];
            }
          } else {
            for (int f = lane; f < nfl; f += 32) {
              gv[f] = myst[f];
              if (p.v_shaped) p.v_shaped[((size_t)b * m.V + v0w) * 3 + f] = myst[100 + f];
            }
          }
        }
      }
      if (et == 0 && (unsigned)(lt - p.dbg_lt) < 2u) stamp(6 + 3 * (lt - p.dbg_lt));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(31);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

size_t lbs_smem_bytes(int J, bool tv) {
  const size_t aj = tv ? (size_t)kTvOpBytes + 2 * kTvWBytes : (size_t)LG * J * 12 * 4;
  return 1024 + LStages * kStageBytes + kLbsMaxKB * kCoefBlkBytes + aj + LG * 12 * 4 + kLbsEpi * 2 * 100 * 4 +
         (size_t)J * 3 * 12 * 4 + 16 + 256;   // + 0.5 KB of static tables (s_parents ...)
}

// Skinning weights folded onto the rotated joints and packed for the blend GEMM: [Vpad][hi 32 | lo 32] fp16 (x 1024),
// one 2-D tensor map with 128-row boxes.  Built once per n_rot.
static const shapy_smplx::WTiles *get_wtiles(const shapy_smplx *mm, int n_rot) {
  std::lock_guard<std::mutex> lk(mm->wmu);
  auto it = mm->wtiles.find(n_rot);
  if (it != mm->wtiles.end()) return it->second.ok ? &it->second : nullptr;
  shapy_smplx::WTiles &w = mm->wtiles[n_rot];
  const SmplxDev &d = mm->d;
  if (n_rot > kTvJ || mm->h_lbs_weights.empty() || !get_encode()) return nullptr;
  std::vector<int> anc(d.J);
  for (int j = 0; j < d.J; ++j) anc[j] = j < n_rot ? j : anc[std::max(mm->h_parents[j], 0)];
  std::vector<__half> pk((size_t)d.Vpad * 64, __float2half_rn(0.f));
  std::vector<float> row(kTvJ);
  for (int v = 0; v < d.V; ++v) {
    std::fill(row.begin(), row.end(), 0.f);
    for (int j = 0; j < d.J; ++j) row[anc[j]] += mm->h_lbs_weights[(size_t)v * d.J + j];
    for (int k = 0; k < kTvJ; ++k) {
      const float x = row[k] * kTvWScale;
      const __half h = __float2half_rn(x);
      pk[(size_t)v * 64 + k] = h;
      pk[(size_t)v * 64 + 32 + k] = __float2half_rn(x - __half2float(h));
    }
  }
  if (cudaMalloc((void **)&w.dev, pk.size() * sizeof(__half)) != cudaSuccess) { cudaGetLastError(); w.dev = nullptr; return nullptr; }
  if (cudaMemcpy(w.dev, pk.data(), pk.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  cuuint64_t dims[2] = {64, (cuuint64_t)d.Vpad};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, 128};
  w.ok = encode(&w.map, w.dev, 2, dims, strides, box, 64);
  return w.ok ? &w : nullptr;
}

int launch_lbs_fused(const shapy_smplx *mm, const float *betas, const float *rot, int n_rot, int B, float *vertices,
                     float *v_shaped, float *joints, int *lut, cudaStream_t st) {
  const SmplxDev &d = mm->d;
  const int Kp = (n_rot - 1) * 9, nkb = std::max(1, ceil_div(Kp, LKB));
  static const bool off = []() { const char *e = getenv("SHAPY_LBS_FUSED"); return e && e[0] == '0'; }();
  static const bool tv_off = []() { const char *e = getenv("SHAPY_LBS_TV"); return e && e[0] == '0'; }();
  if (off || !mm->fused_ok || !vertices || d.NB != kLbsNB || nkb > kLbsMaxKB) return SHAPY_ERR_UNSUPPORTED;
  const shapy_smplx::WTiles *wt = tv_off ? nullptr : get_wtiles(mm, n_rot);
  const bool tv = wt != nullptr;
  if (lbs_smem_bytes(d.J, tv) > 226 * 1024) return SHAPY_ERR_UNSUPPORTED;
  LbsParams p;
  memset(&p, 0, sizeof(p));
  p.basis = mm->basis_map;
  { const char *e = getenv("SHAPY_LBS_DBGFLAGS"); p.dbg_flags = e ? atoi(e) : 0; }
  if (tv) p.wmap = wt->map;
  p.m = d;
  p.betas = betas; p.rot = rot; p.n_rot = n_rot; p.B = B; p.Kp = Kp; p.nkb = nkb;
  p.n_vt = d.Vpad / LV;
  p.n_items = ceil_div(B, LG) * p.n_vt;
  p.vertices = vertices; p.v_shaped = v_shaped; p.joints = joints; p.lut = lut;
  p.idesc64 = (1u << 4) | ((uint32_t)((2 * LG) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.idesc32 = (1u << 4) | ((uint32_t)(LG >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.idesc48 = (1u << 4) | ((uint32_t)(kTvN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static std::atomic<unsigned long long> attr_done{0}, attr_done_tv{0};
  if (tv) SHAPY_CUDA_TRY(set_max_dynamic_smem(smplx_lbs_kernel<true>, 226 * 1024, attr_done_tv));   // + 1.3 KB static <= 227 KB
  else SHAPY_CUDA_TRY(set_max_dynamic_smem(smplx_lbs_kernel<false>, 226 * 1024, attr_done));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(p.n_items, sms);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kLbsThreads);
  cfg.dynamicSmemBytes = lbs_smem_bytes(d.J, tv);
  cfg.stream = st;
  static const bool dbg = getenv("SHAPY_LBS_DEBUG") != nullptr;
  if (dbg) {   // synchronous: per-CTA cycle stamps of the roles (prologue / MMA / epilogue phases)
    long long *d = nullptr;
    cudaMalloc(&d, (size_t)grid * 32 * 8);
    cudaMemset(d, 0, (size_t)grid * 32 * 8);
    p.dbg = d;
    p.dbg_lt = p.n_items / grid > 10 ? 8 : 0;
    if (tv) SHAPY_CUDA_TRY(cudaLaunchKernelEx(&cfg, smplx_lbs_kernel<true>, p));
    else SHAPY_CUDA_TRY(cudaLaunchKernelEx(&cfg, smplx_lbs_kernel<false>, p));
    cudaStreamSynchronize(st);
    std::vector<long long> h((size_t)grid * 32);
    cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    const char *names[32] = {"", "coef_ready", "pairs_loaded", "chain_done", "it0_consts", "it0_acc_full", "it0_epi_done",
                             "it1_consts", "it1_acc_full", "it1_epi_done", "it1_top", "mma_it0_chunks_done", "mma_it0_chunk0", "mma_it0_g1_open", "mma_it1_g1_open",
                             "regs_granted", "mma_coef_wait", "mma_it0_issued", "mma_it1_issued", "c0_ready", "c1_ready", "c2_ready",
                             "c3_ready", "c4_ready", "c5_ready", "c6_ready", "c7_ready", "c6_wait", "c7_wait", "", "", "exit"};
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
  if (tv) SHAPY_CUDA_TRY(cudaLaunchKernelEx(&cfg, smplx_lbs_kernel<true>, p));
  else SHAPY_CUDA_TRY(cudaLaunchKernelEx(&cfg, smplx_lbs_kernel<false>, p));
  count_launch();
  return SHAPY_OK;
}

}  // namespace shapy
