// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (hand-written PTX, no CUTLASS).
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel shifted by tap, cin] * Wt[tap][cout][cin]
//
// GEMM view: M = 128 output pixels, N = cout tile (<= 128), K = taps * cin walked in 16-channel k-steps.
// Two kernels share the PTX helpers, the epilogue and the weight layout:
//
//  conv_umma_kernel ("per-tap"): 1x1 convs, 64-channel-block 3x3 convs, anything the halo variant declines.
//   * A: one 4-D TMA box per (tap, k-block) straight out of the NHWC activation plane, the box origin shifted by
//     the tap; out-of-image rows/columns are ZERO-FILLED by the TMA unit = the padding (no im2col, no halo
//     storage).  Stride-2 layers use four parity views of the same plane so every tap is again a dense box.
//     Channel counts that are not a multiple of 64 still use 64-channel boxes: TMA zero-fills the channels past cin.
//   * B: 3-D TMA box of the pre-packed [tap][cout][cin] fp16 weights (BN scale folded in).
//   * warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2.. = kEpiWarps epilogue warps; smem ring
//     of `stages` slots with full/empty mbarriers; two TMEM accumulator buffers so tile t's epilogue overlaps
//     tile t+1's main loop; persistent, one CTA per SM.
//
//  conv_halo_kernel ("halo-resident"): 3x3 stride-1 and stride-2 convs with <= 32-channel k-blocks (C = 48 / 96:
//   45 % of the network's time).  See the comment above the kernel.
//
// Common: operands land in the canonical K-major SWIZZLE_{32,64,128}B layout the UMMA shared-memory descriptor names,
// so no thread ever touches operand data; one elected lane issues tcgen05.mma (M=128, N=cout tile, K=16) into fp32
// TMEM accumulators.  Split-fp16 parity mode keeps activations and weights as fp16 hi + fp16 lo (x = hi + lo/2048) and
// issues A_hi.[B_hi|B_lo] as ONE N = 2 NT MMA into [D0|D1] plus A_lo.B_hi into D1 (out = D0 + D1/2048).
// Epilogue: tcgen05.ld -> +bias (+residual) -> ReLU -> split to fp16 hi/lo -> 256-bit global stores.
//
// Measured facts this design rests on (tools/umma_bench.cu, profiles/r01_umma_microbench.txt):
//   * one tcgen05.mma M=128 K=16 costs max(32 + N/4, N/2) cycles: operands are read from shared memory at 128 B/cycle
//     (4 KB of A + 32 N bytes of B) unless the tensor pipe (N/2) is slower; N >= 128 reaches the tensor peak, N = 48
//     is shared-memory bound at 54 %.  M = 64 costs the same as M = 128.  A from TMEM would remove the 32 cycles.
//   * accumulating into the same TMEM tile back to back costs nothing extra; swizzle mode of A / B and row-shifted
//     start addresses cost nothing; the swizzle is a function of the absolute shared-memory address (row shifts
//     need no base-offset field).
//   * a single issuing lane sustains ~70 cycles per MMA once its descriptor arithmetic is included: the loops below
//     keep descriptors as 32-bit low words with hoisted high words, unroll m-tiles at compile time and use two
//     issuer warps in the halo kernel.
//   * 16-byte epilogue stores are bound by the request rate (one half-filled sector each): 256-bit STG/LDG.
// Replaces cuDNN conv + BatchNorm + ReLU + residual add (4 launches, 4 HBM round trips) of
// regressor/human_shape/models/backbone/hrnet.py and torchvision BasicBlock / Bottleneck.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "conv.cuh"
#include "umma.cuh"

namespace shapy {

// ------------------------------------------------------------------------------------ parameters
struct alignas(64) UmmaParams {
  CUtensorMap a_hi[4], a_lo[4];  // [parity] (stride 2) or [0] (stride 1)
  CUtensorMap b_hi, b_lo;
  int N, Ho, Wo;                 // output extent
  int TW, TH, TN;                // M-tile box (TW * TH * TN <= 128 rows)
  int tiles_w, tiles_h, tiles_n;
  int cin, cout, NT, ksize, stride;
  int kpt;                       // k-blocks per tap = cin / KCH
  int G;                         // k-blocks per pipeline stage
  int stages;
  int relu;
  int pdl_late;                  // 1: release the programmatic dependents when this CTA starts its LAST tile, not at entry
  int vec32;                     // out / residual rows are 32-byte aligned: 256-bit epilogue accesses
  uint32_t idesc, idesc2;        // idesc: N = NT;  idesc2: N = 2 NT ([B_hi | B_lo] in one MMA, split mode)
  uint32_t tmem_cols;
  uint32_t a_bytes, b_bytes, b_stride;  // per k-block: TMA bytes of A / B, smem pitch of a B block
  const float *bias;
  __half *out_hi, *out_lo;
  int out_ctot, out_coff;
  const __half *res_hi, *res_lo;
  int res_ctot, res_coff;
};

// Persistent, warp-specialised kernel: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
//   warp 0      TMA producer  (smem ring runs ahead across tile boundaries: no pipeline drain per tile)
//   warp 1      TMEM allocator + MMA issuer (alternates between two TMEM accumulator buffers)
//   warps 2..9  epilogue: drains accumulator buffer t & 1 while the MMA warp fills the other one
#ifndef SHAPY_EPI_WARPS
#define SHAPY_EPI_WARPS 16
#endif
constexpr int kEpiWarps = SHAPY_EPI_WARPS;   // multiple of 4: warp w may only read TMEM lanes 32 (w % 4) .. +31
constexpr int kEpiParts = kEpiWarps / 4;     // warps sharing a lane quarter split the 16-column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps;

template <int KCH, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BLK = 128 * KCH * 2;
  const uint32_t B_BLK = p.b_bytes;          // lo rows follow the hi rows directly: [B_hi | B_lo] is one N = 2 NT operand
  const uint32_t kblk_bytes = A_BLK * (SPLIT ? 2 : 1) + p.b_stride;   // b_stride = padded size of the hi(+lo) pair
  const uint32_t stage_bytes = kblk_bytes * p.G;
  const uint32_t bar_base = smem_base + stage_bytes * p.stages;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t acc_full0 = bar_base + 16u * p.stages;   // [2]
  const uint32_t acc_empty0 = acc_full0 + 16u;            // [2]
  const uint32_t tmem_slot = acc_empty0 + 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.cout / p.NT;
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * n_tiles;
  const int taps = p.ksize * p.ksize;
  const int iters = taps * p.kpt / p.G;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full0 + 8u * i, 1); mbar_init(acc_empty0 + 8u * i, kEpiWarps); }
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
  // Programmatic dependent launch: the next kernel of this stream may start its prologue once every CTA here has
  // released it.  Released at entry, its CTAs occupy SMs (which another lane's kernel could use) for this kernel's whole
  // duration; released when a CTA starts its last tile (pdl_late), they only cover this kernel's tail.
  if (!p.pdl_late) pdl_launch_dependents();
  pdl_wait();                // everything above touched no global memory; inputs of this kernel are now complete

  auto tile_coords = [&](int id, int &w0, int &h0, int &n0, int &c_out0) {
    c_out0 = (id % n_tiles) * p.NT; id /= n_tiles;
    w0 = (id % p.tiles_w) * p.TW; id /= p.tiles_w;
    h0 = (id % p.tiles_h) * p.TH; id /= p.tiles_h;
    n0 = id * p.TN;
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    // (the whole warp runs the loops so that control flow stays warp-uniform; one elected lane issues)
    {
      int s = 0;
      uint32_t ph = 0;
      const int pad = p.ksize / 2;
      const int per_tap = p.kpt / p.G;
      const uint32_t tx = (p.a_bytes + p.b_bytes) * (SPLIT ? 2u : 1u) * p.G;
      if (elect_one()) {
        for (int id = blockIdx.x; id < total_tiles; id += gridDim.x) {
          int w0, h0, n0, c_out0;
          tile_coords(id, w0, h0, n0, c_out0);
          if (p.pdl_late && id + (int)gridDim.x >= total_tiles) pdl_launch_dependents();   // last tile of this CTA
#pragma unroll 1
          for (int it = 0; it < iters; ++it) {
            const int tap = it / per_tap, cb0 = (it % per_tap) * p.G;
            const int ky = tap / p.ksize, kx = tap % p.ksize;
            int mi = 0, x, y;
            if (p.stride == 1) {
              x = w0 + kx - pad; y = h0 + ky - pad;
            } else {  // input row 2*oh + ky - 1: ky=0 -> odd rows, index oh-1; ky=1 -> even rows, oh; ky=2 -> odd rows, oh
              mi = (ky != 1) * 2 + (kx != 1);
              x = w0 + (kx == 0 ? -1 : 0); y = h0 + (ky == 0 ? -1 : 0);
            }
            mbar_wait(empty_bar(s), ph ^ 1);
            mbar_expect_tx(full_bar(s), tx);
            const uint32_t sbase = smem_base + stage_bytes * s;
            for (int g = 0; g < p.G; ++g) {
              const uint32_t kb = sbase + kblk_bytes * g;
              const int c0 = (cb0 + g) * KCH;
              tma_load_4d(kb, &p.a_hi[mi], full_bar(s), c0, x, y, n0);
              if (SPLIT) tma_load_4d(kb + A_BLK, &p.a_lo[mi], full_bar(s), c0, x, y, n0);
              const uint32_t bb = kb + A_BLK * (SPLIT ? 2 : 1);
              tma_load_3d(bb, &p.b_hi, full_bar(s), c0, c_out0, tap);
              if (SPLIT) tma_load_3d(bb + B_BLK, &p.b_lo, full_bar(s), c0, c_out0, tap);  // directly after the hi rows
            }
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    {
      int s = 0, lt = 0;
      uint32_t ph = 0;
      for (int id = blockIdx.x; id < total_tiles; id += gridDim.x, ++lt) {
        const int buf = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        mbar_wait(acc_empty0 + 8u * buf, aph ^ 1);   // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t d0 = tmem_base + buf * 256u, d1 = d0 + p.NT;
        if (elect_one()) {
          int s_l = s;
          uint32_t ph_l = ph;
          const uint32_t dhi = desc_hi_swz<KCH>();
          const uint32_t idesc = p.idesc, idesc2 = p.idesc2;
          const uint32_t kblk16 = kblk_bytes >> 4, stage16 = stage_bytes >> 4;
          const uint32_t base16 = desc_lo_swz(smem_base);
          const int G = p.G;
          uint32_t first = 0;
#pragma unroll 1
          for (int it = 0; it < iters; ++it) {
            mbar_wait(full_bar(s_l), ph_l);
            tc_fence_after();
            uint32_t a = base16 + stage16 * s_l;      // low descriptor word of this k-block's A_hi
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
              const uint32_t al = a + (A_BLK >> 4), b = a + ((A_BLK * (SPLIT ? 2 : 1)) >> 4);
#pragma unroll
              for (int ks = 0; ks < KCH / 16; ++ks) {   // +32 bytes along K = +2 descriptor units
                if (SPLIT) {
                  // [D0 | D1] (+)= A_hi . [B_hi | B_lo]^T  (one N = 2 NT MMA), then D1 += A_lo . B_hi^T
                  umma_f16_lh(d0, a + 2 * ks, dhi, b + 2 * ks, dhi, idesc2, ks ? 1u : first);
                  umma_f16_lh(d1, al + 2 * ks, dhi, b + 2 * ks, dhi, idesc, 1u);
                } else {
                  umma_f16_lh(d0, a + 2 * ks, dhi, b + 2 * ks, dhi, idesc, ks ? 1u : first);
                }
              }
              first = 1;
              a += kblk16;
            }
            umma_commit(empty_bar(s_l));
            if (++s_l == p.stages) { s_l = 0; ph_l ^= 1; }
          }
          umma_commit(acc_full0 + 8u * buf);
        }
        __syncwarp();
        for (int it = 0; it < iters; ++it) { if (++s == p.stages) { s = 0; ph ^= 1; } }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;        // which of every kEpiParts 16-column chunks this warp handles
    const int row = q * 32 + lane;           // tile row == TMEM lane
    const int dw = row % p.TW, dh = (row / p.TW) % p.TH, dn = row / (p.TW * p.TH);
    int lt = 0;
    for (int id = blockIdx.x; id < total_tiles; id += gridDim.x, ++lt) {
      int w0, h0, n0, c_out0;
      tile_coords(id, w0, h0, n0, c_out0);
      const int ow = w0 + dw, oh = h0 + dh, on = n0 + dn;
      const bool ok = dn < p.TN && ow < p.Wo && oh < p.Ho && on < p.N;
      const size_t pix = ((size_t)on * p.Ho + oh) * p.Wo + ow;
      const int buf = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      mbar_wait(acc_full0 + 8u * buf, aph);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16);
      for (int c = part * 16; c < p.NT; c += 16 * kEpiParts) {
        const int co = c_out0 + c;
        // residual rows are fetched before the TMEM load completes so the two latencies overlap
        uint4 rh0 = make_uint4(0, 0, 0, 0), rh1 = rh0, rl0 = rh0, rl1 = rh0;
        const bool has_res = ok && p.res_hi != nullptr;
        if (has_res) {
          const size_t ro = pix * p.res_ctot + p.res_coff + co;
          const uint4 *rh = reinterpret_cast<const uint4 *>(p.res_hi + ro);
          if (p.vec32) ldg256(rh, rh0, rh1); else { rh0 = __ldg(rh); rh1 = __ldg(rh + 1); }
          if (SPLIT && p.res_lo) {
            const uint4 *rl = reinterpret_cast<const uint4 *>(p.res_lo + ro);
            if (p.vec32) ldg256(rl, rl0, rl1); else { rl0 = __ldg(rl); rl1 = __ldg(rl + 1); }
          }
        }
        uint32_t v0[16], v1[16];
        tmem_ld16(lane_addr + c, v0);
        if (SPLIT) tmem_ld16(lane_addr + p.NT + c, v1);
        tmem_ld_wait();
        if (!ok) continue;
        float r[16];
        {
          const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + co);   // co % 16 == 0: 64-byte aligned
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 bq = __ldg(b4 + q4);
            r[4 * q4] = bq.x; r[4 * q4 + 1] = bq.y; r[4 * q4 + 2] = bq.z; r[4 * q4 + 3] = bq.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = __uint_as_float(v0[j]);
          if (SPLIT) x += __uint_as_float(v1[j]) * kLoInv;
          r[j] += x;
        }
        if (has_res) {
          const __half *hh0 = reinterpret_cast<const __half *>(&rh0), *hh1 = reinterpret_cast<const __half *>(&rh1);
          const __half *ll0 = reinterpret_cast<const __half *>(&rl0), *ll1 = reinterpret_cast<const __half *>(&rl1);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            r[j] += __half2float(hh0[j]) + __half2float(ll0[j]) * kLoInv;
            r[8 + j] += __half2float(hh1[j]) + __half2float(ll1[j]) * kLoInv;
          }
        }
        __align__(16) __half hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split_store(p.relu ? fmaxf(r[j], 0.f) : r[j], hi[j], lo[j]);
        const size_t oo = pix * p.out_ctot + p.out_coff + co;
        const uint4 *hv = reinterpret_cast<const uint4 *>(hi), *lv = reinterpret_cast<const uint4 *>(lo);
        uint4 *oh4 = reinterpret_cast<uint4 *>(p.out_hi + oo);
        if (p.vec32) stg256(oh4, hv[0], hv[1]); else { oh4[0] = hv[0]; oh4[1] = hv[1]; }
        if (SPLIT && p.out_lo) {
          uint4 *ol4 = reinterpret_cast<uint4 *>(p.out_lo + oo);
          if (p.vec32) stg256(ol4, lv[0], lv[1]); else { ol4[0] = lv[0]; ol4[1] = lv[1]; }
        }
      }
      // this warp is done reading the buffer: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty0 + 8u * buf) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =================================================================================================
// 3x3 convolutions, "halo-resident" variant.
//
// The per-tap variant re-reads every input pixel 9 times from L2 (one shifted box per tap; ncu: L2 52 %, tensor pipe
// 15 %) and, for stride 2, is bound by the TMA unit's sector requests (each gathered pixel is its own 32..96-byte
// request, 2.25x per pixel).  Here the input of a work item (a band of R output rows of one image, or TN whole small
// images) is loaded ONCE per channel group by TMA, with its halo (out-of-image rows / columns are zero-filled = the
// padding), as K-major swizzled rows  slice[position][KCH channels],  position = (n * Hb + row) * Wp + col  on the
// padded grid.  The A operand of tap (ky, kx) for output positions o .. o+127 is the same slice read from position
// o + ky * Wp + kx: all 9 taps x MT m-tiles are addressed by moving the start address of the UMMA descriptor by whole
// rows (legal for any row count: the hardware swizzle is a function of the absolute address).  Stride 2 keeps four
// parity planes (row parity x column parity) of the input per slice; tap (ky, kx) reads plane (ky != 1, kx != 1) at
// position o (+ Wp for ky == 2) (+ 1 for kx == 2).  Outputs are computed for the padding columns too (and discarded
// by the epilogue); in exchange activations cross L2 ~1.3x instead of 9x and the weights of one (tap, channel group)
// are shared by up to 4 m-tiles held in TMEM at once (C = 48: all 27 weight blocks stay resident in shared memory).
struct alignas(64) HaloParams {
  CUtensorMap a_hi[4], a_lo[4];  // stride 1: [0] = (C, W, H, N); stride 2: the four parity views [row parity * 2 + column
                                 // parity] of the input, each (C, W/2, H/2, N); box (KCH, Wp, Hb, TN), swizzled rows
  CUtensorMap b_hi, b_lo;  // (cin, cout, 9) swizzled, box (KCH, NT, 1)
  int N, H, W;
  int Wp, Hb, TN, R, bands, n_super, MT;
  int cin, cout, NT, n_tiles, ncg, bstages, acc_bufs, b_resident, n_iss;
  int stride;       // 1 or 2 (3x3 both); for stride 2, N / H / W are the OUTPUT extent and the A slice holds 4 parity planes
  uint32_t plane_bytes;  // size of one plane of an A slice part (stride 1: == part_bytes)
  int row_sched;    // 1: each CTA owns a contiguous range of the N*H output rows, walked in chunks of <= R rows
  int a_baseoff;    // 1: put (start >> 7) & 7 into the descriptor's base-offset field for row-shifted starts
  uint32_t part_bytes;   // offset of the lo part inside an A slice
  uint32_t a_slice_bytes, a_tx, b_bytes, b_stride;
  int n_slices;          // A slices in flight (2 or 3): the load of a slice takes about as long as its MMAs, so two
                         // leave the issuers waiting ~300 cycles per slice (r02_phases.txt: wait_a); three when smem allows
  uint32_t idesc, idesc2;
  int relu;
  int pdl_late;          // 1: release the programmatic dependents when this CTA starts its LAST item, not at entry
  int vec32;             // out / residual rows are 32-byte aligned: 256-bit epilogue accesses
  const float *bias;
  __half *out_hi, *out_lo;
  int out_ctot, out_coff;
  const __half *res_hi, *res_lo;
  int res_ctot, res_coff;
  unsigned long long *dbg;       // optional [gridDim.x][16] phase cycle counters (SHAPY_CONV_PHASES=1)
  int dbg_flags;                 // phases mode only: 1 skip the A fills, 2 skip the epilogue work, 4 skip the MMAs
};

// Halo kernel warp roles: warp 0 = A-slice TMA producer, warp 1 = weight TMA producer, warps 2..3 = MMA issuers
// (m-tiles dealt round-robin), then kEpiWarps epilogue warps.
constexpr int kHaloProd = 2;
constexpr int kHaloIssue = 2;   // MMA-issuing warps: m-tiles are dealt round-robin, each warp owns its accumulators
constexpr int kHaloThreads = 32 * (kHaloProd + kHaloIssue + kEpiWarps);

template <int KCH, bool SPLIT>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int PARTS = SPLIT ? 2 : 1;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;                                  // 2 slices
  const uint32_t b_base = a_base + (uint32_t)p.n_slices * p.a_slice_bytes;   // ring of bstages x PARTS blocks
  const uint32_t b_stage_bytes = p.b_stride;                          // padded size of the [B_hi | B_lo] pair
  const uint32_t bar_base = b_base + b_stage_bytes * p.bstages;
  auto b_full = [&](int s) { return bar_base + 8u * s; };
  auto b_empty = [&](int s) { return bar_base + 8u * (p.bstages + s); };
  const uint32_t a_full0 = bar_base + 16u * p.bstages;   // [3]
  const uint32_t a_empty0 = a_full0 + 24u;               // [3]
  const uint32_t acc_full0 = a_empty0 + 24u;             // [2]
  const uint32_t acc_empty0 = acc_full0 + 16u;           // [2]
  const uint32_t tmem_slot = acc_empty0 + 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = p.n_super * p.n_tiles;
  unsigned long long g_entry = 0;
  long long c_entry = 0;
  if (p.dbg && threadIdx.x == 0) { asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_entry)); c_entry = clock64(); }

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.bstages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), p.n_iss); }
    for (int i = 0; i < 3; ++i) { mbar_init(a_full0 + 8u * i, 1); mbar_init(a_empty0 + 8u * i, p.n_iss); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full0 + 8u * i, p.n_iss); mbar_init(acc_empty0 + 8u * i, kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kHaloProd) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (!p.pdl_late) pdl_launch_dependents();   // see conv_umma_kernel
  // Everything above touched no global memory.  The wait for the predecessor grid is taken per role: the weights are
  // constants, so their producer (warp 1) starts at once and the 83 KB of a resident 48-channel layer land while the
  // predecessor's last CTAs are still running; the activation producer and the epilogue (residual reads, output
  // writes) wait; the MMA issuers only touch shared memory / TMEM.
  if (p.pdl_late < 2 || warp == 0 || warp >= kHaloProd + kHaloIssue) pdl_wait();   // pdl_late == 2: per-role wait
  if (p.dbg && threadIdx.x == 0) {
    p.dbg[blockIdx.x * 16 + 8] = (unsigned long long)(clock64() - c_entry);   // prologue cycles
    p.dbg[blockIdx.x * 16 + 9] = g_entry;                                     // ns timestamp at entry
  }

  auto item_coords = [&](int id, int &n0, int &h0, int &c_out0) {
    c_out0 = (id % p.n_tiles) * p.NT; id /= p.n_tiles;
    h0 = (id % p.bands) * p.R; id /= p.bands;
    n0 = id * p.TN;
  };
  const uint32_t acc_cols = (uint32_t)p.NT * PARTS;
  // Work items.  Default: ids blockIdx.x, +gridDim.x, ... over (super-tile, cout tile).  Row scheduling (band mode with
  // one cout tile): CTA b owns output rows [b T / G, (b + 1) T / G) of the T = N * H rows and walks them in chunks of
  // at most R rows that do not cross an image boundary; 896 equal items on 148 SMs are 7 waves of 6.05, the row
  // ranges differ by one row at most.
  struct HaloItem { long long g, hi; int id, rows; };
  auto item_begin = [&]() {
    HaloItem it;
    it.id = blockIdx.x; it.rows = 0; it.g = 0; it.hi = 0;
    if (p.row_sched) {
      const long long T = (long long)p.N * p.H;
      it.g = T * blockIdx.x / gridDim.x;
      it.hi = T * (blockIdx.x + 1) / gridDim.x;
    }
    return it;
  };
  auto item_valid = [&](const HaloItem &it) { return p.row_sched ? it.g < it.hi : it.id < total_items; };
  auto item_get = [&](HaloItem &it, int &n0, int &h0, int &c_out0, int &rows) {
    if (p.row_sched) {
      n0 = (int)(it.g / p.H); h0 = (int)(it.g - (long long)n0 * p.H); c_out0 = 0;
      rows = min(min(p.R, p.H - h0), (int)(it.hi - it.g));
    } else {
      item_coords(it.id, n0, h0, c_out0);
      rows = p.R;
    }
    it.rows = rows;
  };
  auto item_next = [&](HaloItem &it) { it.g += it.rows; it.id += gridDim.x; };

  if (warp < kHaloProd) {
    // ===================================================================== producers
    // warp 0: one elected lane loads each A slice (the band of KCH channels with its halo; out-of-image rows and
    // columns are zero-filled by the TMA unit = the padding) as two boxes; warp 1: the weight blocks.
    if (warp == 0) {
      if (elect_one()) {
        int slice = 0;
        uint32_t aph = 0;
        for (HaloItem it = item_begin(); item_valid(it); item_next(it)) {
          int n0, h0, c_out0, rows;
          item_get(it, n0, h0, c_out0, rows);
          if (p.pdl_late) {   // last item of this CTA?
            HaloItem nx = it;
            item_next(nx);
            if (!item_valid(nx)) pdl_launch_dependents();
          }
          for (int cg = 0; cg < p.ncg; ++cg) {
            long long t0 = clock64();
            mbar_wait(a_empty0 + 8u * slice, aph ^ 1);
            long long t1 = clock64();
            const uint32_t sb = a_base + slice * p.a_slice_bytes, fb = a_full0 + 8u * slice;
            if (p.dbg_flags & 1) {
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb) : "memory");
            } else {
              mbar_expect_tx(fb, p.a_tx);
              if (p.stride == 1) {
                tma_load_4d(sb, &p.a_hi[0], fb, cg * KCH, -1, h0 - 1, n0);
                if (SPLIT) tma_load_4d(sb + p.part_bytes, &p.a_lo[0], fb, cg * KCH, -1, h0 - 1, n0);
              } else {
                // plane (py, px) holds input rows 2y + py, columns 2x + px: the odd planes start one row / column early
                // (row -1 / column -1 are out of range = the zero padding)
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                  const int x0 = (pl & 1) ? -1 : 0, y0 = h0 + ((pl >> 1) ? -1 : 0);
                  tma_load_4d(sb + pl * p.plane_bytes, &p.a_hi[pl], fb, cg * KCH, x0, y0, n0);
                  if (SPLIT) tma_load_4d(sb + p.part_bytes + pl * p.plane_bytes, &p.a_lo[pl], fb, cg * KCH, x0, y0, n0);
                }
              }
            }
            if (p.dbg) { long long t2 = clock64(); atomicAdd(p.dbg + blockIdx.x * 16 + 0, (unsigned long long)(t1 - t0)); atomicAdd(p.dbg + blockIdx.x * 16 + 1, (unsigned long long)(t2 - t1)); }
            if (++slice == p.n_slices) { slice = 0; aph ^= 1; }
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (elect_one()) {
        if (p.b_resident) {
          // all 9 x ncg weight blocks of the (single) cout tile stay in shared memory for the whole kernel; one barrier
          // per channel group, so the first MMAs start after a third of the weights has landed
          for (int cg = 0; cg < p.ncg; ++cg) {
            mbar_expect_tx(b_full(cg), p.b_bytes * PARTS * 9u);
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t bb = b_base + b_stage_bytes * (cg * 9 + tap);
              tma_load_3d(bb, &p.b_hi, b_full(cg), cg * KCH, 0, tap);
              if (SPLIT) tma_load_3d(bb + p.b_bytes, &p.b_lo, b_full(cg), cg * KCH, 0, tap);
            }
          }
        } else {
          int bs = 0;
          uint32_t bph = 0;
          for (HaloItem it = item_begin(); item_valid(it); item_next(it)) {
            int n0, h0, c_out0, rows;
            item_get(it, n0, h0, c_out0, rows);
            for (int cg = 0; cg < p.ncg; ++cg)
              for (int tap = 0; tap < 9; ++tap) {
                mbar_wait(b_empty(bs), bph ^ 1);
                mbar_expect_tx(b_full(bs), p.b_bytes * PARTS);
                const uint32_t bb = b_base + b_stage_bytes * bs;
                tma_load_3d(bb, &p.b_hi, b_full(bs), cg * KCH, c_out0, tap);
                if (SPLIT) tma_load_3d(bb + p.b_bytes, &p.b_lo, b_full(bs), cg * KCH, c_out0, tap);
                if (++bs == p.bstages) { bs = 0; bph ^= 1; }
              }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp < kHaloProd + kHaloIssue) {
    // ===================================================================== MMA issuers
    // Issuer w handles m-tiles w, w + n_iss, ...: a single thread sustains only one tcgen05.mma per ~70 cycles once
    // its descriptor arithmetic is added, so two warps on two schedulers feed the tensor pipe.  Every accumulator is
    // owned by one warp (deterministic summation order); each warp's tcgen05.commit covers its own MMAs.
    const int iw = warp - kHaloProd;
    if (iw < p.n_iss) {
      int bs = 0, slice = 0, lt = 0;
      uint32_t bph = 0, aph = 0;
      // Descriptor low words are advanced with 32-bit adds only; everything else is hoisted here.  (Measured with
      // SHAPY_CONV_DBGFLAGS=7: the previous per-tap 64-bit descriptor set-up cost ~250 cycles per tap and made the
      // whole kernel issue-bound.)
      const uint32_t b_hi32 = desc_hi_swz<KCH>();
      // A operand: K-major swizzled rows of KCH channels ([position][KCH], one TMA box per part); a tap or an m-tile is
      // a whole-row shift of the descriptor start address (the swizzle is a function of the absolute shared-memory
      // address, so any row offset is legal with base offset 0 -- verified on hardware against the oracle).
      const uint32_t a_hi32 = desc_hi_swz<KCH>();
      const uint32_t lo_off = p.part_bytes >> 4;
      const uint32_t ks_step = 2u;
      const uint32_t row_u = (uint32_t)(KCH / 8);                                    // descriptor units (16 B) per position
      const uint32_t a0_16 = desc_lo_swz(a_base);
      const bool baseoff = p.a_baseoff != 0;
      const uint32_t a_slice16 = p.a_slice_bytes >> 4;
      const uint32_t b0_16 = desc_lo_swz(b_base), bst16 = b_stage_bytes >> 4;
      const uint32_t idesc = p.idesc, idesc2 = p.idesc2, NT = p.NT, wp = p.Wp;
      const int MT = p.MT, ncg = p.ncg, bstages = p.bstages;
      const uint32_t m_a = (uint32_t)iw * 128u * row_u, m_d = (uint32_t)iw * acc_cols;
      const uint32_t st_a = (uint32_t)p.n_iss * 128u * row_u, st_d = (uint32_t)p.n_iss * acc_cols;
      const bool rec = p.dbg && lane == 0 && iw == 0;
      const bool no_mma = (p.dbg_flags & 4) != 0;
      // one tap: MT m-tiles x KCH/16 k-steps x (2 | 1) MMAs; at / bt = low descriptor words of the tap's A / B block.
      // MTC > 0: the m loop is unrolled at compile time (every operand becomes base + immediate); 0: runtime loop.
      int mt_rt = 0;   // m-tile count of the runtime-loop variant (MTC == 0)
      auto issue_tap = [&](auto mtc, uint32_t at, uint32_t a_hi_t, uint32_t bt, uint32_t d, uint32_t acc0) {
        constexpr int MTC = decltype(mtc)::value;
        auto one = [&](uint32_t a_m, uint32_t d_m) {
#pragma unroll
          for (int ks = 0; ks < KCH / 16; ++ks) {
            if (SPLIT) {
              umma_f16_lh(d_m, a_m + ks * ks_step, a_hi_t, bt + 2 * ks, b_hi32, idesc2, ks ? 1u : acc0);        // [D0|D1] (+)= A_hi [B_hi|B_lo]^T
              umma_f16_lh(d_m + NT, a_m + ks * ks_step + lo_off, a_hi_t, bt + 2 * ks, b_hi32, idesc, 1u);       // D1 += A_lo B_hi^T
            } else {
              umma_f16_lh(d_m, a_m + ks * ks_step, a_hi_t, bt + 2 * ks, b_hi32, idesc, ks ? 1u : acc0);
            }
          }
        };
        if (no_mma) return;
        if constexpr (MTC > 0) {
#pragma unroll
          for (int m = 0; m < MTC; ++m) one(at + m * st_a, d + m * st_d);   // 128 rows x 16 B = 128 descriptor units per m-tile
        } else {
#pragma unroll 1
          for (int m = 0; m < mt_rt; ++m) { one(at, d); at += st_a; d += st_d; }
        }
      };
      // all 9 taps of one channel group (the elected lane only)
      // per tap: start shift in descriptor units, and the high word (base offset = 128-byte line of the start, mod 8)
      // stride 1: tap (ky, kx) reads position o + ky Wp + kx.  stride 2: input (2 oy + ky - 1, 2 ox + kx - 1) lives in
      // parity plane (ky != 1, kx != 1) at position o (+ Wp for ky == 2) (+ 1 for kx == 2)
      const uint32_t plane16 = p.plane_bytes >> 4;
      const bool s2 = p.stride == 2;
      auto tap_shift = [&](int tap) {
        const uint32_t ky = (uint32_t)(tap / 3), kx = (uint32_t)(tap % 3);
        return s2 ? ((ky != 1) * 2u + (kx != 1)) * plane16 + ((ky == 2) * wp + (kx == 2)) * row_u : (ky * wp + kx) * row_u;
      };
      auto tap_hi = [&](int tap) {
        const uint32_t line = (tap_shift(tap) << 4) >> 7;
        return baseoff ? (a_hi32 | ((line & 7u) << 17)) : a_hi32;
      };
      auto issue_cg = [&](auto mtc, int cg, uint32_t as, uint32_t dbase, int bs_l, uint32_t bph_l) {
        if (p.b_resident) {
          // weights are resident: the taps go out back to back
          const uint32_t bt0 = b0_16 + (uint32_t)(cg * 9) * bst16;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
            issue_tap(mtc, as + tap_shift(tap), tap_hi(tap), bt0 + tap * bst16, dbase, (tap | cg) ? 1u : 0u);
        } else {
          // streamed weights: the elected lane also waits for each weight block (a per-tap elect + __syncwarp
          // costs ~150 cycles per tap)
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(b_full(bs_l), bph_l);
            tc_fence_after();
            issue_tap(mtc, as + tap_shift(tap), tap_hi(tap), b0_16 + bs_l * bst16, dbase, (tap | cg) ? 1u : 0u);
            umma_commit(b_empty(bs_l));
            if (++bs_l == bstages) { bs_l = 0; bph_l ^= 1; }
          }
        }
      };
      for (HaloItem it = item_begin(); item_valid(it); item_next(it), ++lt) {
        int n0_, h0_, c0_, rows;
        item_get(it, n0_, h0_, c0_, rows);
        // m-tiles this item really needs (row scheduling hands out partial chunks), and this warp's share of them
        const int mt_item = p.row_sched ? ((rows - 1) * (int)wp + p.W + 127) / 128 : MT;
        const int mt_w = (mt_item - iw + p.n_iss - 1) / p.n_iss;
        const int buf = p.acc_bufs == 2 ? (lt & 1) : 0;
        const uint32_t cph = p.acc_bufs == 2 ? ((lt >> 1) & 1) : (lt & 1);
        long long m0 = clock64();
        mbar_wait(acc_empty0 + 8u * buf, cph ^ 1);
        long long m1 = clock64();
        if (rec) atomicAdd(p.dbg + blockIdx.x * 16 + 2, (unsigned long long)(m1 - m0));
        tc_fence_after();
        const uint32_t dbase = tmem_base + buf * 256u + m_d;
        for (int cg = 0; cg < ncg; ++cg) {
          long long m2 = clock64();
          if (p.b_resident && lt == 0) mbar_wait(b_full(cg), 0);   // resident weights of this channel group have landed
          mbar_wait(a_full0 + 8u * slice, aph);
          if (rec) atomicAdd(p.dbg + blockIdx.x * 16 + 3, (unsigned long long)(clock64() - m2));
          tc_fence_after();
          const uint32_t as = a0_16 + slice * a_slice16 + m_a;
          if (elect_one()) {
            mt_rt = mt_w;
            switch (mt_w) {
              case 1: issue_cg(std::integral_constant<int, 1>{}, cg, as, dbase, bs, bph); break;
              case 2: issue_cg(std::integral_constant<int, 2>{}, cg, as, dbase, bs, bph); break;
              case 3: issue_cg(std::integral_constant<int, 3>{}, cg, as, dbase, bs, bph); break;
              case 4: issue_cg(std::integral_constant<int, 4>{}, cg, as, dbase, bs, bph); break;
              // runtime loop; also mt_w == 0 (a partial chunk leaves this warp no m-tile): the weight-block waits and all
              // commits still happen, so the barrier phases stay in step
              default: issue_cg(std::integral_constant<int, 0>{}, cg, as, dbase, bs, bph); break;
            }
            umma_commit(a_empty0 + 8u * slice);
            if (cg == ncg - 1) umma_commit(acc_full0 + 8u * buf);
          }
          __syncwarp();
          if (!p.b_resident) { bs += 9; while (bs >= bstages) { bs -= bstages; bph ^= 1; } }
          if (++slice == p.n_slices) { slice = 0; aph ^= 1; }
        }
        if (rec) atomicAdd(p.dbg + blockIdx.x * 16 + 4, (unsigned long long)(clock64() - m1));
      }
    }
  } else {
    // ===================================================================== epilogue (8 warps after the MMA warps)
    const int q = warp & 3;
    const int part = (warp - (kHaloProd + kHaloIssue)) >> 2;
    const int row = q * 32 + lane;
    const int per_img = p.Hb * p.Wp;
    int lt = 0;
    for (HaloItem it = item_begin(); item_valid(it); item_next(it), ++lt) {
      int n0, h0, c_out0, rows;
      item_get(it, n0, h0, c_out0, rows);
      const int mt_item = p.row_sched ? ((rows - 1) * p.Wp + p.W + 127) / 128 : p.MT;
      const int buf = p.acc_bufs == 2 ? (lt & 1) : 0;
      const uint32_t cph = p.acc_bufs == 2 ? ((lt >> 1) & 1) : (lt & 1);
      long long e0 = clock64();
      mbar_wait(acc_full0 + 8u * buf, cph);
      long long e1 = clock64();
      tc_fence_after();
      for (int m = 0; m < ((p.dbg_flags & 2) ? 0 : mt_item); ++m) {
        const int o = m * 128 + row;
        const int n = o / per_img, rem = o % per_img, rr = rem / p.Wp, cc = rem % p.Wp;
        const int oh = h0 + rr, on = n0 + n;
        const bool ok = n < p.TN && rr < rows && cc < p.W && oh < p.H && on < p.N;
        const size_t pix = ((size_t)on * p.H + oh) * p.W + cc;
        const uint32_t lane_addr = tmem_base + buf * 256u + m * acc_cols + ((uint32_t)(q * 32) << 16);
        for (int c = part * 16; c < p.NT; c += 16 * kEpiParts) {
          const int co = c_out0 + c;
          uint4 rh0 = make_uint4(0, 0, 0, 0), rh1 = rh0, rl0 = rh0, rl1 = rh0;
          const bool has_res = ok && p.res_hi != nullptr;
          if (has_res) {
            const size_t ro = pix * p.res_ctot + p.res_coff + co;
            const uint4 *rh = reinterpret_cast<const uint4 *>(p.res_hi + ro);
            if (p.vec32) ldg256(rh, rh0, rh1); else { rh0 = __ldg(rh); rh1 = __ldg(rh + 1); }
            if (SPLIT && p.res_lo) {
              const uint4 *rl = reinterpret_cast<const uint4 *>(p.res_lo + ro);
              if (p.vec32) ldg256(rl, rl0, rl1); else { rl0 = __ldg(rl); rl1 = __ldg(rl + 1); }
            }
          }
          uint32_t v0[16], v1[16];
          tmem_ld16(lane_addr + c, v0);
          if (SPLIT) tmem_ld16(lane_addr + p.NT + c, v1);
          tmem_ld_wait();
          if (!ok) continue;
          float r[16];
          {
            const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + co);   // co % 16 == 0: 64-byte aligned
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 bq = __ldg(b4 + q4);
              r[4 * q4] = bq.x; r[4 * q4 + 1] = bq.y; r[4 * q4 + 2] = bq.z; r[4 * q4 + 3] = bq.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float x = __uint_as_float(v0[j]);
            if (SPLIT) x += __uint_as_float(v1[j]) * kLoInv;
            r[j] += x;
          }
          if (has_res) {
            const __half *hh0 = reinterpret_cast<const __half *>(&rh0), *hh1 = reinterpret_cast<const __half *>(&rh1);
            const __half *ll0 = reinterpret_cast<const __half *>(&rl0), *ll1 = reinterpret_cast<const __half *>(&rl1);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              r[j] += __half2float(hh0[j]) + __half2float(ll0[j]) * kLoInv;
              r[8 + j] += __half2float(hh1[j]) + __half2float(ll1[j]) * kLoInv;
            }
          }
          __align__(16) __half hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_store(p.relu ? fmaxf(r[j], 0.f) : r[j], hi[j], lo[j]);
          const size_t oo = pix * p.out_ctot + p.out_coff + co;
          const uint4 *hv = reinterpret_cast<const uint4 *>(hi), *lv = reinterpret_cast<const uint4 *>(lo);
          uint4 *oh4 = reinterpret_cast<uint4 *>(p.out_hi + oo);
          if (p.vec32) stg256(oh4, hv[0], hv[1]); else { oh4[0] = hv[0]; oh4[1] = hv[1]; }
          if (SPLIT && p.out_lo) {
            uint4 *ol4 = reinterpret_cast<uint4 *>(p.out_lo + oo);
            if (p.vec32) stg256(ol4, lv[0], lv[1]); else { ol4[0] = lv[0]; ol4[1] = lv[1]; }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty0 + 8u * buf) : "memory");
      if (p.dbg && threadIdx.x == 32 * (kHaloProd + kHaloIssue)) { long long e2 = clock64(); atomicAdd(p.dbg + blockIdx.x * 16 + 5, (unsigned long long)(e1 - e0)); atomicAdd(p.dbg + blockIdx.x * 16 + 6, (unsigned long long)(e2 - e1)); atomicAdd(p.dbg + blockIdx.x * 16 + 7, 1ull); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    unsigned long long g_exit;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_exit));
    p.dbg[blockIdx.x * 16 + 10] = g_exit;
    p.dbg[blockIdx.x * 16 + 11] = (unsigned long long)(clock64() - c_entry);     // cycles in the kernel
  }
  if (warp == kHaloProd) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host side
// Grid of a persistent kernel over `items` equal work items.  min(items, SMs) leaves a ragged last round: 896 items on
// 148 CTAs are 8 CTAs with 7 items and 140 with 6, and the 140 SMs then sit under the early-launched (PDL) CTAs of the
// next layer of the same lane, which wait for the 8 stragglers.  With CTAs = ceil(items / rounds) (896 -> 128 x 7) every
// CTA finishes together and the SMs the grid does not use run another lane's kernel for the whole time.
// SHAPY_CONV_EVENGRID=0 restores min(items, SMs).
static int pdl_late_enabled() {
  // SHAPY_PDL_LATE: 0 = release at entry, 1 = release at the last tile, 2 = 1 + per-role wait in the halo kernel (measured neutral: profiles/r02_pdl_rolewait_ab.txt); default 1
  static const int v = []() { const char *e = getenv("SHAPY_PDL_LATE"); return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1; }();
  return v;
}

static int even_grid(int items, int sms) {
  static const bool on = []() { const char *e = getenv("SHAPY_CONV_EVENGRID"); return !(e && e[0] == '0'); }();
  if (items <= sms || !on) return std::min(items, sms);
  const int rounds = ceil_div(items, sms);
  return ceil_div(items, rounds);
}

struct UmmaPlan {
  UmmaParams p;
  HaloParams hp;
  bool halo = false;
  int kch;
  bool split;
  dim3 grid;
  size_t smem;
};

static int pick_kch(int cin) { return cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 16); }
static int pick_nt(int cout) {
  for (int nt = 128; nt >= 16; nt -= 16)
    if (cout % nt == 0) return nt;
  return 0;
}

bool umma_supported(const ConvW &w, const ActView &in, const ActView &out) {
  if (w.cin % 16 || w.cout % 16 || !pick_nt(w.cout)) return false;
  if (!((w.ksize == 3 || w.ksize == 1) && (w.stride == 1 || (w.stride == 2 && w.ksize == 3)))) return false;
  if (w.stride == 2 && (in.H % 2 || in.W % 2)) return false;
  if (in.Ctot % 8 || in.coff % 8 || out.Ctot % 8 || out.coff % 8) return false;
  if (out.W > 128 && out.W % 128) return false;
  return get_encode() != nullptr;
}

// 256-bit epilogue accesses need every row of the output (and residual) slice to start on a 32-byte boundary
static int rows_vec32(const ActView &out, const ActView *res) {
  auto ok = [](const ActView &v) {
    return v.Ctot % 16 == 0 && v.coff % 16 == 0 && ((uintptr_t)v.hi & 31) == 0 && (!v.lo || ((uintptr_t)v.lo & 31) == 0);
  };
  return ok(out) && (!res || ok(*res)) ? 1 : 0;
}

static bool halo_enabled() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("SHAPY_CONV_HALO"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
// Measured (profiles/r01_*): with 64-channel k-blocks the halo variant's 16-byte cp.async fill of ~90 KB slices is
// slower than re-streaming taps through TMA, so those layers stay on the per-tap kernel.
static int halo_max_kch() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("SHAPY_CONV_HALO_MAXKCH"); v = e ? atoi(e) : 32; }
  return v;
}

// stride-2 layers: 0 disables the halo variant for them (SHAPY_CONV_HALO_S2_MAXKCH)
static int halo_s2_max_kch() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("SHAPY_CONV_HALO_S2_MAXKCH"); v = e ? atoi(e) : 32; }
  return v;
}

// Chooses the super-tile of the halo-resident kernel by a small traffic / MMA cost model.
static bool halo_configure(UmmaPlan *pl, const ConvW &w, const ActView &in, const ActView &out, const ActView *res,
                           bool relu) {
  const bool split = in.lo != nullptr;
  const int parts = split ? 2 : 1;
  const int kch = pick_kch(w.cin);
  // stride 1: one plane of the input band with a 1-pixel halo.  stride 2: four parity planes (row parity x column parity)
  // of output size + 1; H, W are the OUTPUT extent in both cases.
  const int stride = w.stride;
  const int H = out.H, W = out.W, N = out.N;
  const int Wp = stride == 1 ? W + 2 : W + 1, hpad = stride == 1 ? 2 : 1, nplanes = stride == 1 ? 1 : 4;
  const int maxshift = stride == 1 ? 2 * Wp + 2 : Wp + 1;
  if (Wp > 256) return false;
  struct Cfg { int NT, TN, R, MT, Hb, bands, n_super, bres; double cost; } best = {0, 0, 0, 0, 0, 0, 0, 0, 1e30};
  int f_nt = 0, f_tn = 0, f_r = 0;   // experiments: SHAPY_HALO_FORCE="NT,TN,R" pins the tile configuration
  if (const char *e = getenv("SHAPY_HALO_FORCE")) sscanf(e, "%d,%d,%d", &f_nt, &f_tn, &f_r);
  // L2 -> SM bytes per cycle per SM the TMA loads can count on: the per-tap C=192 layer moves 296 MB in 30 us = 9.9 TB/s
  // = 41 B/cycle/SM at the 1.63 GHz these kernels run at (the first-generation kernels only reached 19.6, which made
  // the model avoid streamed weights: C=96 with R=4 / MT=1 / double-buffered accumulators is 9 % faster than R=7 / MT=2)
  const double sm_count = 148.0, l2_bpc = []() { const char *e = getenv("SHAPY_HALO_L2BPC"); return e ? atof(e) : 40.0; }();
  for (int NT = 128; NT >= 16; NT -= 16) {
    if (w.cout % NT || (f_nt && NT != f_nt)) continue;
    const int mt_max = std::min(8, 512 / (NT * parts));
    if (mt_max < 1) continue;
    const int n_tiles = w.cout / NT;
    const size_t b_blk = align_up((size_t)NT * kch * 2 * parts, 1024);
    const size_t b_all = b_blk * 9 * (w.cin / kch);
    const bool bres = n_tiles == 1 && b_all <= 120 * 1024;
    auto consider = [&](int TN, int R) {
      if (f_nt && (TN != f_tn || R != f_r)) return;
      const int Hb = R + hpad;
      const int bands = ceil_div(H, R);
      const int P = TN * Hb * Wp;
      const int last = ((TN - 1) * Hb + (R - 1)) * Wp + W;   // positions that produce valid outputs
      const int MT = ceil_div(last, 128);
      if (MT > mt_max) return;
      const size_t slice = (size_t)parts * nplanes * align_up((size_t)std::max(P, MT * 128 + maxshift) * kch * 2, 1024);
      if (2 * slice + (bres ? b_all : 2 * b_blk) > 216 * 1024) return;
      const int n_super = ceil_div(N, TN) * bands;
      const double items = (double)n_super * n_tiles;
      const double a_bytes = (double)nplanes * P * w.cin * 2 * parts;
      const double b_bytes = bres ? 0.0 : 9.0 * NT * w.cin * 2 * parts;
      const double mem = (a_bytes + b_bytes) / l2_bpc;
      // MMA time per k-step.  tools/umma_bench.cu measures max(32 + N/4, N/2) cycles (operand reads at 128 B/cycle or
      // the tensor pipe); inside the kernel the issuing warps add ~30 %, which the 85 B/cycle figure below stands for
      // (the tile configurations it selects were validated layer by layer, see profiles/README.md).
      auto mma_cyc = [](double n) { return std::max((4096.0 + 32.0 * n) / 85.0, n * 0.5); };
      const double per_k16 = split ? mma_cyc(2.0 * NT) + mma_cyc(NT) : mma_cyc(NT);
      const double mma = MT * 9.0 * (w.cin / 16) * per_k16;
      const bool dbl = MT * NT * parts <= 256;
      const double epi = (dbl ? 0.25 : 1.0) * MT * (NT / 32.0) * 1500.0;
      const double cost = std::ceil(items / sm_count) * (std::max(mem, mma) + epi + 1500.0);
      if (cost < best.cost) best = {NT, TN, R, MT, Hb, bands, n_super, bres ? 1 : 0, cost};
    };
    for (int TN = 1; TN <= std::min(N, 8); ++TN) consider(TN, H);   // whole images
    for (int R = 1; R < H; ++R) consider(1, R);                     // bands of one image
  }
  if (best.NT && stride == 1) {
    // per-tap kernel estimate: every (m-tile, n-tile) streams 9 taps of A and B through L2
    const int TWt = std::min(W, 128), THt = std::max(1, std::min(H, 128 / TWt));
    const int TNt = (TWt == W && THt == H) ? std::max(1, std::min(N, 128 / (TWt * THt))) : 1;
    const double m_tiles = (double)ceil_div(W, TWt) * ceil_div(H, THt) * ceil_div(N, TNt);
    const int NTt = pick_nt(w.cout);
    const double bytes = m_tiles * (w.cout / NTt) * 9.0 * w.cin * 2 * parts * (TWt * THt * TNt + NTt);
    const double pt_cost = bytes / l2_bpc / sm_count + 3000.0;
    if (getenv("SHAPY_CONV_DEBUG"))
      fprintf(stderr, "[halo] model: halo %.0f cycles vs per-tap %.0f cycles\n", best.cost, pt_cost);
    if (best.cost > pt_cost) return false;
  }
  if (!best.NT) return false;
  HaloParams &p = pl->hp;
  memset(&p, 0, sizeof(p));
  p.stride = stride;
  p.N = N; p.H = H; p.W = W; p.Wp = Wp; p.Hb = best.Hb; p.TN = best.TN; p.R = best.R; p.bands = best.bands;
  p.n_super = best.n_super; p.MT = best.MT;
  p.cin = w.cin; p.cout = w.cout; p.NT = best.NT; p.n_tiles = w.cout / best.NT; p.ncg = w.cin / kch;
  p.acc_bufs = (best.MT * best.NT * parts <= 256) ? 2 : 1;
  p.n_iss = best.MT >= 2 ? kHaloIssue : 1;
  {
    const char *e = getenv("SHAPY_CONV_ROWSCHED");
    p.row_sched = (stride == 1 && best.TN == 1 && p.n_tiles == 1 && best.R < H && e && e[0] == '1') ? 1 : 0;   // off by default: no gain measured
  }
  if (const char *e = getenv("SHAPY_CONV_ISS")) p.n_iss = std::max(1, std::min(p.n_iss, atoi(e)));   // experiments
  const int P = best.TN * best.Hb * Wp;
  {
    const char *b = getenv("SHAPY_CONV_BASEOFF");
    p.a_baseoff = (b && b[0] == '1') ? 1 : 0;   // measured: the swizzle is a function of the absolute address, no base offset needed
  }
  // [position][kch] rows; rows past the box (read by the last m-tile's shifted taps) are never written
  p.plane_bytes = (uint32_t)align_up((size_t)std::max(P, best.MT * 128 + maxshift) * kch * 2, 1024);
  p.part_bytes = p.plane_bytes * nplanes;
  p.a_slice_bytes = p.part_bytes * parts;
  p.a_tx = (uint32_t)((size_t)parts * nplanes * P * kch * 2);
  p.b_bytes = (uint32_t)best.NT * kch * 2;
  p.b_stride = (uint32_t)align_up((size_t)p.b_bytes * parts, 1024);
  const size_t b_stage = (size_t)p.b_stride;
  p.b_resident = best.bres;
  p.n_slices = 2;
  if (p.b_resident) {
    p.bstages = 9 * p.ncg;
  } else {
    const size_t avail = 214 * 1024 - 2 * (size_t)p.a_slice_bytes;
    p.bstages = (int)std::max<size_t>(2, std::min<size_t>(9, avail / b_stage));
  }
  if (2 * (size_t)p.a_slice_bytes + b_stage * p.bstages > 222 * 1024) return false;
  {
    // a third A slice when it fits without shrinking the weight ring (SHAPY_CONV_ASLICES=2 for A/B runs)
    static const int want = []() { const char *e = getenv("SHAPY_CONV_ASLICES"); return (e && e[0] == '2') ? 2 : 3; }();
    if (want == 3 && 3 * (size_t)p.a_slice_bytes + b_stage * p.bstages <= 222 * 1024) p.n_slices = 3;
  }
  p.idesc = (1u << 4) | ((uint32_t)(best.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.idesc2 = (1u << 4) | ((uint32_t)((2 * best.NT) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.relu = relu;
  p.pdl_late = pdl_late_enabled();
  p.vec32 = rows_vec32(out, res);
  p.bias = w.bias;
  p.out_hi = out.hi; p.out_lo = out.lo; p.out_ctot = out.Ctot; p.out_coff = out.coff;
  p.res_hi = res ? res->hi : nullptr; p.res_lo = res ? res->lo : nullptr;
  p.res_ctot = res ? res->Ctot : 0; p.res_coff = res ? res->coff : 0;
  p.dbg = nullptr;
  pl->smem = (size_t)p.n_slices * p.a_slice_bytes + b_stage * p.bstages + 16 * p.bstages + 128 + 1024;
  if (getenv("SHAPY_CONV_DEBUG"))
    fprintf(stderr, "[halo] cin %d cout %d %dx%d N %d kch %d: NT %d TN %d R %d MT %d bands %d items %d part %u slice %u bstages %d acc_bufs %d bres %d smem %zu\n",
            w.cin, w.cout, H, W, N, kch, p.NT, p.TN, p.R, p.MT, p.bands, p.n_super * p.n_tiles, p.part_bytes,
            p.a_slice_bytes, p.bstages, p.acc_bufs, p.b_resident, pl->smem);
  if (getenv("SHAPY_CONV_DEBUG")) fprintf(stderr, "[halo]   A slices in flight: %d\n", p.n_slices);
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    pl->grid = dim3(p.row_sched ? std::min(N * H, sms) : even_grid(p.n_super * p.n_tiles, sms), 1);
  }
  cuuint32_t box[4] = {(cuuint32_t)kch, (cuuint32_t)Wp, (cuuint32_t)best.Hb, (cuuint32_t)best.TN};
  bool ok = true;
  for (int pl_i = 0; pl_i < nplanes && ok; ++pl_i) {
    cuuint64_t dims[4], strides[3];
    size_t off;
    if (stride == 1) {
      dims[0] = in.C; dims[1] = in.W; dims[2] = in.H; dims[3] = in.N;
      strides[0] = (cuuint64_t)in.Ctot * 2; strides[1] = (cuuint64_t)in.W * in.Ctot * 2;
      strides[2] = (cuuint64_t)in.H * in.W * in.Ctot * 2;
      off = in.coff;
    } else {   // parity view: rows 2y + py, columns 2x + px of the input
      const int py = pl_i >> 1, px = pl_i & 1;
      dims[0] = in.C; dims[1] = in.W / 2; dims[2] = in.H / 2; dims[3] = in.N;
      strides[0] = (cuuint64_t)in.Ctot * 4; strides[1] = (cuuint64_t)in.W * in.Ctot * 4;
      strides[2] = (cuuint64_t)in.H * in.W * in.Ctot * 2;
      off = (size_t)in.coff + ((size_t)py * in.W + px) * in.Ctot;
    }
    ok = ok && encode(&p.a_hi[pl_i], in.hi + off, 4, dims, strides, box, kch);
    if (split) ok = ok && encode(&p.a_lo[pl_i], in.lo + off, 4, dims, strides, box, kch);
  }
  cuuint64_t bd[3] = {(cuuint64_t)w.cin, (cuuint64_t)w.cout, 9};
  cuuint64_t bs[2] = {(cuuint64_t)w.cin * 2, (cuuint64_t)w.cin * w.cout * 2};
  cuuint32_t bb[3] = {(cuuint32_t)kch, (cuuint32_t)best.NT, 1};
  ok = ok && encode(&p.b_hi, w.w_hi, 3, bd, bs, bb, kch);
  if (split) ok = ok && encode(&p.b_lo, w.w_lo, 3, bd, bs, bb, kch);
  return ok;
}

UmmaPlan *umma_plan_create(const ConvW &w, const ActView &in, const ActView &out, const ActView *res, bool relu) {
  if (!umma_supported(w, in, out)) {
    set_error("umma_plan_create: unsupported convolution (cin %d cout %d k %d s %d)", w.cin, w.cout, w.ksize, w.stride);
    return nullptr;
  }
  auto *pl = new UmmaPlan();
  memset(&pl->p, 0, sizeof(pl->p));
  UmmaParams &p = pl->p;
  const bool split = in.lo != nullptr;
  int kch = pick_kch(w.cin);
  if (w.ksize == 3 && halo_enabled() && kch <= (w.stride == 1 ? halo_max_kch() : halo_s2_max_kch()) &&
      halo_configure(pl, w, in, out, res, relu)) {
    pl->kch = kch;
    pl->split = split;
    pl->halo = true;
    return pl;
  }
  // K padding: channel counts that are not a multiple of 64 still use 64-channel (128-byte, SWIZZLE_128B) k-blocks; the
  // TMA unit zero-fills the channels past cin of both operands.  These layers are bound by TMA requests and barrier
  // round trips, not by the tensor pipe: one k-block per tap instead of three was 10-18 % faster on every cin = 48 / 96
  // layer (SHAPY_CONV_KPAD=0 restores exact-width k-blocks).
  {
    const char *e = getenv("SHAPY_CONV_KPAD");
    if (!(e && e[0] == '0') && kch < 64) kch = 64;
  }
  pl->kch = kch;
  pl->split = split;
  p.N = out.N; p.Ho = out.H; p.Wo = out.W;
  p.TW = std::min(out.W, 128);
  p.TH = std::max(1, std::min(out.H, 128 / p.TW));
  p.TN = (p.TW == out.W && p.TH == out.H) ? std::max(1, std::min(out.N, 128 / (p.TW * p.TH))) : 1;
  p.tiles_w = ceil_div(out.W, p.TW); p.tiles_h = ceil_div(out.H, p.TH); p.tiles_n = ceil_div(out.N, p.TN);
  p.cin = w.cin; p.cout = w.cout; p.ksize = w.ksize; p.stride = w.stride;
  {
    // cout tile: balance SM utilisation (tiles per 148 CTAs) against the per-MMA operand-read cost (see halo model)
    const double m_tiles = (double)p.tiles_w * p.tiles_h * p.tiles_n;
    auto mma_cyc = [](double n) { return std::max((4096.0 + 32.0 * n) / 128.0, n * 0.5); };
    double best = 1e30;
    int best_nt = pick_nt(w.cout);
    for (int nt = 128; nt >= 16; nt -= 16) {
      if (w.cout % nt) continue;
      const double per_k = split ? mma_cyc(2.0 * nt) + mma_cyc(nt) : mma_cyc(nt);
      const double cost = std::ceil(m_tiles * (w.cout / nt) / 148.0) * (per_k * w.ksize * w.ksize * (w.cin / 16) + 400.0 * nt / 16);
      if (cost < best) { best = cost; best_nt = nt; }
    }
    p.NT = best_nt;
    if (const char *e = getenv("SHAPY_CONV_NT")) { const int v = atoi(e); if (v >= 16 && v <= 128 && w.cout % v == 0) p.NT = v; }   // experiments
  }
  p.kpt = ceil_div(w.cin, kch);
  p.relu = relu;
  p.pdl_late = pdl_late_enabled();
  p.vec32 = rows_vec32(out, res);
  // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=f16 (0), K-major A and B, N>>3 at 17, M>>4 at 24
  p.idesc = (1u << 4) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t cols = p.NT * (split ? 2 : 1), pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  p.tmem_cols = pow2;
  const int rows = p.TW * p.TH * p.TN;
  p.a_bytes = (uint32_t)rows * kch * 2;
  p.b_bytes = (uint32_t)p.NT * kch * 2;
  p.b_stride = (uint32_t)align_up((size_t)p.b_bytes * (split ? 2 : 1), 1024);
  p.idesc2 = (1u << 4) | ((uint32_t)((2 * p.NT) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t kblk = 128u * kch * 2 * (split ? 2 : 1) + p.b_stride;
  // k-blocks per stage: amortise barrier round trips for thin k-blocks, stay <= 48 KB per stage
  int G = 1;
  for (int g = p.kpt; g >= 1; --g)
    if (p.kpt % g == 0 && (size_t)g * kblk <= 49152) { G = g; break; }
  p.G = G;
  const size_t stage = (size_t)G * kblk;
  if (getenv("SHAPY_CONV_DEBUG"))
    fprintf(stderr, "[per-tap] cin %d cout %d k %d s %d %dx%d N %d kch %d: NT %d tile %dx%dx%d tiles %d G %d\n", w.cin, w.cout, w.ksize,
            w.stride, out.H, out.W, out.N, kch, p.NT, p.TW, p.TH, p.TN, p.tiles_w * p.tiles_h * p.tiles_n * (w.cout / p.NT), G);
  const int iters = p.ksize * p.ksize * p.kpt / G;
  // one persistent CTA per SM owns (almost) all of its shared memory
  int stages = (int)std::min<size_t>(8, (200 * 1024) / stage);
  stages = std::max(2, stages);
  (void)iters;
  p.stages = stages;
  pl->smem = stage * stages + 16 * stages + 64 + 1024;
  p.bias = w.bias;
  p.out_hi = out.hi; p.out_lo = out.lo; p.out_ctot = out.Ctot; p.out_coff = out.coff;
  p.res_hi = res ? res->hi : nullptr; p.res_lo = res ? res->lo : nullptr;
  p.res_ctot = res ? res->Ctot : 0; p.res_coff = res ? res->coff : 0;
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int total = p.tiles_w * p.tiles_h * p.tiles_n * (w.cout / p.NT);
    pl->grid = dim3(even_grid(total, sms), 1);
  }
  bool ok = true;
  // activation maps
  const int nmaps = w.stride == 2 ? 4 : 1;
  for (int mi = 0; mi < nmaps && ok; ++mi) {
    const int phy = mi >> 1, phx = mi & 1;
    cuuint64_t dims[4], strides[3];
    size_t off;
    if (w.stride == 1) {
      dims[0] = in.C; dims[1] = in.W; dims[2] = in.H; dims[3] = in.N;
      strides[0] = (cuuint64_t)in.Ctot * 2; strides[1] = (cuuint64_t)in.W * in.Ctot * 2;
      strides[2] = (cuuint64_t)in.H * in.W * in.Ctot * 2;
      off = in.coff;
    } else {
      dims[0] = in.C; dims[1] = in.W / 2; dims[2] = in.H / 2; dims[3] = in.N;
      strides[0] = (cuuint64_t)in.Ctot * 4; strides[1] = (cuuint64_t)in.W * in.Ctot * 4;
      strides[2] = (cuuint64_t)in.H * in.W * in.Ctot * 2;
      off = (size_t)in.coff + ((size_t)phy * in.W + phx) * in.Ctot;
    }
    cuuint32_t box[4] = {(cuuint32_t)kch, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
    ok = ok && encode(&p.a_hi[mi], in.hi + off, 4, dims, strides, box, kch);
    if (split) ok = ok && encode(&p.a_lo[mi], in.lo + off, 4, dims, strides, box, kch);
  }
  if (ok) {
    cuuint64_t dims[3] = {(cuuint64_t)w.cin, (cuuint64_t)w.cout, (cuuint64_t)(w.ksize * w.ksize)};
    cuuint64_t strides[2] = {(cuuint64_t)w.cin * 2, (cuuint64_t)w.cin * w.cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)kch, (cuuint32_t)p.NT, 1};
    ok = ok && encode(&p.b_hi, w.w_hi, 3, dims, strides, box, kch);
    if (split) ok = ok && encode(&p.b_lo, w.w_lo, 3, dims, strides, box, kch);
  }
  if (!ok) { delete pl; return nullptr; }
  return pl;
}

void umma_plan_destroy(UmmaPlan *p) { delete p; }

template <typename Kernel, typename Params>
static int launch_pdl(Kernel kernel, dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &params, bool pdl) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  SHAPY_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, params));
  shapy::count_launch();
  return SHAPY_OK;
}

template <int KCH, bool SPLIT>
static int launch_t(const UmmaPlan *pl, cudaStream_t st, bool pdl) {
  static std::atomic<unsigned long long> attr_done{0};
  SHAPY_CUDA_TRY(set_max_dynamic_smem(conv_umma_kernel<KCH, SPLIT>, 227 * 1024, attr_done));
  return launch_pdl(conv_umma_kernel<KCH, SPLIT>, pl->grid, kThreads, pl->smem, st, pl->p, pdl);
}

template <int KCH, bool SPLIT>
static int launch_halo_t(const UmmaPlan *pl, cudaStream_t st, bool pdl) {
  static std::atomic<unsigned long long> attr_done{0};
  SHAPY_CUDA_TRY(set_max_dynamic_smem(conv_halo_kernel<KCH, SPLIT>, 227 * 1024, attr_done));
  static const bool phases = getenv("SHAPY_CONV_PHASES") != nullptr;
  if (phases) {
    // debug: per-CTA cycle counters of each warp role's waits (synchronous, prints to stderr)
    HaloParams hp = pl->hp;
    unsigned long long *d = nullptr;
    cudaMalloc(&d, 148 * 16 * 8);
    cudaMemset(d, 0, 148 * 16 * 8);
    hp.dbg = d;
    if (const char *f = getenv("SHAPY_CONV_DBGFLAGS")) hp.dbg_flags = atoi(f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    conv_halo_kernel<KCH, SPLIT><<<pl->grid, kHaloThreads, pl->smem, st>>>(hp);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148 * 16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long sum[8] = {0};
    for (int c = 0; c < (int)pl->grid.x; ++c) for (int k = 0; k < 8; ++k) sum[k] += h[c * 16 + k];
    const double n = pl->grid.x;
    fprintf(stderr, "[phases] cin %d cout %d %dx%d MT %d NT %d: %.1f us | per CTA cycles: prod wait_empty %.0f fill %.0f | mma wait_acc %.0f wait_a %.0f busy_total %.0f | epi wait_full %.0f work %.0f items %.1f\n",
            hp.cin, hp.cout, hp.H, hp.W, hp.MT, hp.NT, ms * 1e3, sum[0] / n, sum[1] / n, sum[2] / n, sum[3] / n, sum[4] / n, sum[5] / n, sum[6] / n, sum[7] / n);
    {
      unsigned long long g0 = ~0ull, g0max = 0, g1 = 0, pro = 0, cyc = 0;
      for (int c = 0; c < (int)pl->grid.x; ++c) {
        g0 = std::min(g0, h[c * 16 + 9]); g0max = std::max(g0max, h[c * 16 + 9]); g1 = std::max(g1, h[c * 16 + 10]);
        pro += h[c * 16 + 8]; cyc = std::max(cyc, h[c * 16 + 11]);
      }
      fprintf(stderr, "[phases]   first entry -> last exit %.1f us, entry skew %.1f us, prologue %.0f cycles avg, longest CTA %llu cycles (%.2f GHz)\n",
              (g1 - g0) * 1e-3, (g0max - g0) * 1e-3, pro / n, cyc, cyc / ((g1 - g0) * 1.0));
    }
    cudaFree(d);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    shapy::count_launch();
    return SHAPY_OK;
  }
  return launch_pdl(conv_halo_kernel<KCH, SPLIT>, pl->grid, kHaloThreads, pl->smem, st, pl->hp, pdl);
}

int umma_plan_launch(const UmmaPlan *pl, cudaStream_t st, bool pdl) {
  if (pl->halo) {
    if (pl->split) {
      if (pl->kch == 64) return launch_halo_t<64, true>(pl, st, pdl);
      if (pl->kch == 32) return launch_halo_t<32, true>(pl, st, pdl);
      return launch_halo_t<16, true>(pl, st, pdl);
    }
    if (pl->kch == 64) return launch_halo_t<64, false>(pl, st, pdl);
    if (pl->kch == 32) return launch_halo_t<32, false>(pl, st, pdl);
    return launch_halo_t<16, false>(pl, st, pdl);
  }
  if (pl->split) {
    if (pl->kch == 64) return launch_t<64, true>(pl, st, pdl);
    if (pl->kch == 32) return launch_t<32, true>(pl, st, pdl);
    return launch_t<16, true>(pl, st, pdl);
  }
  if (pl->kch == 64) return launch_t<64, false>(pl, st, pdl);
  if (pl->kch == 32) return launch_t<32, false>(pl, st, pdl);
  return launch_t<16, false>(pl, st, pdl);
}

}  // namespace shapy
