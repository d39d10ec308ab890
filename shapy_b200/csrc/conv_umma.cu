// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (hand-written PTX, no CUTLASS).
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel shifted by tap, cin] * Wt[tap][cout][cin]
//
// GEMM view: M = 128 output pixels (a TW x TH x TN box of the NHWC output), N = cout tile (<= 128),
// K = taps * cin, walked in k-blocks of KCH in {16, 32, 64} channels of one filter tap.
//  * A operand: one 4-D TMA box per (tap, k-block) straight out of the NHWC activation plane, the box
//    origin shifted by the tap; out-of-image rows/columns are ZERO-FILLED by the TMA unit, which is the
//    convolution's padding -- no im2col buffer, no halo storage.  Stride-2 layers use four parity views
//    (even/odd row x even/odd column) of the same plane so every tap is again a dense box.
//  * B operand: 3-D TMA box of the pre-packed [tap][cout][cin] fp16 weights (BN scale folded in).
//  * Both land in shared memory in the canonical K-major SWIZZLE_{32,64,128}B layout (swizzle = KCH*2 B
//    rows) that the UMMA shared-memory descriptor names, so no thread ever touches operand data.
//  * One elected thread issues tcgen05.mma (M=128, N=cout tile, K=16) into fp32 TMEM accumulators;
//    split-fp16 parity mode issues hi*hi into D0 and hi*lo + lo*hi into D1 (out = D0 + D1 / 2048).
//  * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
//    (tcgen05.ld -> +bias (+residual) -> ReLU -> split to fp16 hi/lo -> 16-byte global stores).
//    smem ring of `stages` slots with full/empty mbarriers; tcgen05.commit releases slots and publishes
//    the accumulator.  Several CTAs share an SM (<= ~100 KB smem, <= 256 TMEM columns each) so one
//    CTA's epilogue overlaps another's main loop.
// Replaces cuDNN conv + BatchNorm + ReLU + residual add (4 launches, 4 HBM round trips) of
// regressor/human_shape/models/backbone/hrnet.py and torchvision BasicBlock / Bottleneck.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "conv.cuh"

namespace shapy {

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// Programmatic dependent launch: the next kernel of the stream may be launched while this one is still running
// (its prologue overlaps our tail); it blocks in pdl_wait() until every prior kernel has completed and flushed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

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
  uint32_t idesc, idesc2;        // idesc: N = NT;  idesc2: N = 2 NT ([B_hi | B_lo] in one MMA, split mode)
  uint32_t tmem_cols;
  uint32_t a_bytes, b_bytes, b_stride;  // per k-block: TMA bytes of A / B, smem pitch of a B block
  const float *bias;
  __half *out_hi, *out_lo;
  int out_ctot, out_coff;
  const __half *res_hi, *res_lo;
  int res_ctot, res_coff;
};

template <int KCH>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // K-major, swizzle = KCH * 2 bytes per row; 8-row groups are KCH * 16 bytes apart (SBO); LBO unused (1)
  constexpr uint64_t layout = KCH == 64 ? 2 : (KCH == 32 ? 4 : 6);
  constexpr uint64_t sbo = (KCH * 2 * 8) >> 4;
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// Persistent, warp-specialised kernel: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
//   warp 0      TMA producer  (smem ring runs ahead across tile boundaries: no pipeline drain per tile)
//   warp 1      TMEM allocator + MMA issuer (alternates between two TMEM accumulator buffers)
//   warps 2..9  epilogue: drains accumulator buffer t & 1 while the MMA warp fills the other one
constexpr int kEpiWarps = 8;
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
  pdl_launch_dependents();   // let the next launch start its prologue
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
#pragma unroll 1
          for (int it = 0; it < iters; ++it) {
            mbar_wait(full_bar(s_l), ph_l);
            tc_fence_after();
            const uint32_t sbase = smem_base + stage_bytes * s_l;
            for (int g = 0; g < p.G; ++g) {
              const uint32_t kb = sbase + kblk_bytes * g;
              const uint64_t da0 = make_desc<KCH>(kb), dal0 = make_desc<KCH>(kb + A_BLK);
              const uint64_t db0 = make_desc<KCH>(kb + A_BLK * (SPLIT ? 2 : 1));
              const uint32_t first = (it | g) ? 1u : 0u;
#pragma unroll
              for (int ks = 0; ks < KCH / 16; ++ks) {
                const uint64_t ko = (uint64_t)(ks * 2);   // +32 bytes along K, in 16-byte descriptor units
                if (SPLIT) {
                  // [D0 | D1] (+)= A_hi . [B_hi | B_lo]^T  (one N = 2 NT MMA), then D1 += A_lo . B_hi^T
                  umma_f16(d0, da0 + ko, db0 + ko, p.idesc2, first | (ks ? 1u : 0u));
                  umma_f16(d1, dal0 + ko, db0 + ko, p.idesc, 1);
                } else {
                  umma_f16(d0, da0 + ko, db0 + ko, p.idesc, first | (ks ? 1u : 0u));
                }
              }
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
    const int half = (warp - 2) >> 2;        // which alternate 16-column chunks this warp handles
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
      for (int c = half * 16; c < p.NT; c += 32) {
        const int co = c_out0 + c;
        // residual rows are fetched before the TMEM load completes so the two latencies overlap
        uint4 rh0 = make_uint4(0, 0, 0, 0), rh1 = rh0, rl0 = rh0, rl1 = rh0;
        const bool has_res = ok && p.res_hi != nullptr;
        if (has_res) {
          const size_t ro = pix * p.res_ctot + p.res_coff + co;
          const uint4 *rh = reinterpret_cast<const uint4 *>(p.res_hi + ro);
          rh0 = __ldg(rh); rh1 = __ldg(rh + 1);
          if (SPLIT && p.res_lo) {
            const uint4 *rl = reinterpret_cast<const uint4 *>(p.res_lo + ro);
            rl0 = __ldg(rl); rl1 = __ldg(rl + 1);
          }
        }
        uint32_t v0[16], v1[16];
        tmem_ld16(lane_addr + c, v0);
        if (SPLIT) tmem_ld16(lane_addr + p.NT + c, v1);
        tmem_ld_wait();
        if (!ok) continue;
        float r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = __uint_as_float(v0[j]);
          if (SPLIT) x += __uint_as_float(v1[j]) * kLoInv;
          r[j] = x + __ldg(p.bias + co + j);
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
        uint4 *oh4 = reinterpret_cast<uint4 *>(p.out_hi + oo);
        oh4[0] = reinterpret_cast<const uint4 *>(hi)[0];
        oh4[1] = reinterpret_cast<const uint4 *>(hi)[1];
        if (SPLIT && p.out_lo) {
          uint4 *ol4 = reinterpret_cast<uint4 *>(p.out_lo + oo);
          ol4[0] = reinterpret_cast<const uint4 *>(lo)[0];
          ol4[1] = reinterpret_cast<const uint4 *>(lo)[1];
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
// 3x3 stride-1 convolutions: "halo-resident" variant.
//
// The per-tap variant above re-reads every input pixel 9 times from L2 (one shifted box per tap) and is
// L2-bandwidth bound (ncu: ~5.5 TB/s of L2->SM traffic, tensor pipe 15 %).  Here the input of a
// super-tile (a band of R output rows, or TN whole small images) is loaded ONCE per 16/32/64-channel group,
// with its 1-pixel halo (TMA zero-fills the image border = the padding), into shared memory as UNSWIZZLED
// 8-channel planes:  plane[k8][position][8 x fp16], position = (n * Hb + row) * Wp + col on the padded grid.
// In that layout a row of the MMA A operand is 16 contiguous bytes at pitch 16, so the operand of tap
// (ky, kx) for output positions o .. o+127 is simply the same plane read from position o + ky * Wp + kx:
// all 9 taps x MT m-tiles are addressed by moving the start address of a no-swizzle K-major UMMA
// descriptor (LBO = plane stride, SBO = 128 B).  Outputs are computed for the Wp - W padding columns too
// (and discarded by the epilogue); in exchange activations cross L2 ~1.2-1.5x instead of 9x, and the
// weights of one (tap, channel group) are shared by up to 5 m-tiles held in TMEM at once.
struct alignas(64) HaloParams {
  CUtensorMap a_hi, a_lo;  // (8ch, W, H, N) unswizzled, box (8, Wp, Hb, TN)
  CUtensorMap b_hi, b_lo;  // (cin, cout, 9) swizzled, box (KCH, NT, 1)
  int N, H, W;
  int Wp, Hb, TN, R, bands, n_super, MT;
  int cin, cout, NT, n_tiles, ncg, bstages, acc_bufs, b_resident;
  uint32_t PS, a_slice_bytes, a_tx, b_bytes, b_stride;
  uint32_t idesc, idesc2;
  int relu;
  const float *bias;
  __half *out_hi, *out_lo;
  int out_ctot, out_coff;
  const __half *res_hi, *res_lo;
  int res_ctot, res_coff;
  const __half *in_hi, *in_lo;   // NHWC activation planes (cp.async producers)
  int in_ctot, in_coff;
  unsigned long long *dbg;       // optional [gridDim.x][16] phase cycle counters (SHAPY_CONV_PHASES=1)
};

__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) |
         (1ull << 46);
}

// Halo kernel warp roles: warps 0..7 fill the A planes with 16-byte cp.async (zero-filled outside the image;
// a TMA box with a 16-byte inner row sustains only ~2 B/cycle/SM, measured), warp 0 lane 0 also streams the
// weights by TMA; warp 8 issues the MMAs; warps 9..16 run the epilogue.
constexpr int kHaloProd = 8;
constexpr int kHaloThreads = 32 * (kHaloProd + 1 + kEpiWarps);

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int KCH, bool SPLIT>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int PARTS = SPLIT ? 2 : 1;
  constexpr int PLANES = KCH / 8;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;                                  // 2 slices
  const uint32_t b_base = a_base + 2u * p.a_slice_bytes;              // ring of bstages x PARTS blocks
  const uint32_t b_stage_bytes = p.b_stride;                          // padded size of the [B_hi | B_lo] pair
  const uint32_t bar_base = b_base + b_stage_bytes * p.bstages;
  auto b_full = [&](int s) { return bar_base + 8u * s; };
  auto b_empty = [&](int s) { return bar_base + 8u * (p.bstages + s); };
  const uint32_t a_full0 = bar_base + 16u * p.bstages;   // [2]
  const uint32_t a_empty0 = a_full0 + 16u;               // [2]
  const uint32_t acc_full0 = a_empty0 + 16u;             // [2]
  const uint32_t acc_empty0 = acc_full0 + 16u;           // [2]
  const uint32_t tmem_slot = acc_empty0 + 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = p.n_super * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.bstages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full0 + 8u * i, 32 * kHaloProd); mbar_init(a_empty0 + 8u * i, 1);
      mbar_init(acc_full0 + 8u * i, 1); mbar_init(acc_empty0 + 8u * i, kEpiWarps);
    }
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
  pdl_launch_dependents();   // let the next launch start its prologue
  pdl_wait();                // everything above touched no global memory; inputs of this kernel are now complete

  auto item_coords = [&](int id, int &n0, int &h0, int &c_out0) {
    c_out0 = (id % p.n_tiles) * p.NT; id /= p.n_tiles;
    h0 = (id % p.bands) * p.R; id /= p.bands;
    n0 = id * p.TN;
  };
  const uint32_t acc_cols = (uint32_t)p.NT * PARTS;

  if (warp < kHaloProd) {
    // ===================================================================== producers
    const int pt = threadIdx.x;                      // 0 .. 32 * kHaloProd - 1
    const bool b_thread = pt == 0;
    int bs = 0, slice = 0;
    uint32_t bph = 0, aph = 0;
    if (b_thread && p.b_resident) {
      // all 9 x ncg weight blocks of the (single) cout tile stay in shared memory for the whole kernel
      mbar_expect_tx(b_full(0), p.b_bytes * PARTS * 9u * p.ncg);
      for (int cg = 0; cg < p.ncg; ++cg)
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t bb = b_base + b_stage_bytes * (cg * 9 + tap);
          tma_load_3d(bb, &p.b_hi, b_full(0), cg * KCH, 0, tap);
          if (SPLIT) tma_load_3d(bb + p.b_bytes, &p.b_lo, b_full(0), cg * KCH, 0, tap);
        }
    }
    const int per_img = p.Hb * p.Wp;
    const int P = p.TN * per_img;
    constexpr int PAIRS = KCH / 16;                  // 16 channels = one 32-byte sector = two planes
    for (int id = blockIdx.x; id < total_items; id += gridDim.x) {
      int n0, h0, c_out0;
      item_coords(id, n0, h0, c_out0);
      for (int cg = 0; cg < p.ncg; ++cg) {
        long long t0 = clock64();
        mbar_wait(a_empty0 + 8u * slice, aph ^ 1);
        long long t1 = clock64();
        const uint32_t sb = a_base + slice * p.a_slice_bytes;
        for (int idx = pt; idx < P * PAIRS; idx += 32 * kHaloProd) {
          const int pp = idx / P, pos = idx - pp * P;
          const int n = pos / per_img, rem = pos - n * per_img, hr = rem / p.Wp, hc = rem - hr * p.Wp;
          const int ih = h0 - 1 + hr, iw = hc - 1, in_n = n0 + n;
          const bool ok = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W && in_n < p.N;
          const size_t off = ok ? ((((size_t)in_n * p.H + ih) * p.W + iw) * p.in_ctot + p.in_coff + cg * KCH + pp * 16) : 0;
          const uint32_t nbytes = ok ? 16u : 0u;
          const uint32_t d = sb + (2 * pp) * p.PS + (uint32_t)pos * 16u;
          cp_async16(d, p.in_hi + off, nbytes);
          cp_async16(d + p.PS, p.in_hi + off + 8, nbytes);
          if (SPLIT) {
            cp_async16(d + PLANES * p.PS, p.in_lo + off, nbytes);
            cp_async16(d + (PLANES + 1) * p.PS, p.in_lo + off + 8, nbytes);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tcgen05 (async proxy) reads
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_full0 + 8u * slice) : "memory");
        if (p.dbg && pt == 32) { long long t2 = clock64(); atomicAdd(p.dbg + blockIdx.x * 16 + 0, (unsigned long long)(t1 - t0)); atomicAdd(p.dbg + blockIdx.x * 16 + 1, (unsigned long long)(t2 - t1)); }
        if (++slice == 2) { slice = 0; aph ^= 1; }
        if (b_thread && !p.b_resident) {
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
  } else if (warp == kHaloProd) {
    // ===================================================================== MMA issuer
    {
      int bs = 0, slice = 0, lt = 0;
      uint32_t bph = 0, aph = 0;
      if (p.b_resident) mbar_wait(b_full(0), 0);
      const uint64_t lbo_sbo = ((uint64_t)((p.PS >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      const uint32_t ps16 = p.PS >> 4;
      for (int id = blockIdx.x; id < total_items; id += gridDim.x, ++lt) {
        const int buf = p.acc_bufs == 2 ? (lt & 1) : 0;
        const uint32_t cph = p.acc_bufs == 2 ? ((lt >> 1) & 1) : (lt & 1);
        long long m0 = clock64();
        mbar_wait(acc_empty0 + 8u * buf, cph ^ 1);
        long long m1 = clock64();
        if (p.dbg && lane == 0) atomicAdd(p.dbg + blockIdx.x * 16 + 2, (unsigned long long)(m1 - m0));
        tc_fence_after();
        const uint32_t dbase = tmem_base + buf * 256u;
        for (int cg = 0; cg < p.ncg; ++cg) {
          long long m2 = clock64();
          mbar_wait(a_full0 + 8u * slice, aph);
          if (p.dbg && lane == 0) atomicAdd(p.dbg + blockIdx.x * 16 + 3, (unsigned long long)(clock64() - m2));
          const uint32_t sa = a_base + slice * p.a_slice_bytes;
          // one (tap) step: MT x KCH/16 k-steps x (2 | 1) MMAs
          auto issue_tap = [&](int tap, uint32_t bb) {
            const uint32_t shift = (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 16u;
            const uint64_t db0 = make_desc<KCH>(bb);
            // no-swizzle A descriptor: start address in 16-byte units in the low 14 bits
            const uint64_t da_base = lbo_sbo | (uint64_t)(((sa + shift) & 0x3FFFF) >> 4);
            const uint32_t first = (cg | tap) ? 1u : 0u;
            for (int m = 0; m < p.MT; ++m) {
              const uint32_t d0 = dbase + m * acc_cols, d1 = d0 + p.NT;
              const uint64_t dam = da_base + (uint64_t)(m * 128);   // 128 rows x 16 B = 128 descriptor units
#pragma unroll
              for (int ks = 0; ks < KCH / 16; ++ks) {
                const uint64_t da = dam + (uint64_t)(2 * ks) * ps16;
                const uint64_t ko = (uint64_t)(ks * 2);
                if (SPLIT) {
                  umma_f16(d0, da, db0 + ko, p.idesc2, first | (ks ? 1u : 0u));      // [D0 | D1] (+)= A_hi . [B_hi | B_lo]^T
                  umma_f16(d1, da + (uint64_t)PLANES * ps16, db0 + ko, p.idesc, 1);  // D1 += A_lo . B_hi^T
                } else {
                  umma_f16(d0, da, db0 + ko, p.idesc, first | (ks ? 1u : 0u));
                }
              }
            }
          };
          if (p.b_resident) {
            // weights are resident: all 9 taps of this channel group are issued by one elected lane in one go
            tc_fence_after();
            if (elect_one()) {
#pragma unroll 1
              for (int tap = 0; tap < 9; ++tap) issue_tap(tap, b_base + b_stage_bytes * (cg * 9 + tap));
              umma_commit(a_empty0 + 8u * slice);
              if (cg == p.ncg - 1) umma_commit(acc_full0 + 8u * buf);
            }
            __syncwarp();
          } else {
            // streamed weights: the elected lane also waits for each weight block, so the whole channel group
            // is issued from one uniform region (a per-tap elect + __syncwarp costs ~150 cycles per tap)
            if (elect_one()) {
              int bs_l = bs;
              uint32_t bph_l = bph;
#pragma unroll 1
              for (int tap = 0; tap < 9; ++tap) {
                mbar_wait(b_full(bs_l), bph_l);
                tc_fence_after();
                issue_tap(tap, b_base + b_stage_bytes * bs_l);
                umma_commit(b_empty(bs_l));
                if (++bs_l == p.bstages) { bs_l = 0; bph_l ^= 1; }
              }
              umma_commit(a_empty0 + 8u * slice);
              if (cg == p.ncg - 1) umma_commit(acc_full0 + 8u * buf);
            }
            __syncwarp();
            for (int tap = 0; tap < 9; ++tap) { if (++bs == p.bstages) { bs = 0; bph ^= 1; } }
          }
          if (++slice == 2) { slice = 0; aph ^= 1; }
        }
        if (p.dbg && lane == 0) atomicAdd(p.dbg + blockIdx.x * 16 + 4, (unsigned long long)(clock64() - m1));
      }
    }
  } else {
    // ===================================================================== epilogue (8 warps after the MMA warp)
    const int q = warp & 3;
    const int half = ((warp - (kHaloProd + 1)) >> 2) & 1;
    const int row = q * 32 + lane;
    const int per_img = p.Hb * p.Wp;
    int lt = 0;
    for (int id = blockIdx.x; id < total_items; id += gridDim.x, ++lt) {
      int n0, h0, c_out0;
      item_coords(id, n0, h0, c_out0);
      const int buf = p.acc_bufs == 2 ? (lt & 1) : 0;
      const uint32_t cph = p.acc_bufs == 2 ? ((lt >> 1) & 1) : (lt & 1);
      long long e0 = clock64();
      mbar_wait(acc_full0 + 8u * buf, cph);
      long long e1 = clock64();
      tc_fence_after();
      for (int m = 0; m < p.MT; ++m) {
        const int o = m * 128 + row;
        const int n = o / per_img, rem = o % per_img, rr = rem / p.Wp, cc = rem % p.Wp;
        const int oh = h0 + rr, on = n0 + n;
        const bool ok = n < p.TN && rr < p.R && cc < p.W && oh < p.H && on < p.N;
        const size_t pix = ((size_t)on * p.H + oh) * p.W + cc;
        const uint32_t lane_addr = tmem_base + buf * 256u + m * acc_cols + ((uint32_t)(q * 32) << 16);
        for (int c = half * 16; c < p.NT; c += 32) {
          const int co = c_out0 + c;
          uint4 rh0 = make_uint4(0, 0, 0, 0), rh1 = rh0, rl0 = rh0, rl1 = rh0;
          const bool has_res = ok && p.res_hi != nullptr;
          if (has_res) {
            const size_t ro = pix * p.res_ctot + p.res_coff + co;
            const uint4 *rh = reinterpret_cast<const uint4 *>(p.res_hi + ro);
            rh0 = __ldg(rh); rh1 = __ldg(rh + 1);
            if (SPLIT && p.res_lo) {
              const uint4 *rl = reinterpret_cast<const uint4 *>(p.res_lo + ro);
              rl0 = __ldg(rl); rl1 = __ldg(rl + 1);
            }
          }
          uint32_t v0[16], v1[16];
          tmem_ld16(lane_addr + c, v0);
          if (SPLIT) tmem_ld16(lane_addr + p.NT + c, v1);
          tmem_ld_wait();
          if (!ok) continue;
          float r[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float x = __uint_as_float(v0[j]);
            if (SPLIT) x += __uint_as_float(v1[j]) * kLoInv;
            r[j] = x + __ldg(p.bias + co + j);
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
          uint4 *oh4 = reinterpret_cast<uint4 *>(p.out_hi + oo);
          oh4[0] = reinterpret_cast<const uint4 *>(hi)[0];
          oh4[1] = reinterpret_cast<const uint4 *>(hi)[1];
          if (SPLIT && p.out_lo) {
            uint4 *ol4 = reinterpret_cast<uint4 *>(p.out_lo + oo);
            ol4[0] = reinterpret_cast<const uint4 *>(lo)[0];
            ol4[1] = reinterpret_cast<const uint4 *>(lo)[1];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty0 + 8u * buf) : "memory");
      if (p.dbg && threadIdx.x == 32 * (kHaloProd + 1)) { long long e2 = clock64(); atomicAdd(p.dbg + blockIdx.x * 16 + 5, (unsigned long long)(e1 - e0)); atomicAdd(p.dbg + blockIdx.x * 16 + 6, (unsigned long long)(e2 - e1)); atomicAdd(p.dbg + blockIdx.x * 16 + 7, 1ull); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kHaloProd) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
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

static bool encode(CUtensorMap *m, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
                   const cuuint32_t *box, int kch) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = kch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                    : (kch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : (kch == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void *>(base), dims,
                            strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (rank %d, dims %llu %llu %llu %llu, box %u %u %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], box[2], rank > 3 ? box[3] : 0);
    return false;
  }
  return true;
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

// Chooses the super-tile of the halo-resident kernel by a small traffic / MMA cost model.
static bool halo_configure(UmmaPlan *pl, const ConvW &w, const ActView &in, const ActView &out, const ActView *res,
                           bool relu) {
  const bool split = in.lo != nullptr;
  const int parts = split ? 2 : 1;
  const int kch = pick_kch(w.cin);
  const int H = out.H, W = out.W, N = out.N, Wp = W + 2;
  if (Wp > 256) return false;
  struct Cfg { int NT, TN, R, MT, Hb, bands, n_super, bres; double cost; } best = {0, 0, 0, 0, 0, 0, 0, 0, 1e30};
  const double sm_count = 148.0, l2_bpc = 19.6;   // measured ~5.5 TB/s of L2->SM traffic = 19.6 B/cycle/SM
  for (int NT = 128; NT >= 16; NT -= 16) {
    if (w.cout % NT) continue;
    const int mt_max = std::min(8, 512 / (NT * parts));
    if (mt_max < 1) continue;
    const int n_tiles = w.cout / NT;
    const size_t b_blk = align_up((size_t)NT * kch * 2 * parts, 1024);
    const size_t b_all = b_blk * 9 * (w.cin / kch);
    const bool bres = n_tiles == 1 && b_all <= 120 * 1024;
    auto consider = [&](int TN, int R) {
      const int Hb = R + 2;
      const int bands = ceil_div(H, R);
      const int P = TN * Hb * Wp;
      const int last = ((TN - 1) * Hb + (R - 1)) * Wp + W;   // positions that produce valid outputs
      const int MT = ceil_div(last, 128);
      if (MT > mt_max) return;
      const size_t PS = align_up((size_t)std::max(P, MT * 128 + 2 * Wp + 2) * 16, 128);
      const size_t slice = align_up((size_t)(kch / 8) * parts * PS, 1024);
      if (2 * slice + (bres ? b_all : 2 * b_blk) > 216 * 1024) return;
      const int n_super = ceil_div(N, TN) * bands;
      const double items = (double)n_super * n_tiles;
      const double a_bytes = (double)P * w.cin * 2 * parts * 2.0;                  // x2: 16-byte rows use half a sector
      const double b_bytes = bres ? 0.0 : 9.0 * NT * w.cin * 2 * parts;
      const double mem = (a_bytes + b_bytes) / l2_bpc;
      // MMA time is bound by the shared-memory operand reads (~85 B/cycle measured): 4 KB of A + the B rows
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
  if (best.NT) {
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
  p.N = N; p.H = H; p.W = W; p.Wp = Wp; p.Hb = best.Hb; p.TN = best.TN; p.R = best.R; p.bands = best.bands;
  p.n_super = best.n_super; p.MT = best.MT;
  p.cin = w.cin; p.cout = w.cout; p.NT = best.NT; p.n_tiles = w.cout / best.NT; p.ncg = w.cin / kch;
  p.acc_bufs = (best.MT * best.NT * parts <= 256) ? 2 : 1;
  const int P = best.TN * best.Hb * Wp;
  p.PS = (uint32_t)align_up((size_t)std::max(P, best.MT * 128 + 2 * Wp + 2) * 16, 128);
  p.a_slice_bytes = (uint32_t)align_up((size_t)(kch / 8) * parts * p.PS, 1024);
  p.a_tx = (uint32_t)((size_t)(kch / 8) * parts * P * 16);
  p.b_bytes = (uint32_t)best.NT * kch * 2;
  p.b_stride = (uint32_t)align_up((size_t)p.b_bytes * parts, 1024);
  const size_t b_stage = (size_t)p.b_stride;
  p.b_resident = best.bres;
  if (p.b_resident) {
    p.bstages = 9 * p.ncg;
  } else {
    const size_t avail = 214 * 1024 - 2 * (size_t)p.a_slice_bytes;
    p.bstages = (int)std::max<size_t>(2, std::min<size_t>(9, avail / b_stage));
  }
  if (2 * (size_t)p.a_slice_bytes + b_stage * p.bstages > 222 * 1024) return false;
  p.idesc = (1u << 4) | ((uint32_t)(best.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.idesc2 = (1u << 4) | ((uint32_t)((2 * best.NT) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.relu = relu;
  p.bias = w.bias;
  p.out_hi = out.hi; p.out_lo = out.lo; p.out_ctot = out.Ctot; p.out_coff = out.coff;
  p.res_hi = res ? res->hi : nullptr; p.res_lo = res ? res->lo : nullptr;
  p.res_ctot = res ? res->Ctot : 0; p.res_coff = res ? res->coff : 0;
  p.in_hi = in.hi; p.in_lo = in.lo; p.in_ctot = in.Ctot; p.in_coff = in.coff;
  p.dbg = nullptr;
  pl->smem = 2 * (size_t)p.a_slice_bytes + b_stage * p.bstages + 16 * p.bstages + 128 + 1024;
  if (getenv("SHAPY_CONV_DEBUG"))
    fprintf(stderr, "[halo] cin %d cout %d %dx%d N %d kch %d: NT %d TN %d R %d MT %d bands %d items %d PS %u slice %u bstages %d acc_bufs %d bres %d smem %zu\n",
            w.cin, w.cout, H, W, N, kch, p.NT, p.TN, p.R, p.MT, p.bands, p.n_super * p.n_tiles, p.PS,
            p.a_slice_bytes, p.bstages, p.acc_bufs, p.b_resident, pl->smem);
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    pl->grid = dim3(std::min(p.n_super * p.n_tiles, sms), 1);
  }
  cuuint64_t dims[4] = {(cuuint64_t)in.C, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)in.N};
  cuuint64_t strides[3] = {(cuuint64_t)in.Ctot * 2, (cuuint64_t)in.W * in.Ctot * 2, (cuuint64_t)in.H * in.W * in.Ctot * 2};
  cuuint32_t box[4] = {8, (cuuint32_t)Wp, (cuuint32_t)best.Hb, (cuuint32_t)best.TN};
  bool ok = encode(&p.a_hi, in.hi + in.coff, 4, dims, strides, box, 8);
  if (split) ok = ok && encode(&p.a_lo, in.lo + in.coff, 4, dims, strides, box, 8);
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
  const int kch = pick_kch(w.cin);
  pl->kch = kch;
  pl->split = split;
  if (w.ksize == 3 && w.stride == 1 && halo_enabled() && kch <= halo_max_kch() && halo_configure(pl, w, in, out, res, relu)) {
    pl->halo = true;
    return pl;
  }
  p.N = out.N; p.Ho = out.H; p.Wo = out.W;
  p.TW = std::min(out.W, 128);
  p.TH = std::max(1, std::min(out.H, 128 / p.TW));
  p.TN = (p.TW == out.W && p.TH == out.H) ? std::max(1, std::min(out.N, 128 / (p.TW * p.TH))) : 1;
  p.tiles_w = ceil_div(out.W, p.TW); p.tiles_h = ceil_div(out.H, p.TH); p.tiles_n = ceil_div(out.N, p.TN);
  p.cin = w.cin; p.cout = w.cout; p.ksize = w.ksize; p.stride = w.stride;
  {
    // cout tile: balance SM utilisation (tiles per 148 CTAs) against the per-MMA operand-read cost (see halo model)
    const double m_tiles = (double)p.tiles_w * p.tiles_h * p.tiles_n;
    auto mma_cyc = [](double n) { return std::max((4096.0 + 32.0 * n) / 85.0, n * 0.5); };
    double best = 1e30;
    int best_nt = pick_nt(w.cout);
    for (int nt = 128; nt >= 16; nt -= 16) {
      if (w.cout % nt) continue;
      const double per_k = split ? mma_cyc(2.0 * nt) + mma_cyc(nt) : mma_cyc(nt);
      const double cost = std::ceil(m_tiles * (w.cout / nt) / 148.0) * (per_k * w.ksize * w.ksize * (w.cin / 16) + 400.0 * nt / 16);
      if (cost < best) { best = cost; best_nt = nt; }
    }
    p.NT = best_nt;
  }
  p.kpt = w.cin / kch;
  p.relu = relu;
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
    pl->grid = dim3(std::min(total, sms), 1);
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

static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("SHAPY_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

template <typename Kernel, typename Params>
static int launch_pdl(Kernel kernel, dim3 grid, int threads, size_t smem, cudaStream_t st, const Params &params) {
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
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  SHAPY_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, params));
  shapy::count_launch();
  return SHAPY_OK;
}

template <int KCH, bool SPLIT>
static int launch_t(const UmmaPlan *pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SHAPY_CUDA_TRY(cudaFuncSetAttribute(conv_umma_kernel<KCH, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  return launch_pdl(conv_umma_kernel<KCH, SPLIT>, pl->grid, kThreads, pl->smem, st, pl->p);
}

template <int KCH, bool SPLIT>
static int launch_halo_t(const UmmaPlan *pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SHAPY_CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<KCH, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  static const bool phases = getenv("SHAPY_CONV_PHASES") != nullptr;
  if (phases) {
    // debug: per-CTA cycle counters of each warp role's waits (synchronous, prints to stderr)
    HaloParams hp = pl->hp;
    unsigned long long *d = nullptr;
    cudaMalloc(&d, 148 * 16 * 8);
    cudaMemset(d, 0, 148 * 16 * 8);
    hp.dbg = d;
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
    cudaFree(d);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    shapy::count_launch();
    return SHAPY_OK;
  }
  return launch_pdl(conv_halo_kernel<KCH, SPLIT>, pl->grid, kHaloThreads, pl->smem, st, pl->hp);
}

int umma_plan_launch(const UmmaPlan *pl, cudaStream_t st) {
  if (pl->halo) {
    if (pl->split) {
      if (pl->kch == 64) return launch_halo_t<64, true>(pl, st);
      if (pl->kch == 32) return launch_halo_t<32, true>(pl, st);
      return launch_halo_t<16, true>(pl, st);
    }
    if (pl->kch == 64) return launch_halo_t<64, false>(pl, st);
    if (pl->kch == 32) return launch_halo_t<32, false>(pl, st);
    return launch_halo_t<16, false>(pl, st);
  }
  if (pl->split) {
    if (pl->kch == 64) return launch_t<64, true>(pl, st);
    if (pl->kch == 32) return launch_t<32, true>(pl, st);
    return launch_t<16, true>(pl, st);
  }
  if (pl->kch == 64) return launch_t<64, false>(pl, st);
  if (pl->kch == 32) return launch_t<32, false>(pl, st);
  return launch_t<16, false>(pl, st);
}

}  // namespace shapy
