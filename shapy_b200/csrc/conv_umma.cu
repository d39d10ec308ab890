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
  uint32_t idesc;
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

template <int KCH, bool SPLIT>
__global__ void __launch_bounds__(192) conv_umma_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: ring of stages, each: G x { A_hi, [A_lo], B_hi, [B_lo] }; then barriers
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BLK = 128 * KCH * 2;
  const uint32_t B_BLK = p.b_stride;
  const uint32_t kblk_bytes = (A_BLK + B_BLK) * (SPLIT ? 2 : 1);
  const uint32_t stage_bytes = kblk_bytes * p.G;
  const uint32_t bar_base = smem_base + stage_bytes * p.stages;
  // barriers: full[s] at +8*s, empty[s] at +8*(stages+s), acc_full at +16*stages, tmem ptr after
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t acc_bar = bar_base + 16u * p.stages;
  const uint32_t tmem_slot = acc_bar + 8u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  int tile = blockIdx.x;
  const int tw_i = tile % p.tiles_w; tile /= p.tiles_w;
  const int th_i = tile % p.tiles_h; tile /= p.tiles_h;
  const int tn_i = tile;
  const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;
  const int c_out0 = blockIdx.y * p.NT;
  const int taps = p.ksize * p.ksize;
  const int iters = taps * p.kpt / p.G;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int pad = p.ksize / 2;
      const int per_tap = p.kpt / p.G;
      for (int it = 0; it < iters; ++it) {
        const int tap = it / per_tap, cb0 = (it % per_tap) * p.G;
        const int ky = tap / p.ksize, kx = tap % p.ksize;
        int mi = 0, x, y;
        if (p.stride == 1) {
          x = w0 + kx - pad; y = h0 + ky - pad;
        } else {  // input row 2*oh + ky - 1: ky=0 -> odd rows, index oh-1; ky=1 -> even rows, oh; ky=2 -> odd rows, oh
          const int phy = ky != 1, phx = kx != 1;
          mi = phy * 2 + phx;
          x = w0 + (kx == 0 ? -1 : 0); y = h0 + (ky == 0 ? -1 : 0);
        }
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), (p.a_bytes + p.b_bytes) * (SPLIT ? 2u : 1u) * p.G);
        const uint32_t sbase = smem_base + stage_bytes * s;
        for (int g = 0; g < p.G; ++g) {
          const uint32_t kb = sbase + kblk_bytes * g;
          const int c0 = (cb0 + g) * KCH;
          tma_load_4d(kb, &p.a_hi[mi], full_bar(s), c0, x, y, n0);
          if (SPLIT) tma_load_4d(kb + A_BLK, &p.a_lo[mi], full_bar(s), c0, x, y, n0);
          const uint32_t bb = kb + A_BLK * (SPLIT ? 2 : 1);
          tma_load_3d(bb, &p.b_hi, full_bar(s), c0, c_out0, tap);
          if (SPLIT) tma_load_3d(bb + B_BLK, &p.b_lo, full_bar(s), c0, c_out0, tap);
        }
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t d0 = tmem_base, d1 = tmem_base + p.NT;
      uint32_t acc0 = 0, acc1 = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sbase = smem_base + stage_bytes * s;
        for (int g = 0; g < p.G; ++g) {
          const uint32_t kb = sbase + kblk_bytes * g;
          const uint32_t a_hi = kb, a_lo = kb + A_BLK;
          const uint32_t b_hi = kb + A_BLK * (SPLIT ? 2 : 1), b_lo = b_hi + B_BLK;
#pragma unroll
          for (int ks = 0; ks < KCH / 16; ++ks) {
            const uint64_t da = make_desc<KCH>(a_hi + ks * 32), db = make_desc<KCH>(b_hi + ks * 32);
            umma_f16(d0, da, db, p.idesc, acc0);
            acc0 = 1;
            if (SPLIT) {
              umma_f16(d1, da, make_desc<KCH>(b_lo + ks * 32), p.idesc, acc1);
              acc1 = 1;
              umma_f16(d1, make_desc<KCH>(a_lo + ks * 32), db, p.idesc, 1);
            }
          }
        }
        umma_commit(empty_bar(s));
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(acc_bar);
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;       // tile row == TMEM lane
    const int dw = row % p.TW, dh = (row / p.TW) % p.TH, dn = row / (p.TW * p.TH);
    const int ow = w0 + dw, oh = h0 + dh, on = n0 + dn;
    const bool ok = dn < p.TN && ow < p.Wo && oh < p.Ho && on < p.N;
    const size_t pix = ((size_t)on * p.Ho + oh) * p.Wo + ow;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c = 0; c < p.NT; c += 16) {
      uint32_t v0[16], v1[16];
      tmem_ld16(lane_addr + c, v0);
      if (SPLIT) tmem_ld16(lane_addr + p.NT + c, v1);
      tmem_ld_wait();
      if (!ok) continue;
      float r[16];
      const int co = c_out0 + c;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = __uint_as_float(v0[j]);
        if (SPLIT) x += __uint_as_float(v1[j]) * kLoInv;
        r[j] = x + __ldg(p.bias + co + j);
      }
      if (p.res_hi) {
        const size_t ro = pix * p.res_ctot + p.res_coff + co;
        const uint4 *rh = reinterpret_cast<const uint4 *>(p.res_hi + ro);
        uint4 h0v = rh[0], h1v = rh[1];
        const __half *hh0 = reinterpret_cast<const __half *>(&h0v), *hh1 = reinterpret_cast<const __half *>(&h1v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { r[j] += __half2float(hh0[j]); r[8 + j] += __half2float(hh1[j]); }
        if (SPLIT && p.res_lo) {
          const uint4 *rl = reinterpret_cast<const uint4 *>(p.res_lo + ro);
          uint4 l0v = rl[0], l1v = rl[1];
          const __half *ll0 = reinterpret_cast<const __half *>(&l0v), *ll1 = reinterpret_cast<const __half *>(&l1v);
#pragma unroll
          for (int j = 0; j < 8; ++j) { r[j] += __half2float(ll0[j]) * kLoInv; r[8 + j] += __half2float(ll1[j]) * kLoInv; }
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
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
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
  CUtensorMapSwizzle sw = kch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
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
  p.N = out.N; p.Ho = out.H; p.Wo = out.W;
  p.TW = std::min(out.W, 128);
  p.TH = std::max(1, std::min(out.H, 128 / p.TW));
  p.TN = (p.TW == out.W && p.TH == out.H) ? std::max(1, std::min(out.N, 128 / (p.TW * p.TH))) : 1;
  p.tiles_w = ceil_div(out.W, p.TW); p.tiles_h = ceil_div(out.H, p.TH); p.tiles_n = ceil_div(out.N, p.TN);
  p.cin = w.cin; p.cout = w.cout; p.NT = pick_nt(w.cout); p.ksize = w.ksize; p.stride = w.stride;
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
  p.b_stride = (uint32_t)align_up(p.b_bytes, 1024);
  const uint32_t kblk = (128u * kch * 2 + p.b_stride) * (split ? 2 : 1);
  // k-blocks per stage: amortise barrier round trips for thin k-blocks, stay <= 32 KB per stage
  int G = 1;
  for (int g = p.kpt; g >= 1; --g)
    if (p.kpt % g == 0 && (size_t)g * kblk <= 32768) { G = g; break; }
  p.G = G;
  const size_t stage = (size_t)G * kblk;
  const int iters = p.ksize * p.ksize * p.kpt / G;
  int stages = (int)std::min<size_t>(8, (96 * 1024) / stage);
  stages = std::max(2, std::min(stages, std::max(2, iters)));
  p.stages = stages;
  pl->smem = stage * stages + 16 * stages + 64 + 1024;
  p.bias = w.bias;
  p.out_hi = out.hi; p.out_lo = out.lo; p.out_ctot = out.Ctot; p.out_coff = out.coff;
  p.res_hi = res ? res->hi : nullptr; p.res_lo = res ? res->lo : nullptr;
  p.res_ctot = res ? res->Ctot : 0; p.res_coff = res ? res->coff : 0;
  pl->grid = dim3(p.tiles_w * p.tiles_h * p.tiles_n, w.cout / p.NT);
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

template <int KCH, bool SPLIT>
static int launch_t(const UmmaPlan *pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SHAPY_CUDA_TRY(cudaFuncSetAttribute(conv_umma_kernel<KCH, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  conv_umma_kernel<KCH, SPLIT><<<pl->grid, 192, pl->smem, st>>>(pl->p);
  SHAPY_LAUNCH_CHECK();
  return SHAPY_OK;
}

int umma_plan_launch(const UmmaPlan *pl, cudaStream_t st) {
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
