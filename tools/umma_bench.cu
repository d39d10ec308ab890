// Micro-benchmark of tcgen05.mma issue / completion rates on sm_100a (one CTA per SM, one issuing thread).
// Answers the design questions behind shapy_b200/csrc/conv_umma.cu:
//   * cycles per MMA as a function of N when consecutive MMAs accumulate into the SAME TMEM tile,
//   * the same with 2 / 3 / 4 independent accumulators interleaved,
//   * M = 64 versus M = 128, and A read from TMEM instead of shared memory.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/umma_bench tools/umma_bench.cu
// run  : gpurun_out/umma_bench            (prints one line per pattern)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

struct Pattern {
  int n_ops;
  int dcol[16];   // TMEM column of the accumulator
  int N[16];      // MMA N
  int M[16];      // 64 or 128
  int a_tmem[16]; // 1: A operand from TMEM
  int reps;       // pattern repetitions
  int a_mode;     // 0 SW128, 1 SW64, 2 SW32 (K-major rows of 128/64/32 B), 3 no swizzle (8-channel planes)
  int a_shift;    // byte offset added to the A start address (row / tap shift)
  int b_mode;     // 0 SW128, 1 SW64, 2 SW32
  char name[96];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int NOPS, bool ATMEM>
__global__ void __launch_bounds__(128, 1) bench_kernel(const Pattern pat, long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  // A: 128 rows x 64 fp16 (16 KB, SW128); B: 256 rows x 64 fp16 (32 KB)
  uint8_t *base = (uint8_t *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x) ((uint32_t *)base)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) mbar_init(smem_u32(&bar), 1);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x < 32) {
    const uint32_t a_addr = smem_u32(base), b_addr = smem_u32(base + 16 * 1024);
    uint32_t idesc[NOPS], dst[NOPS];
#pragma unroll
    for (int i = 0; i < NOPS; ++i) {
      idesc[i] = (1u << 4) | ((uint32_t)(pat.N[i] >> 3) << 17) | ((uint32_t)(pat.M[i] >> 4) << 24);
      dst[i] = tmem + (uint32_t)pat.dcol[i];
    }
    uint64_t ad;
    {
      const uint32_t sa = a_addr + (uint32_t)pat.a_shift;
      const uint64_t lo = (uint64_t)((sa & 0x3FFFF) >> 4);
      if (pat.a_mode == 0) ad = lo | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      else if (pat.a_mode == 1) ad = lo | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
      else if (pat.a_mode == 2) ad = lo | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
      else ad = lo | ((uint64_t)(2304 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);  // LBO = plane stride 2304 B
    }
    uint64_t bd;
    {
      const uint64_t lo = (uint64_t)((b_addr & 0x3FFFF) >> 4);
      if (pat.b_mode == 0) bd = lo | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      else if (pat.b_mode == 1) bd = lo | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
      else bd = lo | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
    }
    const uint32_t bmask = pat.b_mode == 0 ? 3 : (pat.b_mode == 1 ? 1 : 0);
    const uint32_t kadv = pat.a_mode == 3 ? (2 * 2304) >> 4 : 2;   // descriptor advance per k-step of 16
    const int reps = pat.reps;
    long long t0 = 0, t_issue = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
#pragma unroll 1
      for (int r = 0; r < reps; r += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int i = 0; i < NOPS; ++i) {
            if (ATMEM) umma_ts(dst[i], tmem + 448 + k * 8, bd + (k & bmask) * 2, idesc[i], 1);
            else umma_ss(dst[i], ad + (pat.a_mode == 2 ? 0 : (pat.a_mode == 1 ? (k & 1) : k)) * kadv, bd + (k & bmask) * 2, idesc[i], 1);
          }
        }
      }
      t_issue = clock64();
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      t1 = clock64();
      cycles[blockIdx.x * 2] = t1 - t0;
      cycles[blockIdx.x * 2 + 1] = t_issue - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int NOPS>
static void launch(const Pattern &p, int sms, long long *d) {
  if (p.a_tmem[0]) {
    cudaFuncSetAttribute(bench_kernel<NOPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
    bench_kernel<NOPS, true><<<sms, 128, 50 * 1024>>>(p, d);
  } else {
    cudaFuncSetAttribute(bench_kernel<NOPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
    bench_kernel<NOPS, false><<<sms, 128, 50 * 1024>>>(p, d);
  }
}
static void launch_any(const Pattern &p, int sms, long long *d) {
  switch (p.n_ops) {
    case 1: launch<1>(p, sms, d); break;
    case 2: launch<2>(p, sms, d); break;
    case 3: launch<3>(p, sms, d); break;
    case 4: launch<4>(p, sms, d); break;
    case 6: launch<6>(p, sms, d); break;
    case 8: launch<8>(p, sms, d); break;
    default: printf("unsupported n_ops %d\n", p.n_ops); exit(1);
  }
}

static std::vector<Pattern> pats;
static void add(const char *name, std::vector<int> dcol, std::vector<int> N, int M = 128, int a_tmem = 0, int a_mode = 0,
                int a_shift = 0, int b_mode = 0) {
  Pattern p;
  memset(&p, 0, sizeof(p));
  p.n_ops = (int)dcol.size();
  for (int i = 0; i < p.n_ops; ++i) { p.dcol[i] = dcol[i]; p.N[i] = N[i]; p.M[i] = M; p.a_tmem[i] = a_tmem; }
  p.reps = 4096 / p.n_ops / 4 * 4;
  p.a_mode = a_mode; p.a_shift = a_shift; p.b_mode = b_mode;
  snprintf(p.name, sizeof(p.name), "%s", name);
  pats.push_back(p);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  char nm[96];
  for (int N : {48, 96, 192}) {
    for (int bm = 0; bm < 3; ++bm) {
      const char *bn = bm == 0 ? "SW128" : (bm == 1 ? "SW64" : "SW32");
      snprintf(nm, sizeof nm, "A noswz B %-5s N=%d", bn, N); add(nm, {0}, {N}, 128, 0, 3, 16, bm);
      snprintf(nm, sizeof nm, "A SW128 B %-5s N=%d", bn, N); add(nm, {0}, {N}, 128, 0, 0, 0, bm);
      snprintf(nm, sizeof nm, "A tmem  B %-5s N=%d", bn, N); add(nm, {0}, {N}, 128, 1, 0, 0, bm);
    }
  }
  long long *d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * 2 * sms);
  std::vector<long long> h(2 * sms);
  printf("%-28s %10s %10s %10s %8s\n", "pattern", "cyc/MMA", "issue/MMA", "ideal", "eff");
  for (auto &p : pats) {
    for (int it = 0; it < 2; ++it) launch_any(p, sms, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", p.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d_cycles, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
    long long mx = 0, is = 0;
    for (int i = 0; i < sms; ++i) { if (h[2 * i] > mx) mx = h[2 * i]; if (h[2 * i + 1] > is) is = h[2 * i + 1]; }
    double n = (double)p.reps * p.n_ops, ideal = 0;
    for (int i = 0; i < p.n_ops; ++i) ideal += p.M[i] * p.N[i] * 16.0 * 2 / 8192.0;
    ideal /= p.n_ops;
    printf("%-28s %10.1f %10.1f %10.1f %7.0f%%\n", p.name, mx / n, is / n, ideal, 100.0 * ideal / (mx / n));
  }
  return 0;
}
