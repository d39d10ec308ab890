// Micro-benchmark of tcgen05.mma.cta_group::2 (CTA pairs: M = 256 across two SMs, each CTA supplies its 128 rows of A and
// HALF of the N rows of B) for the small-N shapes of the C = 48 layers.  Question for round 2 (DESIGN.md work list #3):
// does halving the B reads move N = 96 / 48 MMAs from max(32 + N/4, N/2) cycles towards N/2?
//
// NOT YET RUN ON HARDWARE (written after the round-1 GPU budget was spent): compile-checked only.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/umma_pair_bench tools/umma_pair_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int PAIR>
__global__ void __launch_bounds__(128, 1) bench(int N, int n_acc, int reps, long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t *base = (uint8_t *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x) ((uint32_t *)base)[i] = 0x3c003c00u;
  uint32_t rank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) mbar_init(smem_u32(&bar), 1);
  if (threadIdx.x < 32) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) cluster_sync(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32 && rank == 0) {
    const uint64_t ad = desc_sw128(smem_u32(base)), bd = desc_sw128(smem_u32(base + 16 * 1024));
    // instruction descriptor: D = f32, A = B = f16, K-major, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256 for a pair)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    if (elect_one()) {
      long long t0 = clock64();
#pragma unroll 1
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tmem + (uint32_t)((r % n_acc) * N);
          if (PAIR)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(ad + k * 2), "l"(bd + k * 2), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(ad + k * 2), "l"(bd + k * 2), "r"(idesc), "r"(1u) : "memory");
        }
      }
      if (PAIR)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      mbar_wait(smem_u32(&bar), 0);
      cycles[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  } else if (threadIdx.x == 0 && PAIR) {
    mbar_wait(smem_u32(&bar), 0);          // the peer CTA's barrier receives the multicast arrive
    cycles[blockIdx.x] = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) cluster_sync(); else __syncthreads();
  if (threadIdx.x < 32) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  sms &= ~1;
  long long *d;
  cudaMalloc(&d, sizeof(long long) * sms);
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
  printf("%-10s %6s %6s %12s %12s\n", "mode", "N", "accs", "cyc/MMA", "per-SM ideal");
  const int reps = 1024;
  for (int pair = 0; pair < 2; ++pair)
    for (int N : {48, 96, 144, 192, 256})
      for (int n_acc : {1, 2}) {
        if (n_acc * N > 448) continue;
        for (int it = 0; it < 2; ++it) {
          if (pair) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(sms); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 50 * 1024;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, bench<1>, N, n_acc, reps, d);
          } else {
            bench<0><<<sms, 128, 50 * 1024>>>(N, n_acc, reps, d);
          }
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s N=%d: %s\n", pair ? "pair" : "single", N, cudaGetErrorString(e)); return 1; }
        long long h[512], mx = 0;
        cudaMemcpy(h, d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        for (int i = 0; i < sms; ++i) if (h[i] > mx) mx = h[i];
        // a pair MMA covers M = 256 x N: per SM it is the same 128 x N x 16 work as a single-CTA MMA
        printf("%-10s %6d %6d %12.1f %12.1f\n", pair ? "pair" : "single", N, n_acc, mx / (reps * 4.0), N / 2.0);
      }
  return 0;
}
