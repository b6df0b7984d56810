// Standalone microbenchmark (NOT part of the product): cycles per tcgen05.mma.cta_group::2.kind::f16 (M = 256, K = 16) as a
// function of N and of where the A operand lives (shared memory "SS" or TMEM "TS"), operands resident (no TMA in the
// loop). One cluster of two CTAs; the leader issues `iters` k-blocks of four MMAs over the same operand tiles and waits
// for one commit. Answers: does an N = 64 SS MMA run at its N / 2-cycle floor, or is it bound by reading the 4-KB A
// slice from shared memory every instruction?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/mma_rate_probe tools/mma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N, int TS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(long long* cyc, int iters, int kblocks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // operand tiles: kblocks x (A 128 x 64 fp16 = 16 KB, then B (N/2) x 64 fp16), zero filled (values do not matter)
  const uint32_t kb_bytes = 16384 + (N / 2) * 128;
  for (uint32_t i = threadIdx.x * 16; i < kblocks * kb_bytes; i += blockDim.x * 16) *(uint4*)(smem + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&holder)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = holder;
  if (rank == 0 && warp == 1) {   // warp-uniform loop, tcgen05 on one elected lane (as the product kernels issue)
    const uint32_t idesc = idesc_f16(256, N);
    const long long t0 = clock64();
    int kb = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t base = smem_u32(smem) + (uint32_t)kb * kb_bytes;
      if (++kb == kblocks) kb = 0;
      const uint64_t da = desc_sw128(base), db = desc_sw128(base + 16384);
      uint32_t pred = 0;
      asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred) : "r"(0xffffffffu));
      if (pred) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
                       "r"(tmem + 384 + 8 * k), "l"(db + 2 * k), "r"(idesc), "r"(1u) : "memory");
        } else {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da + 2 * k),
                       "l"(db + 2 * k), "r"(idesc), "r"(1u) : "memory");
        }
      }
      }
      __syncwarp();
    }
    if (lane == 0) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    uint32_t ok = 0, spins = 0;
    while (!ok && ++spins < 4000000u) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    const long long t1 = clock64();
    cyc[0] = t1 - t0;
    cyc[1] = ok;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int N, int TS>
static void run(int iters, int kblocks) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  const size_t smem = 1024 + (size_t)kblocks * (16384 + (N / 2) * 128);
  cudaFuncSetAttribute(probe<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) probe<N, TS><<<2, 128, smem>>>(d, iters, kblocks);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double per = (double)h[0] / (iters * 4.0);
  printf("N=%3d %s kblocks=%d: %8.1f cycles per MMA (floor N/2 = %d)  -> %.0f%% of the tensor peak   [%s, done=%lld]\n", N,
         TS ? "A in TMEM" : "A in smem", kblocks, per, N / 2, 100.0 * (N / 2) / per, cudaGetErrorString(e), h[1]);
  cudaFree(d);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 4000;
  run<64, 0>(iters, 1);
  run<64, 0>(iters, 6);
  run<128, 0>(iters, 6);
  run<192, 0>(iters, 6);
  run<256, 0>(iters, 4);
  run<64, 1>(iters, 6);
  run<192, 1>(iters, 6);
  run<256, 1>(iters, 4);
  return 0;
}
