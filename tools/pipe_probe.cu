// Standalone microbenchmark (NOT part of the product): per-SM throughput of MUFU.EX2, FFMA, F2FP, LDTM.x32, STTM.x32
// as a function of resident warps, measured with clock64 on one CTA. build: nvcc -gencode arch=compute_100a,code=sm_100a -O2
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&holder)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = holder + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32;
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  uint32_t r[32];
  for (int i = 0; i < 32; ++i) r[i] = i + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) { uint32_t o; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(x[i]), "f"(x[i + 1])); x[i] = __uint_as_float(o); }
    } else if (MODE == 6) {   // ex2 on packed fp16 pairs: two exponentials per MUFU instruction?
#pragma unroll
      for (int i = 0; i < 16; ++i) { uint32_t u = __float_as_uint(x[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
    } else if (MODE == 7) {   // HADD2
#pragma unroll
      for (int i = 0; i < 16; ++i) { uint32_t u = __float_as_uint(x[i]); asm volatile("add.rn.f16x2 %0, %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
    } else if (MODE == 3) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(tb) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else if (MODE == 4) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                   ::"r"(tb), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else if (MODE == 5) {   // 4 LDTM in flight before the wait
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(tb + 0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 16; ++i) acc += x[i];
  for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) *cyc = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(holder) : "memory");
}

template <int MODE>
void run(const char* name, int per_iter_ops, int bytes_per_op) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    probe<MODE><<<1, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_ops = (double)iters * per_iter_ops * warps;
    printf("%-10s warps=%2d  cycles=%9lld  clk per warp-instr per SMSP = %6.2f", name, warps, c, c / (warp_ops / 4));
    if (bytes_per_op) printf("   bytes/clk/SM = %7.1f", warp_ops * bytes_per_op / c);
    printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
  }
}
int main() {
  run<0>("MUFU.EX2", 16, 0);
  run<6>("EX2.F16x2", 16, 0);
  run<7>("HADD2", 16, 0);
  run<1>("FFMA", 16, 0);
  run<2>("F2FP", 8, 0);
  run<3>("LDTM.x32", 1, 4096);
  run<5>("LDTMx4", 4, 4096);
  run<4>("STTM.x32", 1, 4096);
  return 0;
}
