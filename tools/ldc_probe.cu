// Standalone microbenchmark (NOT part of the product): what does a warp-uniform parameter load cost on the shared-memory
// LSU (LDS.128 of one address for all lanes) against the constant bank (by-value kernel parameters, LDC.64 with a
// register index), alone and next to the conflict-free per-lane LDS.128 / STS.128 traffic of the full-row epilogues?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ldc_probe tools/ldc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
struct __align__(16) Params { float2 v[1024]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds_f4(uint32_t a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) { asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

// MODE bit 0: 8 per-lane LDS.128 + 8 per-lane STS.128 (the residual rows), bit 1: 8 broadcast LDS.128 (parameters from
// shared memory), bit 2: 16 LDC.64 (the same parameters from the constant bank)
template <int MODE>
__global__ void probe(const __grid_constant__ Params p, float* out, long long* cyc, int iters) {
  extern __shared__ uint8_t sm[];
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t row = smem_u32(sm) + warp * 4096 + lane * 128;
  const uint32_t par = smem_u32(sm) + 64 * 1024;
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = i * 0.001f;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<float*>(sm + 64 * 1024)[i] = i;
  __syncthreads();
  float acc[4] = {0, 0, 0, 0};
  int idx = warp;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    idx = (idx * 5 + 3) & 511;          // warp-uniform, loop carried
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = make_float4(acc[0], acc[1], acc[2], acc[3]);
      if (MODE & 1) {
        const float4 x = lds_f4(row + ((i ^ (lane & 7)) << 4));
        v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
      }
      if (MODE & 2) {
        const float4 b = lds_f4(par + (uint32_t)(idx + i) * 16);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      if (MODE & 4) {
        const float2 b0 = p.v[idx + 2 * i], b1 = p.v[idx + 2 * i + 1];
        v.x += b0.x; v.y += b0.y; v.z += b1.x; v.w += b1.y;
      }
      if (MODE & 1) sts_f4(row + ((i ^ (lane & 7)) << 4), v);
      acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[3] = v.w;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  out[threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  static Params hp;
  for (int i = 0; i < 1024; ++i) hp.v[i] = make_float2((float)i, (float)-i);
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int iters = 4000;
  for (int warps : {8, 16}) {
    probe<MODE><<<1, warps * 32, 80 * 1024>>>(hp, out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-46s warps=%2d  cycles per iteration (8 x 4 columns per warp) = %7.1f  (%s)\n", name, warps, (double)c / iters,
           cudaGetErrorString(cudaGetLastError()));
  }
}
int main() {
  run<1>("per-lane LDS.128 + STS.128 (8 + 8)");
  run<2>("broadcast LDS.128 (8)");
  run<4>("LDC.64 register-indexed (16)");
  run<3>("per-lane LDS/STS + broadcast LDS.128");
  run<5>("per-lane LDS/STS + LDC.64");
  return 0;
}
