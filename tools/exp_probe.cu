// Standalone microbenchmark (NOT part of the product): cycles for the softmax exp phase of one 64-column half row
// (64 x {FFMA, MUFU.EX2, FADD} + 32 x F2FP per thread) as a function of resident warps. One CTA.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
template <int VARIANT>
__global__ void probe(const float* in, uint32_t* out, long long* cyc, int iters, float sl2) {
  float v[64];
  for (int i = 0; i < 64; ++i) v[i] = in[(threadIdx.x * 64 + i) & 4095];
  float l = 0.f, m = in[threadIdx.x & 4095];
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float mneg = -m;
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t pk[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float p0, p1;
      if (VARIANT == 0) {
        p0 = ex2_ftz(fmaf(v[2 * j], sl2, mneg));
        p1 = ex2_ftz(fmaf(v[2 * j + 1], sl2, mneg));
      } else {
        // variant 1: every 4th pair through a degree-3 polynomial on the FMA pipe (Cody-Waite), rest on MUFU
        const float x0 = fmaf(v[2 * j], sl2, mneg), x1 = fmaf(v[2 * j + 1], sl2, mneg);
        if ((j & 3) == 3) {
          const float xc0 = fmaxf(x0, -126.f), xc1 = fmaxf(x1, -126.f);
          const float r0 = xc0 + 12582912.f, r1 = xc1 + 12582912.f;   // round to nearest integer (magic number)
          const float f0 = xc0 - (r0 - 12582912.f), f1 = xc1 - (r1 - 12582912.f);
          float q0 = fmaf(f0, 0.0555041f, 0.2402265f), q1 = fmaf(f1, 0.0555041f, 0.2402265f);
          q0 = fmaf(q0, f0, 0.6931472f); q1 = fmaf(q1, f1, 0.6931472f);
          q0 = fmaf(q0, f0, 1.0f); q1 = fmaf(q1, f1, 1.0f);
          p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(r0) << 23));
          p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(r1) << 23));
        } else {
          p0 = ex2_ftz(x0);
          p1 = ex2_ftz(x1);
        }
      }
      sum[(2 * j) & 3] += p0; sum[(2 * j + 1) & 3] += p1;
      pk[j] = pack_bf16(p0, p1);
    }
    l += (sum[0] + sum[1]) + (sum[2] + sum[3]);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= pk[j];
    m += 1e-6f * l;   // loop-carried dependency so iterations are not merged
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] += 1e-7f;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = acc + __float_as_uint(l);
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int VARIANT>
void run(const char* name) {
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4); cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  const int iters = 1000;
  for (int warps : {4, 8, 16}) {
    probe<VARIANT><<<1, warps * 32>>>(in, out, cyc, iters, 0.18f);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-8s warps=%2d (%d per SMSP): %7.1f cycles per exp phase  -> %6.1f per warp-phase per SMSP  (%s)\n", name, warps, warps / 4,
           (double)c / iters, (double)c / iters / (warps / 4), cudaGetErrorString(cudaGetLastError()));
  }
}
int main() { run<0>("mufu"); run<1>("mufu+poly"); return 0; }
