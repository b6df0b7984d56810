// Retrieval metrics of the Hisfrag consumer on the device.
//
// The reference turns the similarity matrix into fp16 distances (hisfrag.py:283-296) and calls
// wi19_evaluate.get_metrics (misc/wi19_evaluate.py:12-56): argsort every row, drop the first sorted column ("self"),
// mark same-label items, and derive mAP / top-1 / Pr@10 / Pr@100 from cumulative sums over the N x (N-1) matrix.
// No sort is needed for that: for a query row only the RANKS of its relevant items matter, and the rank of item r is
// the number of items that sort before it. One CTA per query row, the row's sort keys in shared memory:
//   key(j) = (order-preserving 16-bit image of the fp16 distance 1 - fp16(sim[i, j])) << 16 | j
// -- ties in distance are broken by ascending index (numpy's own order among ties is implementation defined);
//   first      = argmin key                                 (the column get_metrics drops)
//   rank'(r)   = #{j : key(j) < key(r)}                     (1-based rank after dropping `first`)
//   c(r)       = 1 + #{r' relevant, r' != first : key(r') < key(r)}
//   AP sum     = sum over relevant r != first of c(r) / rank'(r)          (precision at each relevant rank)
// Integer work on bytes that are read once: N^2 * 4 bytes of similarities per evaluation.
#include <cuda_fp16.h>

#include "kernels.h"

namespace vited {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxRelevant = 2048;   // same-label items per query held in shared memory

__device__ __forceinline__ uint32_t sort_key(float sim, int j) {
  // hisfrag.py:288: scores stored as fp16; :296: distance = 1 - similarity, computed in fp16 (round to nearest even)
  const __half d = __hsub(__float2half_rn(1.0f), __float2half_rn(sim));
  uint16_t b = __half_as_ushort(d);
  if (b == 0x8000u) b = 0;                                   // -0 == +0 for the comparison
  b = (b & 0x8000u) ? (uint16_t)~b : (uint16_t)(b | 0x8000u);  // order-preserving map of the sign-magnitude encoding
  return ((uint32_t)b << 16) | (uint32_t)j;
}

__global__ void __launch_bounds__(kThreads) retrieval_rows_kernel(
    const float* __restrict__ sim, const int* __restrict__ labels, int N, int* __restrict__ n_rel,
    double* __restrict__ ap_sum, int* __restrict__ top1, int* __restrict__ hits10, int* __restrict__ hits100,
    int* __restrict__ overflow) {
  extern __shared__ uint32_t keys[];                          // [N]
  __shared__ uint32_t s_rel[kMaxRelevant];                    // keys of the relevant items
  __shared__ double s_term[kMaxRelevant];
  __shared__ uint32_t s_min[kThreads / 32];
  __shared__ int s_count;
  __shared__ uint32_t s_first;
  const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int my_label = labels[i];
  if (threadIdx.x == 0) s_count = 0;
  uint32_t mn = 0xffffffffu;
  for (int j = threadIdx.x; j < N; j += kThreads) {
    const uint32_t k = sort_key(sim[(size_t)i * N + j], j);
    keys[j] = k;
    mn = min(mn, k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if (lane == 0) s_min[warp] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t m = s_min[0];
    for (int w = 1; w < kThreads / 32; ++w) m = min(m, s_min[w]);
    s_first = m;
  }
  __syncthreads();
  const uint32_t first = s_first;
  // relevant = same label (self included, misc/wi19_evaluate.py:26-27), minus the dropped first column
  for (int j = threadIdx.x; j < N; j += kThreads) {
    if (labels[j] == my_label && keys[j] != first) {
      const int slot = atomicAdd(&s_count, 1);
      if (slot < kMaxRelevant) s_rel[slot] = keys[j];
    }
  }
  __syncthreads();
  const int R = s_count;
  if (R > kMaxRelevant) {                                     // reported by the launcher; nothing is written for this row
    if (threadIdx.x == 0) atomicExch(overflow, 1);
    return;
  }
  int h1 = 0, h10 = 0, h100 = 0;
  for (int r = warp; r < R; r += kThreads / 32) {             // one warp per relevant item
    const uint32_t kr = s_rel[r];
    int less = 0, less_rel = 0;
    for (int j = lane; j < N; j += 32) less += keys[j] < kr;
    for (int q = lane; q < R; q += 32) less_rel += s_rel[q] < kr;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      less += __shfl_xor_sync(0xffffffffu, less, o);
      less_rel += __shfl_xor_sync(0xffffffffu, less_rel, o);
    }
    // `less` counts the dropped first column too, so it IS the 1-based rank among the remaining N - 1 columns
    if (lane == 0) {
      s_term[less_rel] = (double)(less_rel + 1) / (double)less;   // slot = rank among the relevant items: fixed order
      h1 += less == 1;
      h10 += less <= 10;
      h100 += less <= 100;
    }
  }
  __shared__ int s_h[3];
  if (threadIdx.x < 3) s_h[threadIdx.x] = 0;
  __syncthreads();
  if (lane == 0) { atomicAdd(&s_h[0], h1); atomicAdd(&s_h[1], h10); atomicAdd(&s_h[2], h100); }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int r = 0; r < R; ++r) s += s_term[r];              // ascending rank, as a cumulative sum would
    n_rel[i] = R;
    ap_sum[i] = s;
    top1[i] = s_h[0];
    hits10[i] = s_h[1];
    hits100[i] = s_h[2];
  }
}

}  // namespace

int retrieval_rows(const float* sim, const int* labels, int N, int* n_rel, double* ap_sum, int* top1, int* hits10,
                   int* hits100, cudaStream_t stream) {
  VITED_CHECK(sim && labels && n_rel && ap_sum && top1 && hits10 && hits100, "retrieval_rows: null pointer");
  VITED_CHECK(N >= 2 && N <= 49152, "retrieval_rows: N=%d out of range (2..49152: one row of keys lives in shared memory)", N);
  const size_t smem = (size_t)N * sizeof(uint32_t);
  static PerDeviceOnce once;   // the opt-in is raised once per device to the largest row the kernel accepts
  if (once.first())
    VITED_CUDA_OK(cudaFuncSetAttribute(retrieval_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 * 4));
  int* d_overflow = nullptr;
  VITED_CUDA_OK(cudaMallocAsync(&d_overflow, sizeof(int), stream));
  int overflow = 0;
  // every path below frees d_overflow (stream ordered) before it returns
  cudaError_t ce = cudaMemsetAsync(d_overflow, 0, sizeof(int), stream);
  if (ce == cudaSuccess) {
    retrieval_rows_kernel<<<N, kThreads, smem, stream>>>(sim, labels, N, n_rel, ap_sum, top1, hits10, hits100, d_overflow);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&overflow, d_overflow, sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);
  cudaFreeAsync(d_overflow, stream);
  VITED_CHECK(ce == cudaSuccess, "retrieval_rows: %s", cudaGetErrorString(ce));
  VITED_CHECK(overflow == 0, "retrieval_rows: a query has more than %d same-label items", kMaxRelevant);
  return 0;
}

}  // namespace vited
