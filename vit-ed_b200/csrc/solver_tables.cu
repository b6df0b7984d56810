// Distance tables of the Paikin-Tal solver from the [N, N, 4] score matrix, on the device.
//
// The reference fills these tables through 4*N*(N-1) Python callbacks into the closure of evaluation.py:116-131 and
// three more O(N^2) Python loops (paikin_tal_solver/inter_piece_distance.py:189-240 distances + minimum / second best,
// :325-372 asymmetric compatibility, :489-524 mutual compatibility, :626-648 best buddies) -- minutes for a 1000-piece
// puzzle once scoring itself takes seconds. Here: two launches over tables that stay L2-resident (N = 1000: 48 MB),
// integer / fp32 / fp64 arithmetic ordered exactly as the Python expressions so every table is bit-identical.
//
//   row (i, s): piece at list position i, side s (PuzzlePieceSide: top 0, right 1, bottom 2, left 3); the neighbour's
//   side is the complementary one (type-1 puzzle, :816-819), so a row has one entry per other piece j.
//   score bin of side s: (s + 3) & 3   (evaluation.py:118-129: right->0, bottom->1, left->2, top->3)
#include "kernels.h"

namespace vited {
namespace {

constexpr long long kMaxSize = 0x7fffffffffffffffLL;   // sys.maxsize (inter_piece_distance.py:283-288 initial values)
constexpr int kRowThreads = 256;

// the two smallest values of a multiset, duplicates counted
struct Two { unsigned long long a, b; };
__device__ __forceinline__ Two two_merge(Two x, Two y) {
  Two r;
  r.a = min(x.a, y.a);
  r.b = min(max(x.a, y.a), min(x.b, y.b));
  return r;
}
__device__ __forceinline__ Two two_push(Two x, unsigned long long v) {
  if (v < x.a) { x.b = x.a; x.a = v; }
  else if (v < x.b) x.b = v;
  return x;
}

// 1 - sigmoid(logit) exactly as torch computes it on the device for evaluation.py:109-114: 1 / (1 + exp(-x)) in fp32
// with IEEE division, then the fp32 subtraction. (No fast-math flags in this build.)
__device__ __forceinline__ float one_minus_sigmoid(float x) { return 1.0f - 1.0f / (1.0f + expf(-x)); }

// One CTA per row (i, s). Pass 1: distances (uint32 truncation of dist * 1000, :229) + the two smallest; pass 2:
// asymmetric compatibility in fp64 rounded to fp32 (:354-360), the number of pieces at the minimum and the lowest such j.
__global__ void __launch_bounds__(kRowThreads) tables_rows_kernel(
    const float* __restrict__ scores, int flags, const int* __restrict__ order, int N,
    uint32_t* __restrict__ asym, long long* __restrict__ min_d, long long* __restrict__ second_d,
    int* __restrict__ n_cand, int* __restrict__ cand, float* __restrict__ compat) {
  const bool scores_are_logits = (flags & 1) != 0, f32_product = (flags & 2) != 0;
  const int row = blockIdx.x, i = row >> 2, s = row & 3, bin = (s + 3) & 3;
  const int oi = order != nullptr ? order[i] : i;
  const size_t base = (size_t)row * N;
  __shared__ Two s_two[kRowThreads / 32];
  __shared__ int s_cnt[kRowThreads / 32], s_first[kRowThreads / 32];
  __shared__ Two s_row;

  // the initial (min, second) = (maxsize - 1, maxsize) take part like two more elements (:283-288, :256-272)
  Two t = {(unsigned long long)kMaxSize, (unsigned long long)kMaxSize};
  if (threadIdx.x == 0) t.a = (unsigned long long)(kMaxSize - 1);
  for (int j = threadIdx.x; j < N; j += kRowThreads) {
    uint32_t dist = 0x7fffffffu;                                   // fill value of the diagonal (:204-207)
    if (j != i) {
      const int oj = order != nullptr ? order[j] : j;
      float v = scores[((size_t)oi * N + oj) * 4 + bin];
      if (scores_are_logits) v = one_minus_sigmoid(v);
      // evaluation.py:118-129 `pred[k] * 1000.`: np.float32 scalar times a Python float. Under the NumPy 1.x the
      // reference's pinned stack runs on (torch~=2.1, scipy~=1.9.1) that promotes to float64; NumPy >= 2 (NEP 50)
      // keeps float32. The products differ in the last place and the truncating uint32 store (:229) then differs by
      // one whenever the fp32 rounding crosses an integer (~1e-5 of the entries). Default: the float64 product.
      dist = f32_product ? (uint32_t)__fmul_rn(v, 1000.0f) : (uint32_t)__dmul_rn((double)v, 1000.0);
      t = two_push(t, dist);
    }
    asym[base + j] = dist;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Two u;
    u.a = __shfl_xor_sync(0xffffffffu, t.a, o);
    u.b = __shfl_xor_sync(0xffffffffu, t.b, o);
    t = two_merge(t, u);
  }
  if ((threadIdx.x & 31) == 0) s_two[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    Two r = s_two[0];
    for (int w = 1; w < kRowThreads / 32; ++w) r = two_merge(r, s_two[w]);
    s_row = r;
    min_d[row] = (long long)r.a;
    second_d[row] = (long long)r.b;
  }
  __syncthreads();
  const unsigned long long mn = s_row.a, sec = s_row.b;

  int cnt = 0, first = 0x7fffffff;
  for (int j = threadIdx.x; j < N; j += kRowThreads) {
    float c = __int_as_float(0x7f800000);                          // +inf on the diagonal (:340-343)
    if (j != i) {
      const uint32_t dist = asym[base + j];
      if (dist == 0u) c = 1.0f;                                    // :354-355
      else if (sec == 0ull) c = (float)(-kMaxSize);                // :356-357
      else c = (float)(1.0 - (double)dist / (double)sec);          // :359-360, float64 then the float32 store
      if ((unsigned long long)dist == mn) { ++cnt; first = min(first, j); }
    }
    compat[base + j] = c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  }
  if ((threadIdx.x & 31) == 0) { s_cnt[threadIdx.x >> 5] = cnt; s_first[threadIdx.x >> 5] = first; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0, f = 0x7fffffff;
    for (int w = 0; w < kRowThreads / 32; ++w) { c += s_cnt[w]; f = min(f, s_first[w]); }
    n_cand[row] = c;
    cand[row] = c > 0 ? f : -1;
  }
}

// mutual[i, s, j] = (compat[i, s, j] + compat[j, s^2, i]) / 2 in fp32 (:507-516); best buddy of (i, s): the single
// piece at the minimum whose own single minimum on the complementary side is i (:626-648 with
// _ALLOW_MULTIPLE_BEST_BUDDIES = False, :76-84), else -1.
__global__ void __launch_bounds__(kRowThreads) tables_mutual_kernel(
    const float* __restrict__ compat, const int* __restrict__ n_cand, const int* __restrict__ cand, int N,
    float* __restrict__ mutual, int* __restrict__ best_buddy) {
  const int row = blockIdx.x, i = row >> 2, s = row & 3, cs = s ^ 2;
  const size_t base = (size_t)row * N;
  for (int j = threadIdx.x; j < N; j += kRowThreads) {
    float m = __int_as_float(0x7f800000);
    if (j != i) m = __fmul_rn(__fadd_rn(compat[base + j], compat[((size_t)j * 4 + cs) * N + i]), 0.5f);
    mutual[base + j] = m;
  }
  if (threadIdx.x == 0) {
    int bb = -1;
    if (n_cand[row] == 1) {
      const int j = cand[row];
      if (n_cand[j * 4 + cs] == 1 && cand[j * 4 + cs] == i) bb = j;
    }
    best_buddy[row] = bb;
  }
}

}  // namespace

int puzzle_tables(const float* scores, int flags, const int* order, int N, uint32_t* asym, long long* min_d,
                  long long* second_d, int* n_cand, int* cand, float* compat, float* mutual, int* best_buddy,
                  cudaStream_t stream) {
  VITED_CHECK(N >= 1 && N <= (1 << 20), "puzzle_tables: N=%d out of range", N);
  VITED_CHECK(scores && asym && min_d && second_d && n_cand && cand && compat && mutual && best_buddy,
              "puzzle_tables: null pointer");
  VITED_CHECK((flags & ~3) == 0, "puzzle_tables: unknown flag bits 0x%x", flags);
  tables_rows_kernel<<<4 * N, kRowThreads, 0, stream>>>(scores, flags, order, N, asym, min_d, second_d,
                                                        n_cand, cand, compat);
  VITED_CUDA_OK(cudaGetLastError());
  tables_mutual_kernel<<<4 * N, kRowThreads, 0, stream>>>(compat, n_cand, cand, N, mutual, best_buddy);
  VITED_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vited
