// Self- and cross-attention for the ViT-ED blocks (reference: models/vision_transformer.py:56-80 Attention.forward,
// :174-200 CrossAttention.forward; SDPA default scale head_dim^-0.5, no mask, no dropout at eval).
//
// Flash-style: one CTA = 64 patch queries of one (sequence, head); K/V stream through shared memory in 64-key
// chunks; softmax statistics stay in registers and are combined with warp shuffles. The class token is never
// padded into a 16-row MMA tile: as a KEY it seeds the online-softmax state (m = q.k_cls, l = 1, O = v_cls),
// as a QUERY it is handled by a fifth warp with plain FMAs. That keeps the 64-token puzzle sequences exactly
// one MMA tile wide instead of two half-empty ones.
//
// Token rows live in the "split" layout: n_seq*n_patch patch rows, then n_seq cls rows.
#include "kernels.h"

namespace vited {

constexpr float kLog2e = 1.4426950408889634f;

template <int HD>
__global__ void __launch_bounds__(160) attn_mma_kernel(AttnArgs a) {
  constexpr int LD = HD + 8;        // padded smem row (elements): conflict-free ldmatrix (80 B / 144 B strides)
  constexpr int PIECES = HD / 8;    // 16-byte pieces per head row
  __shared__ __align__(16) bf16 Qs[64 * LD];
  __shared__ __align__(16) bf16 Ks[64 * LD];
  __shared__ __align__(16) bf16 Vs[64 * LD];
  __shared__ float qcls[HD], kcls[HD], vcls[HD];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qblocks = (a.nq_patch + 63) / 64;
  const int b = blockIdx.x / qblocks, qb = blockIdx.x % qblocks;
  const int h = blockIdx.y;
  const int q0 = qb * 64;
  const int kvb = a.kv_index ? a.kv_index[b] : b;
  const bool do_cls_q = a.q_has_cls && qb == 0;
  const float sl2 = a.scale * kLog2e;

  // ---- stage the Q tile and the class-token vectors ----
  for (int idx = tid; idx < 64 * PIECES; idx += 160) {
    const int row = idx / PIECES, pc = idx % PIECES;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (q0 + row < a.nq_patch)
      val = *reinterpret_cast<const uint4*>(a.q + ((size_t)b * a.nq_patch + q0 + row) * a.q_ld + h * HD + pc * 8);
    *reinterpret_cast<uint4*>(&Qs[row * LD + pc * 8]) = val;
  }
  if (tid < HD) {
    if (do_cls_q) qcls[tid] = __bfloat162float(a.q[((size_t)a.n_seq * a.nq_patch + b) * a.q_ld + h * HD + tid]);
    if (a.k_has_cls) {
      const size_t krow = (size_t)a.n_kv_seq * a.nk_patch + kvb;
      kcls[tid] = __bfloat162float(a.k[krow * a.k_ld + h * HD + tid]);
      vcls[tid] = __bfloat162float(a.v[krow * a.v_ld + h * HD + tid]);
    }
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, ri = lane & 7;

  // ---- per-warp state ----
  uint32_t qf[HD / 16][4];
  float o_acc[HD / 8][4];
  float m_row[2], l_row[2];
  // class-token query state (warp 4)
  float mc = -INFINITY, lc = 0.f;
  float oc[HD / 32];

  if (warp < 4) {
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const int row = warp * 16 + (mi & 1) * 8 + ri;
      const int col = ks * 16 + (mi >> 1) * 8;
      ldsm_x4(smem_u32(&Qs[row * LD + col]), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    if (a.k_has_cls) {
      // seed the online softmax with the class-token key
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int i = 0; i < HD / 4; ++i) {
        const int d = t * (HD / 4) + i;
        const float kc = kcls[d];
        s0 += __bfloat162float(Qs[(warp * 16 + g) * LD + d]) * kc;
        s1 += __bfloat162float(Qs[(warp * 16 + g + 8) * LD + d]) * kc;
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      m_row[0] = s0 * sl2; m_row[1] = s1 * sl2;
      l_row[0] = l_row[1] = (t == 0) ? 1.f : 0.f;   // thread-partial row sums; the quad is reduced at the end
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt) {
        o_acc[nt][0] = o_acc[nt][2] = vcls[nt * 8 + 2 * t];
        o_acc[nt][1] = o_acc[nt][3] = vcls[nt * 8 + 2 * t + 1];
      }
    } else {
      m_row[0] = m_row[1] = -INFINITY;
      l_row[0] = l_row[1] = 0.f;
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt) o_acc[nt][0] = o_acc[nt][1] = o_acc[nt][2] = o_acc[nt][3] = 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < HD / 32; ++i) oc[i] = 0.f;
    if (do_cls_q && a.k_has_cls) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < HD / 32; ++i) s += qcls[lane + 32 * i] * kcls[lane + 32 * i];
      s = warp_sum(s);
      mc = s * sl2;
      lc = 1.f;
#pragma unroll
      for (int i = 0; i < HD / 32; ++i) oc[i] = vcls[lane + 32 * i];
    }
  }

  const int n_chunks = (a.nk_patch + 63) / 64;
  for (int kc = 0; kc < n_chunks; ++kc) {
    const int k0 = kc * 64;
    if (kc > 0) __syncthreads();  // everyone is done with the previous chunk
    for (int idx = tid; idx < 64 * PIECES; idx += 160) {
      const int row = idx / PIECES, pc = idx % PIECES;
      uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (k0 + row < a.nk_patch) {
        const size_t grow = (size_t)kvb * a.nk_patch + k0 + row;
        kv = *reinterpret_cast<const uint4*>(a.k + grow * a.k_ld + h * HD + pc * 8);
        vv = *reinterpret_cast<const uint4*>(a.v + grow * a.v_ld + h * HD + pc * 8);
      }
      *reinterpret_cast<uint4*>(&Ks[row * LD + pc * 8]) = kv;
      *reinterpret_cast<uint4*>(&Vs[row * LD + pc * 8]) = vv;
    }
    __syncthreads();

    if (warp < 4) {
      // ---- S = Q K^T : 16 queries x 64 keys per warp ----
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int np = 0; np < 4; ++np) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const int key = np * 16 + (mi >> 1) * 8 + ri;
          const int dim = ks * 16 + (mi & 1) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4(smem_u32(&Ks[key * LD + dim]), b0, b1, b2, b3);
          mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
          mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
        }
      }
      // ---- online softmax (rows g and g+8 of this warp's tile) ----
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = k0 + nt * 8 + 2 * t;
        const bool v0 = key < a.nk_patch, v1 = key + 1 < a.nk_patch;
        s[nt][0] = v0 ? s[nt][0] * sl2 : -INFINITY;
        s[nt][1] = v1 ? s[nt][1] * sl2 : -INFINITY;
        s[nt][2] = v0 ? s[nt][2] * sl2 : -INFINITY;
        s[nt][3] = v1 ? s[nt][3] * sl2 : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m_row[0], mx0), mn1 = fmaxf(m_row[1], mx1);
      const float c0 = exp2f(m_row[0] - mn0), c1 = exp2f(m_row[1] - mn1);
      m_row[0] = mn0; m_row[1] = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - mn0); s[nt][1] = exp2f(s[nt][1] - mn0);
        s[nt][2] = exp2f(s[nt][2] - mn1); s[nt][3] = exp2f(s[nt][3] - mn1);
        rs0 += s[nt][0] + s[nt][1];
        rs1 += s[nt][2] + s[nt][3];
      }
      l_row[0] = l_row[0] * c0 + rs0;
      l_row[1] = l_row[1] * c1 + rs1;
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt) {
        o_acc[nt][0] *= c0; o_acc[nt][1] *= c0;
        o_acc[nt][2] *= c1; o_acc[nt][3] *= c1;
      }
      // ---- O += P V ----
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        uint32_t pa[4];
        pa[0] = pack_bf16(s[2 * k2][0], s[2 * k2][1]);
        pa[1] = pack_bf16(s[2 * k2][2], s[2 * k2][3]);
        pa[2] = pack_bf16(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
        pa[3] = pack_bf16(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
        for (int dp = 0; dp < HD / 16; ++dp) {
          const int key = k2 * 16 + (mi & 1) * 8 + ri;
          const int dim = dp * 16 + (mi >> 1) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4_trans(smem_u32(&Vs[key * LD + dim]), b0, b1, b2, b3);
          mma_bf16_16816(o_acc[2 * dp], pa, b0, b1);
          mma_bf16_16816(o_acc[2 * dp + 1], pa, b2, b3);
        }
      }
    } else if (do_cls_q) {
      // ---- class-token query: lane owns keys (lane, lane+32) for the scores, head dims for the output ----
      float sc[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int key = lane + 32 * j;
        float acc = 0.f;
#pragma unroll
        for (int pc = 0; pc < PIECES; ++pc) {
          const uint4 kk = *reinterpret_cast<const uint4*>(&Ks[key * LD + pc * 8]);
          const float2 k01 = unpack_bf16(kk.x), k23 = unpack_bf16(kk.y), k45 = unpack_bf16(kk.z), k67 = unpack_bf16(kk.w);
          acc += qcls[pc * 8 + 0] * k01.x + qcls[pc * 8 + 1] * k01.y + qcls[pc * 8 + 2] * k23.x +
                 qcls[pc * 8 + 3] * k23.y + qcls[pc * 8 + 4] * k45.x + qcls[pc * 8 + 5] * k45.y +
                 qcls[pc * 8 + 6] * k67.x + qcls[pc * 8 + 7] * k67.y;
        }
        sc[j] = (k0 + key < a.nk_patch) ? acc * sl2 : -INFINITY;
      }
      const float mx = warp_max(fmaxf(sc[0], sc[1]));
      const float mn = fmaxf(mc, mx);
      const float corr = exp2f(mc - mn);
      const float p0 = exp2f(sc[0] - mn), p1 = exp2f(sc[1] - mn);
      lc = lc * corr + warp_sum(p0 + p1);
      mc = mn;
#pragma unroll
      for (int i = 0; i < HD / 32; ++i) oc[i] *= corr;
      for (int key = 0; key < 64; ++key) {
        const float pk = __shfl_sync(0xffffffffu, key < 32 ? p0 : p1, key & 31);
#pragma unroll
        for (int i = 0; i < HD / 32; ++i) oc[i] += pk * __bfloat162float(Vs[key * LD + lane + 32 * i]);
      }
    }
  }

  // ---- finalize ----
  if (warp < 4) {
    float l0 = l_row[0], l1 = l_row[1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    // each warp only overwrites the Q rows it alone consumed
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
      *reinterpret_cast<uint32_t*>(&Qs[(warp * 16 + g) * LD + nt * 8 + 2 * t]) =
          pack_bf16(o_acc[nt][0] * i0, o_acc[nt][1] * i0);
      *reinterpret_cast<uint32_t*>(&Qs[(warp * 16 + g + 8) * LD + nt * 8 + 2 * t]) =
          pack_bf16(o_acc[nt][2] * i1, o_acc[nt][3] * i1);
    }
  } else if (do_cls_q) {
    const float inv = 1.f / lc;
    const size_t orow = (size_t)a.n_seq * a.nq_patch + b;
#pragma unroll
    for (int i = 0; i < HD / 32; ++i)
      a.o[orow * a.o_ld + h * HD + lane + 32 * i] = __float2bfloat16_rn(oc[i] * inv);
  }
  __syncthreads();
  for (int idx = tid; idx < 64 * PIECES; idx += 160) {
    const int row = idx / PIECES, pc = idx % PIECES;
    if (q0 + row < a.nq_patch) {
      *reinterpret_cast<uint4*>(a.o + ((size_t)b * a.nq_patch + q0 + row) * a.o_ld + h * HD + pc * 8) =
          *reinterpret_cast<const uint4*>(&Qs[row * LD + pc * 8]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// debugging reference: one thread per query token, fp32 FMAs, keys streamed through shared memory.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t tok_row(int n_seq_total, int n_patch, int has_cls, int b, int s) {
  if (has_cls) return s == 0 ? (size_t)n_seq_total * n_patch + b : (size_t)b * n_patch + (s - 1);
  return (size_t)b * n_patch + s;
}

template <int HD>
__global__ void __launch_bounds__(128) attn_simt_kernel(AttnArgs a) {
  __shared__ float Ks[32][HD];
  __shared__ float Vs[32][HD];
  const int nq = a.nq_patch + a.q_has_cls, nk = a.nk_patch + a.k_has_cls;
  const int qblocks = (nq + 127) / 128;
  const int b = blockIdx.x / qblocks, qb = blockIdx.x % qblocks;
  const int h = blockIdx.y;
  const int kvb = a.kv_index ? a.kv_index[b] : b;
  const int sq = qb * 128 + threadIdx.x;
  const bool valid = sq < nq;
  float q[HD], o[HD];
  float m = -INFINITY, l = 0.f;
  if (valid) {
    const size_t row = tok_row(a.n_seq, a.nq_patch, a.q_has_cls, b, sq);
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = __bfloat162float(a.q[row * a.q_ld + h * HD + d]);
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  for (int k0 = 0; k0 < nk; k0 += 32) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * HD; idx += 128) {
      const int kk = idx / HD, d = idx % HD;
      float kv = 0.f, vv = 0.f;
      if (k0 + kk < nk) {
        const size_t row = tok_row(a.n_kv_seq, a.nk_patch, a.k_has_cls, kvb, k0 + kk);
        kv = __bfloat162float(a.k[row * a.k_ld + h * HD + d]);
        vv = __bfloat162float(a.v[row * a.v_ld + h * HD + d]);
      }
      Ks[kk][d] = kv;
      Vs[kk][d] = vv;
    }
    __syncthreads();
    if (valid) {
      const int kmax = min(32, nk - k0);
      for (int kk = 0; kk < kmax; ++kk) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) s += q[d] * Ks[kk][d];
        s *= a.scale;
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn);
        const float p = __expf(s - mn);
        l = l * corr + p;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = o[d] * corr + p * Vs[kk][d];
        m = mn;
      }
    }
  }
  if (valid) {
    const size_t row = tok_row(a.n_seq, a.nq_patch, a.q_has_cls, b, sq);
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < HD; ++d) a.o[row * a.o_ld + h * HD + d] = __float2bfloat16_rn(o[d] * inv);
  }
}

int attention(const AttnArgs& a, int impl, cudaStream_t stream) {
  VITED_CHECK(a.head_dim == 32 || a.head_dim == 64, "attention: head_dim %d not supported (32 or 64)", a.head_dim);
  VITED_CHECK(a.q_ld % 8 == 0 && a.k_ld % 8 == 0 && a.v_ld % 8 == 0 && a.o_ld % 8 == 0,
              "attention: row strides must be multiples of 8 elements");
  VITED_CHECK(a.nq_patch > 0 && a.nk_patch > 0, "attention: empty sequences");
  if (a.n_seq == 0) return 0;
  if (impl == IMPL_REF) {
    const int nq = a.nq_patch + a.q_has_cls;
    dim3 grid((unsigned)((size_t)a.n_seq * ((nq + 127) / 128)), a.n_heads);
    if (a.head_dim == 32) attn_simt_kernel<32><<<grid, 128, 0, stream>>>(a);
    else attn_simt_kernel<64><<<grid, 128, 0, stream>>>(a);
  } else {
    dim3 grid((unsigned)((size_t)a.n_seq * ((a.nq_patch + 63) / 64)), a.n_heads);
    if (a.head_dim == 32) attn_mma_kernel<32><<<grid, 160, 0, stream>>>(a);
    else attn_mma_kernel<64><<<grid, 160, 0, stream>>>(a);
  }
  VITED_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vited
