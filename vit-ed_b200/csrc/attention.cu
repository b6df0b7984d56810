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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros, nothing read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 2^x on the MUFU unit, flush-to-zero (inputs are <= 0 or -inf here); plain exp2f() adds a denormal-range fix-up
// (3-4 extra instructions per element) that this instruction-bound kernel cannot afford.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD>
struct AttnSmem {
  static constexpr int LD = HD + 8;              // padded smem row (elements): conflict-free ldmatrix (80 B / 144 B)
  static constexpr int ROWS = 80;                // 64 patch rows + the class-token row (64) + 15 zero rows
  static constexpr int T = ROWS * LD;            // one tile (elements)
  static constexpr int STAGE = 3 * T;            // Q, K, V tiles
  static constexpr int NST = HD == 32 ? 3 : 2;   // cp.async ring depth
  static constexpr int BYTES = (NST * STAGE + T) * 2;  // ring + output staging
};

// Persistent, software-pipelined flash attention for one (sequence, head, 64-query block) per step and 64-key chunk:
// the loads of step s+NST-1 are in flight (cp.async ring) while step s is computed.
//  * Five identical MMA warps: warps 0-3 own 16 patch queries each; warp 4 owns a 16-row tile whose row 0 is the
//    class-token query (rows 1-15 zero), so the class token rides the same tensor-core path.
//  * The class-token KEY/VALUE sit in row 64 of the K/V tiles and are consumed as a ninth 8-key MMA tile (first
//    chunk only), so 65-token sequences cost 9/8 of a 64-token one instead of 2x.
//  * cls_only: only warp 4 computes (last decoder layer: only row 0 of each sequence reaches the head).
// The kernel is instruction-issue bound (64x64x32 tiles), so per-step address arithmetic is hoisted out of the loop.
template <int HD>
__global__ void __launch_bounds__(160) attn_mma_kernel(AttnArgs a, int n_items, int cls_only) {
  using SM = AttnSmem<HD>;
  constexpr int LD = SM::LD;
  constexpr int NST = SM::NST;
  constexpr int PIECES = HD / 8;                      // 16-byte pieces per head row
  constexpr int NLD = (64 * PIECES + 159) / 160;      // tile pieces per thread
  extern __shared__ __align__(16) uint8_t attn_smem_raw[];
  bf16* smem = reinterpret_cast<bf16*>(attn_smem_raw);
  bf16* sO = smem + NST * SM::STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qblocks = cls_only ? 1 : (a.nq_patch + 63) / 64;
  const int n_chunks = (a.nk_patch + 63) / 64;
  const bool ragged_k = (a.nk_patch & 63) != 0;
  const float sl2 = a.scale * kLog2e;
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, ri = lane & 7;

  // rows 65..79 of every tile are never loaded: zero them once (they multiply into the MMAs as exact zeros)
  for (int i = tid; i < NST * 3 * 15 * LD; i += 160) {
    const int tile = i / (15 * LD), r = i % (15 * LD);
    smem[tile * SM::T + 65 * LD + r] = __float2bfloat16(0.f);
  }

  // per-thread load slots: (row, 16-byte piece) pairs are the same every step
  int row_of[NLD], soff[NLD], qoff[NLD], koff[NLD], voff[NLD];
#pragma unroll
  for (int i = 0; i < NLD; ++i) {
    const int idx = tid + i * 160;
    const int row = idx / PIECES, pc = idx % PIECES;
    row_of[i] = idx < 64 * PIECES ? row : 1 << 20;    // out-of-range slots never pass the row test
    soff[i] = row * LD + pc * 8;
    qoff[i] = row * a.q_ld + pc * 8;
    koff[i] = row * a.k_ld + pc * 8;
    voff[i] = row * a.v_ld + pc * 8;
  }
  const bool full_tiles = (a.nq_patch & 63) == 0 && !ragged_k;   // no partial query / key tile anywhere

  // work-item cursor in mixed radix (qb, h, b): stepping by gridDim.x needs no division
  struct Cursor { int item, qb, h, b; };
  const int H = a.n_heads;
  const int step_qb = (int)(gridDim.x % qblocks);
  const int step_h = (int)((gridDim.x / qblocks) % H);
  const int step_b = (int)(gridDim.x / qblocks / H);
  auto advance = [&](Cursor& c) {
    c.item += gridDim.x;
    c.qb += step_qb;
    if (c.qb >= qblocks) { c.qb -= qblocks; c.h += 1; }
    c.h += step_h;
    if (c.h >= H) { c.h -= H; c.b += 1; }
    c.b += step_b;
  };
  Cursor c0;
  c0.item = blockIdx.x;
  c0.qb = (int)(blockIdx.x % qblocks);
  c0.h = (int)((blockIdx.x / qblocks) % H);
  c0.b = (int)(blockIdx.x / qblocks / H);

  auto issue_loads = [&](const Cursor& c, int kc, bf16* st) {
    const int kvb = a.kv_index ? __ldg(a.kv_index + c.b) : c.b;
    bf16* Qs = st;
    bf16* Ks = st + SM::T;
    bf16* Vs = st + 2 * SM::T;
    if (kc == 0) {
      if (!cls_only) {
        const int q0 = c.qb * 64;
        const bf16* qbase = a.q + ((size_t)c.b * a.nq_patch + q0) * a.q_ld + c.h * HD;
        if (full_tiles) {
#pragma unroll
          for (int i = 0; i < NLD; ++i)
            if (row_of[i] < 64) cp_async16(Qs + soff[i], qbase + qoff[i], true);
        } else {
          const int qrows = a.nq_patch - q0;
#pragma unroll
          for (int i = 0; i < NLD; ++i) {
            if (row_of[i] < 64) {
              const bool ok = row_of[i] < qrows;
              cp_async16(Qs + soff[i], ok ? qbase + qoff[i] : a.q, ok);
            }
          }
        }
      }
      if (tid < 3 * PIECES) {
        const int which = tid / PIECES, pc = tid % PIECES;
        if (which == 0) {   // class-token query -> row 64 of the Q tile
          const bool ok = a.q_has_cls && c.qb == 0;
          cp_async16(Qs + 64 * LD + pc * 8,
                     ok ? a.q + ((size_t)a.n_seq * a.nq_patch + c.b) * a.q_ld + c.h * HD + pc * 8 : a.q, ok);
        } else {            // class-token key / value -> row 64 of the K / V tiles
          const bool ok = a.k_has_cls;
          const size_t krow = (size_t)a.n_kv_seq * a.nk_patch + kvb;
          const bf16* src = which == 1 ? a.k + krow * a.k_ld + c.h * HD + pc * 8 : a.v + krow * a.v_ld + c.h * HD + pc * 8;
          cp_async16((which == 1 ? Ks : Vs) + 64 * LD + pc * 8, ok ? src : a.q, ok);
        }
      }
    }
    const int k0 = kc * 64;
    const bf16* kbase = a.k + ((size_t)kvb * a.nk_patch + k0) * a.k_ld + c.h * HD;
    const bf16* vbase = a.v + ((size_t)kvb * a.nk_patch + k0) * a.v_ld + c.h * HD;
    if (full_tiles) {
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        if (row_of[i] < 64) {
          cp_async16(Ks + soff[i], kbase + koff[i], true);
          cp_async16(Vs + soff[i], vbase + voff[i], true);
        }
      }
    } else {
      const int krows = a.nk_patch - k0;
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        if (row_of[i] < 64) {
          const bool ok = row_of[i] < krows;
          cp_async16(Ks + soff[i], ok ? kbase + koff[i] : a.k, ok);
          cp_async16(Vs + soff[i], ok ? vbase + voff[i] : a.v, ok);
        }
      }
    }
  };

  // ---- per-warp state (lives across the chunks of one item) ----
  uint32_t qf[HD / 16][4];
  float o_acc[HD / 8][4];
  float m_row[2], l_row[2];

  // fetch cursor runs NST-1 steps ahead of the compute cursor; exactly one commit group per step (possibly empty)
  Cursor fc = c0;
  int f_kc = 0, f_stage = 0;
  auto fetch_next = [&]() {
    if (fc.item < n_items) {
      issue_loads(fc, f_kc, smem + f_stage * SM::STAGE);
      if (++f_kc == n_chunks) { f_kc = 0; advance(fc); }
    }
    cp_async_commit();
    if (++f_stage == NST) f_stage = 0;
  };
  __syncthreads();  // zero rows visible before any ldmatrix
#pragma unroll
  for (int i = 0; i < NST - 1; ++i) fetch_next();

  Cursor cc = c0;
  int kc = 0, stage = 0;
  while (cc.item < n_items) {
    fetch_next();
    cp_async_wait<NST - 1>();
    __syncthreads();

    bf16* st = smem + stage * SM::STAGE;
    const int qb = cc.qb;
    const bool active = warp == 4 ? (a.q_has_cls && qb == 0) : !cls_only;
    const bf16* Qs = st;
    const bf16* Ks = st + SM::T;
    const bf16* Vs = st + 2 * SM::T;

    if (active) {
      if (kc == 0) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const int row = warp * 16 + (mi & 1) * 8 + ri;
          const int col = ks * 16 + (mi >> 1) * 8;
          ldsm_x4(smem_u32(&Qs[row * LD + col]), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
        }
        m_row[0] = m_row[1] = -INFINITY;
        l_row[0] = l_row[1] = 0.f;
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) o_acc[nt][0] = o_acc[nt][1] = o_acc[nt][2] = o_acc[nt][3] = 0.f;
      }
      const bool with_cls_key = a.k_has_cls && kc == 0;   // ninth key tile: row 64 = class-token key

      // ---- S = Q K^T : 16 queries x (64 [+8]) keys per warp, raw (unscaled) scores ----
      float s[9][4];
#pragma unroll
      for (int nt = 0; nt < 9; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
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
      if (with_cls_key) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ks += 2) {
          // one ldmatrix.x4 = the (keys 64..71) B fragments of two k-steps
          const int key = 64 + ri;
          const int dim = ks * 16 + mi * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4(smem_u32(&Ks[key * LD + dim]), b0, b1, b2, b3);
          mma_bf16_16816(s[8], qf[ks], b0, b1);
          mma_bf16_16816(s[8], qf[ks + 1], b2, b3);
        }
      }
      // ---- online softmax (rows g and g+8 of this warp's tile) ----
      if (ragged_k) {
        const int k0 = kc * 64;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int key = k0 + nt * 8 + 2 * t;
          if (key >= a.nk_patch) s[nt][0] = s[nt][2] = -INFINITY;
          if (key + 1 >= a.nk_patch) s[nt][1] = s[nt][3] = -INFINITY;
        }
      }
      // only key 64 of the ninth tile exists (thread t == 0, first element)
      if (!(with_cls_key && t == 0)) s[8][0] = s[8][2] = -INFINITY;
      s[8][1] = s[8][3] = -INFINITY;
      float mx0 = fmaxf(s[8][0], s[8][1]), mx1 = fmaxf(s[8][2], s[8][3]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m_row[0], mx0 * sl2), mn1 = fmaxf(m_row[1], mx1 * sl2);   // scaled (log2) units
      const float c0 = ex2_ftz(m_row[0] - mn0), c1 = ex2_ftz(m_row[1] - mn1);
      m_row[0] = mn0; m_row[1] = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 9; ++nt) {
        s[nt][0] = ex2_ftz(fmaf(s[nt][0], sl2, -mn0)); s[nt][1] = ex2_ftz(fmaf(s[nt][1], sl2, -mn0));
        s[nt][2] = ex2_ftz(fmaf(s[nt][2], sl2, -mn1)); s[nt][3] = ex2_ftz(fmaf(s[nt][3], sl2, -mn1));
        rs0 += s[nt][0] + s[nt][1];
        rs1 += s[nt][2] + s[nt][3];
      }
      l_row[0] = l_row[0] * c0 + rs0;    // thread-partial row sums; the quad is reduced at the end
      l_row[1] = l_row[1] * c1 + rs1;
      if (kc > 0) {
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
          o_acc[nt][0] *= c0; o_acc[nt][1] *= c0;
          o_acc[nt][2] *= c1; o_acc[nt][3] *= c1;
        }
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
      if (with_cls_key) {
        // fifth k-step: keys 64..79 (64 = class token, 65..79 are zero rows with zero probabilities)
        uint32_t pa[4];
        pa[0] = pack_bf16(s[8][0], s[8][1]);
        pa[1] = pack_bf16(s[8][2], s[8][3]);
        pa[2] = 0u;
        pa[3] = 0u;
#pragma unroll
        for (int dp = 0; dp < HD / 16; ++dp) {
          const int key = 64 + (mi & 1) * 8 + ri;
          const int dim = dp * 16 + (mi >> 1) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4_trans(smem_u32(&Vs[key * LD + dim]), b0, b1, b2, b3);
          mma_bf16_16816(o_acc[2 * dp], pa, b0, b1);
          mma_bf16_16816(o_acc[2 * dp + 1], pa, b2, b3);
        }
      }

      // ---- finalize the item after its last chunk ----
      if (kc == n_chunks - 1) {
        const int h = cc.h;
        const int b = cc.b;
        const int q0 = qb * 64;
        float l0 = l_row[0], l1 = l_row[1];
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.f / l0, i1 = 1.f / l1;
        // each warp stages and stores its own 16 rows (only this warp ever touches these sO rows)
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
          *reinterpret_cast<uint32_t*>(&sO[(warp * 16 + g) * LD + nt * 8 + 2 * t]) =
              pack_bf16(o_acc[nt][0] * i0, o_acc[nt][1] * i0);
          *reinterpret_cast<uint32_t*>(&sO[(warp * 16 + g + 8) * LD + nt * 8 + 2 * t]) =
              pack_bf16(o_acc[nt][2] * i1, o_acc[nt][3] * i1);
        }
        __syncwarp();
        if (warp < 4) {
          bf16* obase = a.o + ((size_t)b * a.nq_patch + q0) * a.o_ld + h * HD;
          for (int idx = lane; idx < 16 * PIECES; idx += 32) {
            const int row = warp * 16 + idx / PIECES, pc = idx % PIECES;
            if (q0 + row < a.nq_patch)
              *reinterpret_cast<uint4*>(obase + (size_t)row * a.o_ld + pc * 8) =
                  *reinterpret_cast<const uint4*>(&sO[row * LD + pc * 8]);
          }
        } else if (lane < PIECES) {
          const size_t orow = (size_t)a.n_seq * a.nq_patch + b;   // class-token output row
          *reinterpret_cast<uint4*>(a.o + orow * a.o_ld + h * HD + lane * 8) =
              *reinterpret_cast<const uint4*>(&sO[64 * LD + lane * 8]);
        }
        __syncwarp();
      }
    }
    __syncthreads();  // everyone is done with this stage before a later fetch refills it
    if (++kc == n_chunks) { kc = 0; advance(cc); }
    if (++stage == NST) stage = 0;
  }
  cp_async_wait<0>();
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

static int attention_launch(const AttnArgs& a, int cls_only, cudaStream_t stream);

// class-token query only (last decoder layer): same kernel, only the fifth warp computes.
int attention_cls(const AttnArgs& a, cudaStream_t stream) {
  VITED_CHECK(a.head_dim == 32 || a.head_dim == 64, "attention_cls: head_dim %d not supported (32 or 64)", a.head_dim);
  VITED_CHECK(a.q_has_cls, "attention_cls: the query sequences have no class token");
  if (a.n_seq == 0) return 0;
  return attention_launch(a, 1, stream);
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
    return attention_launch(a, 0, stream);
  }
  VITED_CUDA_OK(cudaGetLastError());
  return 0;
}

static int attention_launch(const AttnArgs& a, int cls_only, cudaStream_t stream) {
  const size_t items = (size_t)a.n_seq * a.n_heads * (cls_only ? 1 : (a.nq_patch + 63) / 64);
  VITED_CHECK(items < (size_t)1 << 31, "attention: too many work items");
  static int sms = 0, per_sm32 = 0, per_sm64 = 0;   // resident CTAs per SM (registers / shared memory)
  if (sms == 0) {
    int dev = 0;
    VITED_CUDA_OK(cudaGetDevice(&dev));
    VITED_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    VITED_CUDA_OK(cudaFuncSetAttribute(attn_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<32>::BYTES));
    VITED_CUDA_OK(cudaFuncSetAttribute(attn_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<64>::BYTES));
    VITED_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm32, attn_mma_kernel<32>, 160, AttnSmem<32>::BYTES));
    VITED_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm64, attn_mma_kernel<64>, 160, AttnSmem<64>::BYTES));
    if (per_sm32 < 1) per_sm32 = 1;
    if (per_sm64 < 1) per_sm64 = 1;
  }
  const int per_sm = a.head_dim == 32 ? per_sm32 : per_sm64;
  size_t grid = (size_t)sms * per_sm;
  if (grid > items) grid = items;
  AttnArgs b = a;
  if (cls_only) {
    // with cls_only the 64-query blocks are not walked: one item per (sequence, head); the class-token query row is
    // still addressed as row n_seq*nq_patch + b of the q buffer
    if (a.head_dim == 32) attn_mma_kernel<32><<<(unsigned)grid, 160, AttnSmem<32>::BYTES, stream>>>(b, (int)items, 1);
    else attn_mma_kernel<64><<<(unsigned)grid, 160, AttnSmem<64>::BYTES, stream>>>(b, (int)items, 1);
  } else {
    if (a.head_dim == 32) attn_mma_kernel<32><<<(unsigned)grid, 160, AttnSmem<32>::BYTES, stream>>>(b, (int)items, 0);
    else attn_mma_kernel<64><<<(unsigned)grid, 160, AttnSmem<64>::BYTES, stream>>>(b, (int)items, 0);
  }
  VITED_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vited
