// attn_p64x2_kernel -- puzzle-shape attention (64 patch tokens (+ class token) per sequence, head_dim 32) with TWO
// (sequence, head) units per 128-row tcgen05 tile (reference: models/vision_transformer.py:56-80 Attention.forward,
// :174-200 CrossAttention.forward; SDPA scale head_dim^-0.5, no mask, eval mode).
//
// Why: in attn_p64_kernel (attention_tc.cu) a unit fills rows 0..64 of a 128-row tile. TMEM lane quarter q can only be
// read by warps with warp % 4 == q, and those warps live on SM sub-partition q -- so the softmax of every unit ran on
// sub-partitions 0 and 1 (rows 0-31, 32-63), sub-partition 2 spent the same 65 MUFU instructions per unit on ONE live
// row (the class-token query) and sub-partition 3 only carried the service warps: ncu MUFU 53 % = three busy
// sub-partitions at ~70 % and one idle. Here the tile holds the patch queries of unit A in rows 0-63 and of unit B in
// rows 64-127:
//     S = [Q_A; Q_B] [K_A; K_B; kcls_A; 0; kcls_B; 0...]^T     one 128 x 144 (128 without class-token keys) MMA pair
//     row of A: softmax over columns 0-63 and 128, row of B: over columns 64-127 and 130; the other unit's columns get
//     exact zeros in P, so O = P [V_A; V_B; vcls_A; 0; vcls_B; 0...] is block diagonal: rows 0-63 = O_A, rows 64-127 = O_B.
// The wasted off-diagonal score blocks cost tensor-pipe time that is free here (the pipe ran at 16-21 %); every lane
// quarter, i.e. every sub-partition, now carries 32 live rows per tile. The two class-token QUERY rows of a pair no
// longer fit in the tile: four CUDA-core warps (one per sub-partition) compute them straight from the K / V tiles in
// shared memory (65 keys x 32 dims per row, fp32, warp-shuffle softmax) while the tile pipeline works on the same stage.
//
// Warps (16, one CTA per SM): 0-3 = softmax of TMEM stage 0 (lane quarters 0-3), 4-7 = TMEM stage 1, 8-11 = class-token
// warps (pair i goes to warp 8 + i % 4), 12 = TMA producer, 13 = MMA issuer (warp-uniform, elected lane), 14 = TMEM
// allocator. Shared memory: ring of NS pair stages {Q 128 x 64 B, K 144 x 64 B, V 144 x 64 B, 2 class-token query
// rows}, hardware 64B swizzle. TMEM per stage: S / P at +0 (144 columns), O at +160 (32 columns).
// Operand conventions (SW64 K-major Q / K with N = 144, MN-major V with K = 144, A operand from TMEM) were pinned on
// B200 by tools/umma_probe.cu (profiles/r01b_umma_probe.txt: qk32 N=144, pv32ts K=144).
// MEASURED NEGATIVE RESULT (profiles/README.md, round 2 second session): correct on every attention test, but 0.190 /
// 0.215 ms against 0.147 / 0.159 ms per 262 k rows (self / cross): the kernel is bound by the per-stage latency chain
// (scores -> softmax -> P -> PV -> O), and two 176-column TMEM stages of two units each keep fewer chains in flight
// than four one-unit stages. Only compiled into -DVITED_EXPERIMENTAL builds (tools/build_variants.sh).
#include "kernels.h"

#ifdef VITED_EXPERIMENTAL
namespace vited {

namespace {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct PX2 {
  static constexpr int HD = 32;
  static constexpr int RB = 64;                 // bytes per tile row
  static constexpr int QB = 128 * RB;           // Q tile: rows 0-63 unit A, 64-127 unit B
  static constexpr int KROWS = 144;             // 64 + 64 patch keys, class-token key of A (row 128) and B (row 130), zero rows
  static constexpr int KB = KROWS * RB;         // 9216
  static constexpr int QC_OFF = QB + 2 * KB;    // the two class-token query rows (64 B each)
  static constexpr int STAGE = 27 * 1024;       // 8192 + 9216 + 9216 + 128, rounded up to the 1 KB tile alignment
  static constexpr int NS = 7;                  // pair stages in the shared-memory ring (14 units in flight)
  static constexpr int NT = 2;                  // TMEM stages
  static constexpr int TCOLS = 192;             // TMEM columns per stage
  static constexpr int OCOL = 160;
  static constexpr int THREADS = 512;
  static constexpr int CLS_WARP0 = 8, PRODUCER_WARP = 12, ISSUER_WARP = 13, ALLOC_WARP = 14;
  static constexpr int BAR_BYTES = (2 * NS + 3 * NT) * 8 + 16;   // full, empty, s_full, o_full, p_ready + TMEM holder
  static constexpr int BYTES = 1024 + NS * STAGE + BAR_BYTES;
  static_assert(QB + 2 * KB + 2 * 128 <= STAGE, "stage layout");
  static_assert(BYTES <= 232448, "shared memory budget");
};

struct PX2Maps {
  CUtensorMap q_tile, q_row, k_tile, k_row, v_tile, v_row;   // boxes {32, 64} and {32, 1}, 64B swizzle
};

// byte offset of 16-byte chunk c of tile row r (64-byte rows, hardware 64B swizzle, tile base 512-byte aligned)
__device__ __forceinline__ uint32_t sw64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// 8 elements of a key row (one 16-byte chunk) against q[8c .. 8c+7]
__device__ __forceinline__ void dot8(const uint4& kk, const float* q8, float (&acc)[2]) {
  const float2 k0 = unpack_act(kk.x), k1 = unpack_act(kk.y), k2 = unpack_act(kk.z), k3 = unpack_act(kk.w);
  acc[0] = fmaf(k0.x, q8[0], acc[0]); acc[1] = fmaf(k0.y, q8[1], acc[1]);
  acc[0] = fmaf(k1.x, q8[2], acc[0]); acc[1] = fmaf(k1.y, q8[3], acc[1]);
  acc[0] = fmaf(k2.x, q8[4], acc[0]); acc[1] = fmaf(k2.y, q8[5], acc[1]);
  acc[0] = fmaf(k3.x, q8[6], acc[0]); acc[1] = fmaf(k3.y, q8[7], acc[1]);
}

__global__ void __launch_bounds__(PX2::THREADS, 1)
attn_p64x2_kernel(AttnArgs a, const __grid_constant__ PX2Maps maps, int n_units) {
  using C = PX2;
  extern __shared__ uint8_t attn_pair_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(attn_pair_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NS * C::STAGE);
  uint64_t* full = bars;                       // [NS] TMA bytes of a pair stage landed (count 1 + tx)
  uint64_t* empty = bars + C::NS;              // [NS] PV of the pair has read the stage (commit) + its class-token warp
  uint64_t* s_full = bars + 2 * C::NS;         // [NT] S ready in TMEM (commit)
  uint64_t* o_full = s_full + C::NT;           // [NT] O ready in TMEM (commit)
  uint64_t* p_ready = o_full + C::NT;          // [NT] P written to TMEM (the stage's 4 softmax warps)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(p_ready + C::NT);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const int H = a.n_heads;
  const int n_pairs = (n_units + 1) >> 1;
  const int n_my = (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // pairs of this CTA
  const int NK = a.k_has_cls ? 144 : 128;      // key columns of S

  // Everything the TMA does not overwrite must be finite (zeros): key / value rows 128..143 (rows 128 / 129 take the
  // class-token keys when present), and the B half of every tile in case the launch ends on a single unit.
  for (int i = tid; i < C::NS * C::STAGE / 16; i += C::THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < C::NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }
    for (int t = 0; t < C::NT; ++t) { mbar_init(&s_full[t], 1); mbar_init(&o_full[t], 1); mbar_init(&p_ready[t], 4); }
    fence_mbar_init();
    tma_prefetch_desc(&maps.q_tile);
    tma_prefetch_desc(&maps.k_tile);
    tma_prefetch_desc(&maps.v_tile);
  }
  if (warp == C::ALLOC_WARP) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_launch_dependents();
  pdl_wait();           // the prologue above overlapped the previous kernel's tail; global memory only from here on

  if (warp == C::PRODUCER_WARP) {
    // ===================== TMA producer =====================
    // Lane 0 issues. With a key/value index (cross-attention: kv_index[b]) the whole warp fetches the indices of the
    // next 32 units (16 pairs) at once instead of one dependent __ldg per unit in the issuing thread's chain.
    int s = 0;
    uint32_t ph = 0;
    auto issue_pair = [&](int i, int kvb_a, int kvb_b) {     // lane 0: the TMA loads of local pair i
      const int u0 = 2 * ((int)blockIdx.x + i * (int)gridDim.x);
      const int nu = u0 + 1 < n_units ? 2 : 1;
      mbar_wait(&empty[s], ph ^ 1, 50);
      uint8_t* st = smem + s * C::STAGE;
      const uint32_t per_unit = 3 * 64 * C::RB + (a.q_has_cls ? C::RB : 0) + (a.k_has_cls ? 2 * C::RB : 0);
      mbar_arrive_expect_tx(&full[s], nu * per_unit);
      for (int x = 0; x < nu; ++x) {
        const int u = u0 + x;
        const int b = u / H, h = u - b * H;
        const int kvb = (x ? kvb_b : kvb_a) < 0 ? b : (x ? kvb_b : kvb_a);
        const int col = h * C::HD;
        tma_load_2d(&maps.q_tile, &full[s], st + x * 64 * C::RB, col, b * 64);
        tma_load_2d(&maps.k_tile, &full[s], st + C::QB + x * 64 * C::RB, col, kvb * 64);
        tma_load_2d(&maps.v_tile, &full[s], st + C::QB + C::KB + x * 64 * C::RB, col, kvb * 64);
        // single rows go to 128-byte-aligned addresses (a TMA destination must be): query rows 128 B apart, the
        // class-token key / value of unit x in tile row 128 + 2x
        if (a.q_has_cls) tma_load_2d(&maps.q_row, &full[s], st + C::QC_OFF + x * 128, col, a.n_seq * 64 + b);
        if (a.k_has_cls) {
          tma_load_2d(&maps.k_row, &full[s], st + C::QB + (128 + 2 * x) * C::RB, col, a.n_kv_seq * 64 + kvb);
          tma_load_2d(&maps.v_row, &full[s], st + C::QB + C::KB + (128 + 2 * x) * C::RB, col, a.n_kv_seq * 64 + kvb);
        }
      }
      if (++s == C::NS) { s = 0; ph ^= 1; }
    };
    if (a.kv_index == nullptr) {
      if (lane == 0)
        for (int i = 0; i < n_my; ++i) issue_pair(i, -1, -1);
    } else {
      for (int i0 = 0; i0 < n_my; i0 += 16) {
        // lane 2j + x holds the index of unit x of local pair i0 + j
        int my_kvb = 0;
        {
          const int i = i0 + (lane >> 1);
          const int u = 2 * ((int)blockIdx.x + i * (int)gridDim.x) + (lane & 1);
          if (i < n_my && u < n_units) my_kvb = __ldg(a.kv_index + u / H);
        }
        const int jn = n_my - i0 < 16 ? n_my - i0 : 16;
        for (int j = 0; j < jn; ++j) {
          const int ka = __shfl_sync(0xffffffffu, my_kvb, 2 * j);
          const int kb = __shfl_sync(0xffffffffu, my_kvb, 2 * j + 1);
          if (lane == 0) issue_pair(i0 + j, ka, kb);
        }
      }
    }
  } else if (warp == C::ISSUER_WARP) {
    // ===================== MMA issuer (warp-uniform code, one elected lane) =====================
    // Pairs alternate between the two TMEM stages. PV(i) as soon as the stage's probabilities are in TMEM, QK^T of the
    // stage's next pair (i + NT) right behind it: the tensor pipe runs in issue order, so those scores may overwrite
    // P(i); O(i) has its own columns and is read by the softmax warps before they publish P(i + NT).
    const uint32_t idesc_qk = umma_idesc_f16(128, NK);
    const uint32_t idesc_pv = umma_idesc_f16(128, C::HD) | kIdescBMajorMN;
    const uint32_t smem0 = smem_u32(smem);
    auto issue_qk = [&](int i) {      // S(stage i % NT) = [Q_A; Q_B] [K_A; K_B; kcls...]^T of local pair i
      const int s = i % C::NS, t = i % C::NT;
      mbar_wait(&full[s], (uint32_t)(i / C::NS) & 1u, 51);
      tc_fence_after();
      const uint32_t q_addr = smem0 + s * C::STAGE;
      const uint64_t dq = umma_desc_sw(q_addr, 64);
      const uint64_t dk = umma_desc_sw(q_addr + C::QB, 64);
      const uint32_t t_col = tmem_base + t * C::TCOLS;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < C::HD / 16; ++k) umma_f16(t_col, dq + 2 * k, dk + 2 * k, idesc_qk, k);
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    for (int i = 0; i < C::NT && i < n_my; ++i) issue_qk(i);
    for (int i = 0; i < n_my; ++i) {
      const int s = i % C::NS, t = i % C::NT;
      const uint32_t t_col = tmem_base + t * C::TCOLS;
      mbar_wait(&p_ready[t], (uint32_t)(i / C::NT) & 1u, 53);
      tc_fence_after();
      const uint64_t dv = umma_desc_sw(smem0 + s * C::STAGE + C::QB + C::KB, 64);
      if (elect_one_sync()) {
        if (NK == 144) {
#pragma unroll
          for (int k = 0; k < 9; ++k)   // 16 keys per step = two 8-key groups of 512 B = +64 in the (addr >> 4) field
            umma_f16_ts(t_col + C::OCOL, t_col + 8 * k, dv + 64 * k, idesc_pv, k);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16_ts(t_col + C::OCOL, t_col + 8 * k, dv + 64 * k, idesc_pv, k);
        }
        umma_commit(&empty[s]);
        umma_commit(&o_full[t]);
      }
      __syncwarp();
      if (i + C::NT < n_my) issue_qk(i + C::NT);
    }
  } else if (warp >= C::CLS_WARP0 && warp < C::CLS_WARP0 + 4) {
    // ===================== class-token query rows of the pairs on the CUDA cores =====================
    // Warp c takes the pairs i = c, c + 4, ...: for each unit of the pair, scores of the class-token query against the
    // unit's 64 (+ 1) keys (lane t: keys t and t + 32, every lane the class-token key), warp-shuffle softmax, then
    // o = P V with lane (half, d2): keys of half `half`, dims 2 d2 and 2 d2 + 1. Reads the stage the tile pipeline is
    // working on; the stage is released when both this warp and PV of the pair are done with it.
    const int c = warp - C::CLS_WARP0;
    const float sl2 = a.scale * kLog2e;
    const uint32_t smem0 = smem_u32(smem);
    const int half = lane >> 4, d2 = lane & 15;
    for (int i = c; i < n_my; i += 4) {
      const int s = i % C::NS;
      mbar_wait(&full[s], (uint32_t)(i / C::NS) & 1u, 56);
      if (a.q_has_cls) {
        const int u0 = 2 * ((int)blockIdx.x + i * (int)gridDim.x);
        const uint32_t st = smem0 + s * C::STAGE;
        const uint32_t kt = st + C::QB, vt = kt + C::KB;
        for (int x = 0; x < 2 && u0 + x < n_units; ++x) {
          const int u = u0 + x;
          const int b = u / H, h = u - b * H;
          // the query row: a single 64-byte row at QC_OFF + 128 x; the 64B hardware swizzle XORs the 16-byte chunk
          // index with address bits 7-8, i.e. with x here
          float q[32];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint4 qq = lds_u4(st + C::QC_OFF + x * 128 + ((cc ^ x) << 4));
            const float2 q0 = unpack_act(qq.x), q1 = unpack_act(qq.y), q2 = unpack_act(qq.z), q3 = unpack_act(qq.w);
            q[8 * cc + 0] = q0.x * sl2; q[8 * cc + 1] = q0.y * sl2; q[8 * cc + 2] = q1.x * sl2; q[8 * cc + 3] = q1.y * sl2;
            q[8 * cc + 4] = q2.x * sl2; q[8 * cc + 5] = q2.y * sl2; q[8 * cc + 6] = q3.x * sl2; q[8 * cc + 7] = q3.y * sl2;
          }
          // scores in the log2 domain
          float sc[3];
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const int r = kk < 2 ? x * 64 + kk * 32 + lane : 128 + 2 * x;
            float acc[2] = {0.f, 0.f};
            if (kk < 2 || a.k_has_cls) {
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) dot8(lds_u4(kt + sw64_off(r, cc)), q + 8 * cc, acc);
            }
            sc[kk] = acc[0] + acc[1];
          }
          float m = warp_max(fmaxf(sc[0], sc[1]));
          if (a.k_has_cls) m = fmaxf(m, sc[2]);
          const float p0 = ex2_ftz(sc[0] - m), p1 = ex2_ftz(sc[1] - m);
          const float pc = a.k_has_cls ? ex2_ftz(sc[2] - m) : 0.f;
          const float l = warp_sum(p0 + p1) + pc;
          // o = P V: this lane's half of the keys (the upper half one row ahead, so that the two halves of the warp read
          // rows of opposite parity = disjoint banks), its two dims
          float o0 = 0.f, o1 = 0.f;
#pragma unroll 8
          for (int j = 0; j < 32; ++j) {
            const int jj = (j + half) & 31;                       // key inside the half
            const float pa = __shfl_sync(0xffffffffu, p0, jj);    // p of key jj        (wanted by half 0)
            const float pb = __shfl_sync(0xffffffffu, p1, jj);    // p of key 32 + jj   (wanted by half 1)
            const float p = half ? pb : pa;
            const int r = x * 64 + half * 32 + jj;
            const float2 vv = unpack_act(lds_u32(vt + sw64_off(r, d2 >> 2) + (d2 & 3) * 4));
            o0 = fmaf(p, vv.x, o0);
            o1 = fmaf(p, vv.y, o1);
          }
          o0 += __shfl_xor_sync(0xffffffffu, o0, 16);
          o1 += __shfl_xor_sync(0xffffffffu, o1, 16);
          if (a.k_has_cls) {
            const float2 vv = unpack_act(lds_u32(vt + sw64_off(128 + 2 * x, d2 >> 2) + (d2 & 3) * 4));
            o0 = fmaf(pc, vv.x, o0);
            o1 = fmaf(pc, vv.y, o1);
          }
          if (half == 0) {
            const float inv = 1.f / l;
            uint32_t* dst = reinterpret_cast<uint32_t*>(a.o + ((size_t)a.n_seq * 64 + b) * a.o_ld + h * C::HD);
            dst[d2] = pack_act(o0 * inv, o1 * inv);
          }
        }
      }
      __syncwarp();                                  // every lane has finished reading the stage
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  } else if (warp < 4 * C::NT) {
    // ===================== softmax of TMEM stage `stage`, lane quarter `quarter` =====================
    // Quarters 0 / 1 hold unit A's rows 0-31 / 32-63, quarters 2 / 3 unit B's: one thread = one query row. Scores ->
    // probabilities (packed fp16 over S, exact zeros in the other unit's key columns) -> signal the issuer -> read O.
    const int stage = warp >> 2, quarter = warp & 3;
    const int x = quarter >> 1;                            // unit of the pair
    const int row = (quarter & 1) * 32 + lane;             // query row inside the unit
    const float sl2 = a.scale * kLog2e;
    const uint32_t t_col = tmem_base + stage * C::TCOLS;
    const uint32_t t_lane = t_col + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t ph = 0;
    for (int i = stage; i < n_my; i += C::NT) {
      const int u = 2 * ((int)blockIdx.x + i * (int)gridDim.x) + x;
      const bool live = u < n_units;                       // (a launch with an odd number of units ends on half a pair)
      const int b = u / H, h = u - b * H;
      mbar_wait(&s_full[stage], ph, 54);
      tc_fence_after();
      float l = 1.f;
      {
        uint32_t v0[32], v1[32];
        uint32_t vc[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_ld_32x32b_x32(t_lane + x * 64, v0);
        tmem_ld_32x32b_x32(t_lane + x * 64 + 32, v1);
        if (a.k_has_cls) tmem_ld_32x32b_x8(t_lane + 128, vc);
        tmem_ld_wait();
        float mx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mx[j] = fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j]));
#pragma unroll
        for (int j = 4; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
        float mxa = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        const float scls = x ? __uint_as_float(vc[2]) : __uint_as_float(vc[0]);   // key row 128 + 2x
        if (a.k_has_cls) mxa = fmaxf(mxa, scls);
        const float mneg = -mxa * sl2;
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p0 = ex2_ftz(fmaf(__uint_as_float(v0[2 * j]), sl2, mneg));
          const float p1 = ex2_ftz(fmaf(__uint_as_float(v0[2 * j + 1]), sl2, mneg));
          sum[(2 * j) & 3] += p0; sum[(2 * j + 1) & 3] += p1;
          pk[j] = pack_act(p0, p1);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p0 = ex2_ftz(fmaf(__uint_as_float(v1[2 * j]), sl2, mneg));
          const float p1 = ex2_ftz(fmaf(__uint_as_float(v1[2 * j + 1]), sl2, mneg));
          sum[(2 * j) & 3] += p0; sum[(2 * j + 1) & 3] += p1;
          pk[16 + j] = pack_act(p0, p1);
        }
        // packed columns 0-31 = keys of A, 32-63 = keys of B, 64 / 65 = (class-token key of A / B, zero row), 66-71 = zero rows
        tmem_st_32x32b_x32(t_lane + x * 32, pk);
#pragma unroll
        for (int j = 0; j < 32; ++j) pk[j] = 0u;
        tmem_st_32x32b_x32(t_lane + (x ^ 1) * 32, pk);
        if (a.k_has_cls) {
          const float pc = ex2_ftz(fmaf(scls, sl2, mneg));
          sum[0] += pc;
          const uint32_t pcp = pack_act(pc, 0.f);
          uint32_t pc8[8] = {x ? 0u : pcp, x ? pcp : 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          tmem_st_32x32b_x8(t_lane + 64, pc8);
        }
        l = (sum[0] + sum[1]) + (sum[2] + sum[3]);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[stage]);   // 4 warps: all of P is in TMEM -> the issuer may run PV
      mbar_wait(&o_full[stage], ph, 55);
      tc_fence_after();
      {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(t_lane + C::OCOL, ov);
        tmem_ld_wait();
        if (live) {
          const float inv = 1.f / l;
          uint4* dst = reinterpret_cast<uint4*>(a.o + ((size_t)b * 64 + row) * a.o_ld + h * C::HD);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint4 w;
            w.x = pack_act(__uint_as_float(ov[8 * cc + 0]) * inv, __uint_as_float(ov[8 * cc + 1]) * inv);
            w.y = pack_act(__uint_as_float(ov[8 * cc + 2]) * inv, __uint_as_float(ov[8 * cc + 3]) * inv);
            w.z = pack_act(__uint_as_float(ov[8 * cc + 4]) * inv, __uint_as_float(ov[8 * cc + 5]) * inv);
            w.w = pack_act(__uint_as_float(ov[8 * cc + 6]) * inv, __uint_as_float(ov[8 * cc + 7]) * inv);
            dst[cc] = w;
          }
        }
      }
      ph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C::ALLOC_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool attention_pair_supported(const AttnArgs& a) {
  return a.n_heads >= 1 && a.head_dim == 32 && a.nq_patch == 64 && a.nk_patch == 64 && (a.o_ld % 8) == 0;
}

int attention_pair(const AttnArgs& a, cudaStream_t stream) {
  VITED_CHECK(attention_pair_supported(a), "attention_pair: unsupported shape");
  const size_t units = (size_t)a.n_seq * a.n_heads;
  VITED_CHECK(units < ((size_t)1 << 30), "attention_pair: too many work units");
  static PerDeviceOnce once;
  const int sms = device_sm_count();
  if (once.first())
    VITED_CUDA_OK(cudaFuncSetAttribute(attn_p64x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PX2::BYTES));
  const uint64_t cols = (uint64_t)a.n_heads * 32;
  const uint64_t q_rows = (uint64_t)a.n_seq * 64 + (a.q_has_cls ? a.n_seq : 0);
  const uint64_t k_rows = (uint64_t)a.n_kv_seq * 64 + (a.k_has_cls ? a.n_kv_seq : 0);
  PX2Maps maps;
  if (make_tmap_act_2d(&maps.q_tile, a.q, cols, q_rows, (uint64_t)a.q_ld * 2, 32, 64, 64)) return 1;
  if (make_tmap_act_2d(&maps.q_row, a.q, cols, q_rows, (uint64_t)a.q_ld * 2, 32, 1, 64)) return 1;
  if (make_tmap_act_2d(&maps.k_tile, a.k, cols, k_rows, (uint64_t)a.k_ld * 2, 32, 64, 64)) return 1;
  if (make_tmap_act_2d(&maps.k_row, a.k, cols, k_rows, (uint64_t)a.k_ld * 2, 32, 1, 64)) return 1;
  if (make_tmap_act_2d(&maps.v_tile, a.v, cols, k_rows, (uint64_t)a.v_ld * 2, 32, 64, 64)) return 1;
  if (make_tmap_act_2d(&maps.v_row, a.v, cols, k_rows, (uint64_t)a.v_ld * 2, 32, 1, 64)) return 1;
  const size_t pairs = (units + 1) / 2;
  const unsigned grid = (unsigned)(pairs < (size_t)sms ? pairs : (size_t)sms);
  VITED_CUDA_OK(launch_pdl(attn_p64x2_kernel, dim3(grid), dim3(PX2::THREADS), PX2::BYTES, stream, a, maps, (int)units));
  VITED_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vited
#endif  // VITED_EXPERIMENTAL
