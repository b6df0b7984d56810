// tcgen05 attention kernels for the two hot decoder shapes (reference: models/vision_transformer.py:56-80
// Attention.forward, :174-200 CrossAttention.forward; SDPA scale head_dim^-0.5, no mask, eval mode).
//
// attn_p64_kernel -- puzzle model: 64 patch tokens (+ class token) per sequence, head_dim 32.
//   One work unit = one (sequence, head): S = Q K^T is ONE 128 x {64|80} tcgen05.mma pair (rows 0..63 = patch queries,
//   row 64 = the class-token query, rows 65..127 unused), softmax runs with one thread per query row straight out of
//   TMEM, the probabilities go back to TMEM as packed bf16 (aliasing S) and feed the second MMA as its A operand,
//   O = P V with V consumed in place as an MN-major operand (no transpose, no ldmatrix, no shuffles).
//   Warp roles (16 warps, 1 CTA / SM): warp 3 = TMA producer (12-deep ring of Q/K/V tiles, hardware 64B swizzle),
//   warp 11 = TMEM allocator, warps {4g, 4g+1, 4g+2} = softmax group g of TMEM stage g (rows 0-31, 32-63, 64); every
//   group issues its own MMAs (PV of its unit, then QK^T of its next unit back to back), so no unit waits on another
//   thread between its phases.
// Operand conventions (SW64 K-major, MN-major V, A from TMEM) were pinned on B200 by tools/umma_probe.cu.
#include "kernels.h"

namespace vited {

namespace {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct P64 {
  static constexpr int HD = 32;
  static constexpr int RB = 64;               // bytes per tile row
  static constexpr int TROWS = 80;            // 64 patch rows + class-token row (64) + 15 zero rows
  static constexpr int TB = TROWS * RB;       // 5120
  static constexpr int STAGE = 3 * TB;        // Q, K, V
  static constexpr int NS = 12;               // shared-memory ring depth (units in flight)
  static constexpr int NT = 4;                // TMEM stages = softmax groups
  static constexpr int TCOLS = 128;           // TMEM columns per stage: S/P at +0 (80), O at +96 (32)
  static constexpr int OCOL = 96;
  static constexpr int THREADS = 32 * 4 * NT;
  static constexpr int BAR_BYTES = (2 * NS + 4 * NT) * 8 + 16;
  static constexpr int BYTES = 1024 + NS * STAGE + BAR_BYTES;
};

struct P64Maps {
  CUtensorMap q_tile, q_row, k_tile, k_row, v_tile, v_row;   // boxes {32, 64} and {32, 1}, 64B swizzle
};

__global__ void __launch_bounds__(P64::THREADS, 1)
attn_p64_kernel(AttnArgs a, const __grid_constant__ P64Maps maps, int n_units) {
  using C = P64;
  extern __shared__ uint8_t attn_tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(attn_tc_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NS * C::STAGE);
  uint64_t* full = bars;                       // TMA bytes landed (count 1 + tx)
  uint64_t* empty = bars + C::NS;              // PV of the unit has read the stage (tcgen05.commit)
  uint64_t* s_full = bars + 2 * C::NS;         // S ready in TMEM (commit)
  uint64_t* o_full = s_full + C::NT;           // O ready in TMEM (commit)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(o_full + C::NT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, group = warp >> 2;
  const int H = a.n_heads;
  const int n_my = (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int NK = a.k_has_cls ? 80 : 64;        // key columns of S (64 patch keys [+ class-token key + 15 zero rows])

  // rows 64..79 of every tile: the class-token row is TMA-written when present, everything else must be exact zeros
  for (int i = tid; i < C::NS * 3 * (16 * C::RB / 16); i += C::THREADS) {
    const int tile = i / (16 * C::RB / 16), r = i % (16 * C::RB / 16);
    *reinterpret_cast<uint4*>(smem + tile * C::TB + 64 * C::RB + r * 16) = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int s = 0; s < C::NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int t = 0; t < C::NT; ++t) {
      mbar_init(&s_full[t], 1); mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&maps.q_tile);
    tma_prefetch_desc(&maps.k_tile);
    tma_prefetch_desc(&maps.v_tile);
  }
  if (warp == 11) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (quarter == 3) {
    if (group == 0 && lane == 0) {
      // ===================== TMA producer =====================
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_my; ++i) {
        const int u = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = u / H, h = u - b * H;
        const int kvb = a.kv_index ? __ldg(a.kv_index + b) : b;
        const int col = h * C::HD;
        mbar_wait(&empty[s], ph ^ 1, 50);
        uint8_t* st = smem + s * C::STAGE;
        const uint32_t bytes = 3 * 64 * C::RB + (a.q_has_cls ? C::RB : 0) + (a.k_has_cls ? 2 * C::RB : 0);
        mbar_arrive_expect_tx(&full[s], bytes);
        tma_load_2d(&maps.q_tile, &full[s], st, col, b * 64);
        if (a.q_has_cls) tma_load_2d(&maps.q_row, &full[s], st + 64 * C::RB, col, a.n_seq * 64 + b);
        tma_load_2d(&maps.k_tile, &full[s], st + C::TB, col, kvb * 64);
        tma_load_2d(&maps.v_tile, &full[s], st + 2 * C::TB, col, kvb * 64);
        if (a.k_has_cls) {
          tma_load_2d(&maps.k_row, &full[s], st + C::TB + 64 * C::RB, col, a.n_kv_seq * 64 + kvb);
          tma_load_2d(&maps.v_row, &full[s], st + 2 * C::TB + 64 * C::RB, col, a.n_kv_seq * 64 + kvb);
        }
        if (++s == C::NS) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== softmax group `group`, TMEM lane quarter `quarter` =====================
    // The group owns TMEM stage `group` and issues its own MMAs (lane 0 of its first warp): PV of the current unit as
    // soon as the three warps' probabilities are in TMEM, and QK^T of the group's NEXT unit right behind it (the
    // tensor pipe executes in issue order, so S of the next unit may overwrite P of this one), which means the next
    // scores are ready by the time this unit's output has been written.
    const float sl2 = a.scale * kLog2e;
    const int row = quarter * 32 + lane;                  // query row of the unit's tile
    const bool warp_active = quarter < 2 || a.q_has_cls;  // quarter 2 only carries the class-token query (row 64)
    const bool leader = quarter == 0 && lane == 0;
    const uint32_t t_col = tmem_base + group * C::TCOLS;
    const uint32_t t_stage = t_col + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t idesc_qk = umma_idesc_bf16(128, NK);
    const uint32_t idesc_pv = umma_idesc_bf16(128, C::HD) | kIdescBMajorMN;
    const int pv_steps = NK / 16;
    auto issue_qk = [&](int i) {      // leader only: S(stage) = Q K^T of local unit i
      const int s = i % C::NS;
      mbar_wait(&full[s], (uint32_t)(i / C::NS) & 1u, 51);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(smem + s * C::STAGE);
      const uint64_t dq = umma_desc_sw(q_addr, 64);
      const uint64_t dk = umma_desc_sw(q_addr + C::TB, 64);
#pragma unroll
      for (int k = 0; k < C::HD / 16; ++k) umma_bf16(t_col, dq + 2 * k, dk + 2 * k, idesc_qk, k);
      umma_commit(&s_full[group]);
    };
    if (leader && group < n_my) issue_qk(group);
    uint32_t ph = 0;
    for (int i = group; i < n_my; i += C::NT) {
      const int u = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = u / H, h = u - b * H;
      mbar_wait(&s_full[group], ph, 54);
      tc_fence_after();
      float l = 1.f;
      if (warp_active) {
        uint32_t v0[32], v1[32];
        uint32_t vc[8];
        tmem_ld_32x32b_x32(t_stage, v0);
        tmem_ld_32x32b_x32(t_stage + 32, v1);
        if (a.k_has_cls) tmem_ld_32x32b_x8(t_stage + 64, vc);
        tmem_ld_wait();
        float mx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mx[j] = fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j]));
#pragma unroll
        for (int j = 4; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
        float mxa = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if (a.k_has_cls) mxa = fmaxf(mxa, __uint_as_float(vc[0]));
        const float mneg = -mxa * sl2;
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p0 = ex2_ftz(fmaf(__uint_as_float(v0[2 * j]), sl2, mneg));
          const float p1 = ex2_ftz(fmaf(__uint_as_float(v0[2 * j + 1]), sl2, mneg));
          sum[(2 * j) & 3] += p0; sum[(2 * j + 1) & 3] += p1;
          pk[j] = pack_bf16(p0, p1);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p0 = ex2_ftz(fmaf(__uint_as_float(v1[2 * j]), sl2, mneg));
          const float p1 = ex2_ftz(fmaf(__uint_as_float(v1[2 * j + 1]), sl2, mneg));
          sum[(2 * j) & 3] += p0; sum[(2 * j + 1) & 3] += p1;
          pk[16 + j] = pack_bf16(p0, p1);
        }
        tmem_st_32x32b_x32(t_stage, pk);
        if (a.k_has_cls) {
          const float pc = ex2_ftz(fmaf(__uint_as_float(vc[0]), sl2, mneg));
          sum[0] += pc;
          uint32_t pc8[8] = {pack_bf16(pc, 0.f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          tmem_st_32x32b_x8(t_stage + 32, pc8);
        }
        l = (sum[0] + sum[1]) + (sum[2] + sum[3]);
        tmem_st_wait();
      }
      tc_fence_before();
      asm volatile("bar.sync %0, 96;" ::"r"(group + 1) : "memory");   // the group's three warps: all of P is in TMEM
      if (leader) {
        tc_fence_after();
        const int s = i % C::NS;
        const uint32_t v_addr = smem_u32(smem + s * C::STAGE + 2 * C::TB);
        for (int k = 0; k < pv_steps; ++k)   // 16 keys per step = two 8-key groups of 512 B
          umma_bf16_ts(t_col + C::OCOL, t_col + 8 * k, umma_desc_sw(v_addr + k * 1024, 64), idesc_pv, k);
        umma_commit(&empty[s]);
        umma_commit(&o_full[group]);
        if (i + C::NT < n_my) issue_qk(i + C::NT);
      }
      __syncwarp();
      mbar_wait(&o_full[group], ph, 55);
      tc_fence_after();
      if (warp_active) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(t_stage + C::OCOL, ov);
        tmem_ld_wait();
        const float inv = 1.f / l;
        size_t orow;
        bool valid = true;
        if (row < 64) orow = (size_t)b * 64 + row;
        else if (row == 64 && a.q_has_cls) orow = (size_t)a.n_seq * 64 + b;
        else { valid = false; orow = 0; }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(a.o + orow * a.o_ld + h * C::HD);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(ov[8 * c + 0]) * inv, __uint_as_float(ov[8 * c + 1]) * inv);
            w.y = pack_bf16(__uint_as_float(ov[8 * c + 2]) * inv, __uint_as_float(ov[8 * c + 3]) * inv);
            w.z = pack_bf16(__uint_as_float(ov[8 * c + 4]) * inv, __uint_as_float(ov[8 * c + 5]) * inv);
            w.w = pack_bf16(__uint_as_float(ov[8 * c + 6]) * inv, __uint_as_float(ov[8 * c + 7]) * inv);
            dst[c] = w;
          }
        }
      }
      ph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 11) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// true when the tcgen05 kernel covers this problem (the mma.sync kernel in attention.cu covers everything else)
bool attention_tc_supported(const AttnArgs& a) {
  return a.head_dim == 32 && a.nq_patch == 64 && a.nk_patch == 64 && a.n_heads >= 1;
}

int attention_tc(const AttnArgs& a, cudaStream_t stream) {
  VITED_CHECK(attention_tc_supported(a), "attention_tc: unsupported shape");
  const size_t units = (size_t)a.n_seq * a.n_heads;
  VITED_CHECK(units < ((size_t)1 << 31), "attention_tc: too many work units");
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    VITED_CUDA_OK(cudaGetDevice(&dev));
    VITED_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    VITED_CUDA_OK(cudaFuncSetAttribute(attn_p64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P64::BYTES));
  }
  const uint64_t cols = (uint64_t)a.n_heads * 32;
  const uint64_t q_rows = (uint64_t)a.n_seq * 64 + (a.q_has_cls ? a.n_seq : 0);
  const uint64_t k_rows = (uint64_t)a.n_kv_seq * 64 + (a.k_has_cls ? a.n_kv_seq : 0);
  P64Maps maps;
  if (make_tmap_bf16_2d(&maps.q_tile, a.q, cols, q_rows, (uint64_t)a.q_ld * 2, 32, 64, 64)) return 1;
  if (make_tmap_bf16_2d(&maps.q_row, a.q, cols, q_rows, (uint64_t)a.q_ld * 2, 32, 1, 64)) return 1;
  if (make_tmap_bf16_2d(&maps.k_tile, a.k, cols, k_rows, (uint64_t)a.k_ld * 2, 32, 64, 64)) return 1;
  if (make_tmap_bf16_2d(&maps.k_row, a.k, cols, k_rows, (uint64_t)a.k_ld * 2, 32, 1, 64)) return 1;
  if (make_tmap_bf16_2d(&maps.v_tile, a.v, cols, k_rows, (uint64_t)a.v_ld * 2, 32, 64, 64)) return 1;
  if (make_tmap_bf16_2d(&maps.v_row, a.v, cols, k_rows, (uint64_t)a.v_ld * 2, 32, 1, 64)) return 1;
  const unsigned grid = (unsigned)(units < (size_t)sms ? units : (size_t)sms);
  attn_p64_kernel<<<grid, P64::THREADS, P64::BYTES, stream>>>(a, maps, (int)units);
  VITED_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vited
