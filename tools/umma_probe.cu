// Standalone probe (NOT part of the product library): pins down the tcgen05 operand conventions the attention kernels
// rely on, against a CPU reference, one variant per process:
//   qk32   : D[128x144] = Q[128x32] * K[144x32]^T, both K-major with 64-byte swizzle (head_dim 32)
//   pv32ss : D[128x32]  = P[128x144] * V[144x32], P K-major SW128 in smem, V "MN-major" SW64 (rows = keys)
//   pv32ts : same, P read from TMEM (packed bf16 pairs written with tcgen05.st)
//   pv64ss : D[128x64]  = P[128x128] * V[128x64], V MN-major SW128
//   pv64ts : same, P from TMEM
// argv: <variant> [swap_lbo_sbo]
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ inline uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t swz) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)swz << 61;
  return d;
}

struct Params {
  int M, N, K;            // MMA problem (K multiple of 16)
  int a_bytes, b_bytes;   // smem images
  int a_from_tmem;        // 1: A is [128][K/2] packed words in global -> tcgen05.st -> TMEM
  uint32_t a_lbo, a_sbo, a_swz, a_kstep;   // descriptor fields for A (smem) and byte advance per 16-wide k-step
  uint32_t a_kblock_elems, a_kblock_bytes; // K-major SW128: every 64 k elements jump a_kblock_bytes
  uint32_t b_lbo, b_sbo, b_swz, b_kstep;
  uint32_t b_kblock_elems, b_kblock_bytes;
  uint32_t idesc;
  int d_lane;              // lane offset of the D address (M = 64 layout probe)
  int prefill;             // 1: fill the D columns of all 128 lanes with a sentinel first
};

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, float* out, Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((p.a_bytes + 1023) & ~1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (!p.a_from_tmem)
    for (int i = tid; i < p.a_bytes / 16; i += 128) ((uint4*)sA)[i] = ((const uint4*)a_img)[i];
  for (int i = tid; i < p.b_bytes / 16; i += 128) ((uint4*)sB)[i] = ((const uint4*)b_img)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&holder)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = holder;
  const uint32_t tA = tbase + 256;   // packed P lives at columns 256..
  if (p.a_from_tmem) {
    const uint32_t* w = (const uint32_t*)a_img + (size_t)tid * (p.K / 2);
    for (int c0 = 0; c0 < p.K / 2; c0 += 8) {
      uint32_t v[8];
      for (int i = 0; i < 8; ++i) v[i] = w[c0 + i];
      const uint32_t taddr = tA + c0 + ((uint32_t)(warp * 32) << 16);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
                   "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  if (p.prefill) {
    for (int c0 = 0; c0 < p.N; c0 += 8) {
      const uint32_t sv = __float_as_uint(12345.f);
      const uint32_t taddr = tbase + c0 + ((uint32_t)(warp * 32) << 16);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(sv) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  const uint32_t tD = tbase + ((uint32_t)p.d_lane << 16);
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int ks = 0; ks < p.K / 16; ++ks) {
      const int kel = ks * 16;
      const uint32_t a_off = (kel / p.a_kblock_elems) * p.a_kblock_bytes + ((kel % p.a_kblock_elems) / 16) * p.a_kstep;
      const uint32_t b_off = (kel / p.b_kblock_elems) * p.b_kblock_bytes + ((kel % p.b_kblock_elems) / 16) * p.b_kstep;
      const uint64_t db = make_desc(smem_u32(sB) + b_off, p.b_lbo, p.b_sbo, p.b_swz);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      if (p.a_from_tmem) {
        const uint32_t ta = tA + ks * 8;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(tD),
                     "r"(ta), "l"(db), "r"(p.idesc), "r"(acc)
                     : "memory");
      } else {
        const uint64_t da = make_desc(smem_u32(sA) + a_off, p.a_lbo, p.a_sbo, p.a_swz);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tD),
                     "l"(da), "l"(db), "r"(p.idesc), "r"(acc)
                     : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      if (clock64() - t0 > 2000000000LL) { if (tid == 0) printf("probe: timeout\n"); __trap(); }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < p.N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tbase + c0 + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(size_t)tid * p.N + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

static uint16_t f2bf(float f) { __nv_bfloat16 b = __float2bfloat16(f); uint16_t u; memcpy(&u, &b, 2); return u; }
static float bf2f(uint16_t u) { __nv_bfloat16 b; memcpy(&b, &u, 2); return __bfloat162float(b); }

// physical image of a [rows][cols] bf16 matrix whose rows are `row_bytes` wide boxes with a TMA-style swizzle
// (row_bytes 64 -> SW64, 128 -> SW128); cols beyond one box go to further boxes of rows*row_bytes bytes each.
static void put(std::vector<uint8_t>& img, int rows, int cols, int row_bytes, const std::vector<float>& m) {
  const int box_cols = row_bytes / 2;
  const int nbox = (cols + box_cols - 1) / box_cols;
  img.assign((size_t)nbox * rows * row_bytes, 0);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const int bx = c / box_cols, cc = c % box_cols;
      const int chunk = cc / 8;
      const int sw = row_bytes == 128 ? (r & 7) : row_bytes == 64 ? ((r >> 1) & 3) : 0;
      const size_t off = (size_t)bx * rows * row_bytes + (size_t)r * row_bytes + ((chunk ^ sw) << 4) + (cc % 8) * 2;
      const uint16_t u = f2bf(m[(size_t)r * cols + c]);
      memcpy(&img[off], &u, 2);
    }
}

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: umma_probe <variant> [swap]\n"); return 1; }
  const char* var = argv[1];
  const int swap = argc > 2 ? atoi(argv[2]) : 0;
  Params p;
  memset(&p, 0, sizeof(p));
  std::vector<float> A, B;   // logical A [M][K], B as stored: qk: [N][K] (K-major); pv: V [K][N] (rows = keys)
  std::vector<uint8_t> a_img, b_img;
  srand(1234);
  auto rnd = [] { return bf2f(f2bf((float)(rand() % 2001 - 1000) / 1000.f)); };
  bool b_is_kn = false;
  const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10);
  if (!strcmp(var, "qk32")) {
    p.M = 128; p.N = 144; p.K = 32;
    A.resize(128 * 32); B.resize(144 * 32);
    for (auto& x : A) x = rnd();
    for (auto& x : B) x = rnd();
    put(a_img, 128, 32, 64, A);
    put(b_img, 144, 32, 64, B);
    p.a_lbo = 16; p.a_sbo = 512; p.a_swz = 4; p.a_kstep = 32; p.a_kblock_elems = 32; p.a_kblock_bytes = 0;
    p.b_lbo = 16; p.b_sbo = 512; p.b_swz = 4; p.b_kstep = 32; p.b_kblock_elems = 32; p.b_kblock_bytes = 0;
    p.idesc = idesc_base | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
  } else if (!strncmp(var, "pv32", 4) || !strncmp(var, "pv64", 4)) {
    const int hd = var[2] == '3' ? 32 : 64;
    const int keys = hd == 32 ? 144 : 128;
    p.M = 128; p.N = hd; p.K = keys;
    b_is_kn = true;
    A.resize(128 * keys); B.resize(keys * hd);
    for (auto& x : A) x = rnd();
    for (auto& x : B) x = rnd();
    p.a_from_tmem = !strcmp(var + 4, "ts");
    if (p.a_from_tmem) {
      a_img.resize((size_t)128 * (keys / 2) * 4);
      for (int r = 0; r < 128; ++r)
        for (int c = 0; c < keys / 2; ++c) {
          const uint32_t w = (uint32_t)f2bf(A[r * keys + 2 * c]) | ((uint32_t)f2bf(A[r * keys + 2 * c + 1]) << 16);
          memcpy(&a_img[((size_t)r * (keys / 2) + c) * 4], &w, 4);
        }
    } else {
      put(a_img, 128, keys, 128, A);
      p.a_lbo = 16; p.a_sbo = 1024; p.a_swz = 2; p.a_kstep = 32; p.a_kblock_elems = 64; p.a_kblock_bytes = 128 * 128;
    }
    put(b_img, keys, hd, hd * 2, B);   // rows = keys, hd*2 bytes wide, swizzle = row width
    const uint32_t group = 8 * hd * 2;  // bytes of an 8-key group
    p.b_swz = hd == 32 ? 4 : 2;
    p.b_lbo = swap ? group : 16;
    p.b_sbo = swap ? 16 : group;
    p.b_kstep = 2 * group; p.b_kblock_elems = 1 << 20; p.b_kblock_bytes = 0;
    p.idesc = idesc_base | (1u << 16) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
  } else if (!strcmp(var, "m64")) {
    // D[64x80] = Q[64x32] * K[80x32]^T with M = 64: WHERE in TMEM do the 64 rows land, and can the D address carry a lane
    // offset (argv[2])? The answer decides whether two 64-row units can share the 128 lanes of one column range.
    p.M = 64; p.N = 80; p.K = 32;
    A.resize(64 * 32); B.resize(80 * 32);
    for (auto& x : A) x = rnd();
    for (auto& x : B) x = rnd();
    put(a_img, 64, 32, 64, A);
    put(b_img, 80, 32, 64, B);
    p.a_lbo = 16; p.a_sbo = 512; p.a_swz = 4; p.a_kstep = 32; p.a_kblock_elems = 32; p.a_kblock_bytes = 0;
    p.b_lbo = 16; p.b_sbo = 512; p.b_swz = 4; p.b_kstep = 32; p.b_kblock_elems = 32; p.b_kblock_bytes = 0;
    p.idesc = idesc_base | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
    p.d_lane = swap;       // second argument = lane offset here
    p.prefill = 1;
  } else {
    printf("unknown variant %s\n", var);
    return 1;
  }
  p.a_bytes = (int)a_img.size();
  p.b_bytes = (int)b_img.size();
  uint8_t *da, *db;
  float* dout;
  CK(cudaMalloc(&da, a_img.size()));
  CK(cudaMalloc(&db, b_img.size()));
  CK(cudaMalloc(&dout, (size_t)128 * p.N * 4));
  CK(cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice));
  const int smem = 1024 + ((p.a_bytes + 1023) & ~1023) + p.b_bytes + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(da, db, dout, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> out((size_t)128 * p.N);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  if (!strcmp(var, "m64")) {
    // which logical row does every TMEM lane hold?
    printf("variant=m64 d_lane=%d: lane <- row ('.' = untouched sentinel, '?' = something else)\n", p.d_lane);
    for (int l = 0; l < 128; ++l) {
      int found = -2;
      bool sentinel = true;
      for (int n = 0; n < p.N; ++n) sentinel = sentinel && out[(size_t)l * p.N + n] == 12345.f;
      if (sentinel) found = -1;
      else
        for (int r = 0; r < 64 && found == -2; ++r) {
          double e = 0;
          for (int n = 0; n < p.N; ++n) {
            double ref = 0;
            for (int k = 0; k < p.K; ++k) ref += (double)A[r * p.K + k] * B[n * p.K + k];
            e = fmax(e, fabs(ref - out[(size_t)l * p.N + n]));
          }
          if (e < 1e-2) found = r;
        }
      if (found == -1) printf("  .");
      else if (found == -2) printf("  ?");
      else printf(" %2d", found);
      if ((l & 31) == 31) printf("   | lanes %d-%d\n", l - 31, l);
    }
    return 0;
  }
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < p.N; ++n) {
      double ref = 0;
      for (int k = 0; k < p.K; ++k) ref += (double)A[m * p.K + k] * (b_is_kn ? B[k * p.N + n] : B[n * p.K + k]);
      maxerr = fmax(maxerr, fabs(ref - out[(size_t)m * p.N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("variant=%s swap=%d M=%d N=%d K=%d max_abs_err=%.5f (max |ref| %.3f) %s\n", var, swap, p.M, p.N, p.K, maxerr, maxref,
         maxerr < 1e-2 ? "PASS" : "FAIL");
  return maxerr < 1e-2 ? 0 : 3;
}
