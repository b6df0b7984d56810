// On-device piece preparation: LAB uint8 puzzle image -> [N, 3, S, S] fp32 model input.
//
// The reference prepares BOTH pieces of every pair inside DataLoader workers (data/datasets/pieces_dataset.py:34-56 +
// data/transforms.py:12-26): cv2.cvtColor(LAB2RGB) on the eroded piece, PIL bilinear Resize(S), ToTensor (/255),
// Normalize((x - .5) / .5) -- 2*N*(N-1) times; it is deterministic per piece. Here: one CTA per piece, everything in
// shared memory, the integer arithmetic of both libraries restated so that every output float is bit-identical:
//   1. crop: piece (r, c) of the centred grid, eroded (paikin_tal_solver/puzzle_importer.py:196-232, :430-446);
//   2. Lab -> sRGB: OpenCV's bit-exact 8-bit path (fixed point 2^14, f^-1 by integer division / cube, 3x3 integer
//      matrix, 4096-entry inverse-gamma table; tables in lab_tables.inc, verified on all 2^24 triples);
//   3. resize: Pillow's two-pass 8-bit resampler (horizontal then vertical, 22-bit fixed-point coefficients computed
//      in double on the host exactly as precompute_coeffs does, uint8 rounding between the passes);
//   4. fp32: v / 255 (IEEE division), (x - 0.5) / 0.5.
// HBM-bound byte work: reads s*s*3 bytes and writes 3*S*S*4 bytes per piece.
#include <algorithm>
#include <cmath>
#include <vector>

#include "kernels.h"

namespace vited {
namespace {

#include "lab_tables.inc"

constexpr int kBase = 1 << 14;
constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow Resample.c PRECISION_BITS
constexpr int kMaxSide = 96;                 // eroded piece side limit (shared memory budget)
constexpr int kMaxOut = 96;
constexpr int kMaxTaps = 8;                  // ceil(support) * 2 + 1 with support = max(1, side / out)

__device__ uint16_t g_lab_y[256], g_lab_fy[256];
__device__ __align__(16) uint8_t g_inv_gamma[4096];

struct PrepArgs {
  const uint8_t* lab;   // [H, W, 3]
  int H, W;
  int piece_width, rows, cols, top, left, side, off, out;
  const int* coef;      // [out, taps] fixed-point weights
  const int* bounds;    // [out, 2] first source index, tap count
  int taps;
  float* dst;           // [rows*cols, 3, out, out]
};

__device__ __forceinline__ int ab_to_xz(int i) {
  // OpenCV abToXZ_b: f^-1 in fixed point; C integer division truncates toward zero, which matters for i < 0
  return i <= 3390 ? i * 108 / 841 - kBase * 16 / 116 * 108 / 841 : i * i / kBase * i / kBase;
}

__device__ __forceinline__ void lab_to_rgb(int L, int a, int b, const uint8_t* gamma, uint8_t* rgb) {
  const int y = g_lab_y[L], fy = g_lab_fy[L];
  const int adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * kBase / 500;
  const int bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * kBase / 200 + 1;
  const int x = ab_to_xz(fy + adiv), z = ab_to_xz(fy - bdiv);
  // round(4096 * XYZ->sRGB(D65) * white point), descaled by 14 bits to the 12-bit index of the inverse-gamma table
  const int r = (12615 * x - 6296 * y - 2223 * z + (1 << 13)) >> 14;
  const int g = (-3773 * x + 7684 * y + 185 * z + (1 << 13)) >> 14;
  const int bl = (217 * x - 836 * y + 4715 * z + (1 << 13)) >> 14;
  rgb[0] = gamma[min(max(r, 0), 4095)];
  rgb[1] = gamma[min(max(g, 0), 4095)];
  rgb[2] = gamma[min(max(bl, 0), 4095)];
}

__device__ __forceinline__ int clip8(int v) { return min(max(v >> kPrecisionBits, 0), 255); }

__global__ void __launch_bounds__(256) prep_pieces_kernel(PrepArgs p) {
  extern __shared__ uint8_t smem[];
  uint8_t* gamma = smem;                                   // 4096
  uint8_t* rgb = gamma + 4096;                             // [side][side][3]
  uint8_t* tmp = rgb + p.side * p.side * 3;                // [side][out][3]   after the horizontal pass
  const int piece = blockIdx.x, pr = piece / p.cols, pc = piece % p.cols;
  const int y0 = p.top + pr * p.piece_width + p.off, x0 = p.left + pc * p.piece_width + p.off;
  for (int i = threadIdx.x; i < 4096 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(gamma)[i] = reinterpret_cast<const uint32_t*>(g_inv_gamma)[i];
  __syncthreads();
  for (int i = threadIdx.x; i < p.side * p.side; i += blockDim.x) {
    const int yy = i / p.side, xx = i % p.side;
    const uint8_t* src = p.lab + ((size_t)(y0 + yy) * p.W + (x0 + xx)) * 3;
    lab_to_rgb(src[0], src[1], src[2], gamma, rgb + i * 3);
  }
  __syncthreads();
  // horizontal pass (ImagingResampleHorizontal_8bpc): every source row, `out` columns
  for (int i = threadIdx.x; i < p.side * p.out; i += blockDim.x) {
    const int yy = i / p.out, xx = i % p.out;
    const int xmin = p.bounds[2 * xx], n = p.bounds[2 * xx + 1];
    const int* k = p.coef + xx * p.taps;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < n; ++t) {
      const uint8_t* px = rgb + (yy * p.side + xmin + t) * 3;
      s0 += px[0] * k[t]; s1 += px[1] * k[t]; s2 += px[2] * k[t];
    }
    uint8_t* o = tmp + i * 3;
    o[0] = (uint8_t)clip8(s0); o[1] = (uint8_t)clip8(s1); o[2] = (uint8_t)clip8(s2);
  }
  __syncthreads();
  // vertical pass (ImagingResampleVertical_8bpc) + ToTensor + Normalize, written channel-major
  float* dst = p.dst + (size_t)piece * 3 * p.out * p.out;
  for (int i = threadIdx.x; i < 3 * p.out * p.out; i += blockDim.x) {
    const int c = i / (p.out * p.out), rem = i % (p.out * p.out), yy = rem / p.out, xx = rem % p.out;
    const int ymin = p.bounds[2 * yy], n = p.bounds[2 * yy + 1];
    const int* k = p.coef + yy * p.taps;
    int s = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < n; ++t) s += tmp[((ymin + t) * p.out + xx) * 3 + c] * k[t];
    const float v = __fdiv_rn((float)clip8(s), 255.0f);               // ToTensor
    dst[i] = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);                      // Normalize(mean .5, std .5)
  }
}

// Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0), box = whole axis
void pil_bilinear_coeffs(int in_size, int out_size, int* taps_out, std::vector<int>* coef, std::vector<int>* bounds) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const int ksize = (int)std::ceil(support) * 2 + 1;
  coef->assign((size_t)out_size * ksize, 0);
  bounds->assign((size_t)out_size * 2, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      double v = (x + xmin - center + 0.5) * ss;
      if (v < 0.0) v = -v;
      const double w = v < 1.0 ? 1.0 - v : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double scaled = k[x] * (double)(1 << kPrecisionBits);
      (*coef)[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + scaled) : (int)(0.5 + scaled);
    }
    (*bounds)[2 * xx] = xmin;
    (*bounds)[2 * xx + 1] = xmax;
  }
  *taps_out = ksize;
}

bool g_tables_ready[64] = {};

int upload_tables() {
  int dev = 0;
  VITED_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && g_tables_ready[dev]) return 0;
  VITED_CUDA_OK(cudaMemcpyToSymbol(g_lab_y, kLabToY, sizeof(kLabToY)));
  VITED_CUDA_OK(cudaMemcpyToSymbol(g_lab_fy, kLabToFy, sizeof(kLabToFy)));
  VITED_CUDA_OK(cudaMemcpyToSymbol(g_inv_gamma, kInvGamma, sizeof(kInvGamma)));
  if (dev < 64) g_tables_ready[dev] = true;
  return 0;
}

// [N, S, S, 3] uint8 -> [N, 3, S, S] fp32: ToTensor (v / 255) + Normalize((x - .5) / .5), the arithmetic of
// hisfrag.py:89-93 after its CenterCrop; four pixels of one channel per thread, float4 stores
__global__ void normalize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, size_t n_quads, int plane) {
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += (size_t)gridDim.x * blockDim.x) {
    const size_t e = q * 4;                       // first output element of this quad: ((n * 3 + c) * plane + pix)
    const size_t pix = e % plane, nc = e / plane, c = nc % 3, n = nc / 3;
    const uint8_t* src = in + (n * plane + pix) * 3 + c;
    float4 v;
    v.x = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[0], 255.0f), 0.5f), 0.5f);
    v.y = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[3], 255.0f), 0.5f), 0.5f);
    v.z = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[6], 255.0f), 0.5f), 0.5f);
    v.w = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[9], 255.0f), 0.5f), 0.5f);
    *reinterpret_cast<float4*>(out + e) = v;
  }
}

}  // namespace

int normalize_u8(const uint8_t* in, int N, int S, float* out, cudaStream_t stream) {
  VITED_CHECK(in != nullptr && out != nullptr && N >= 1 && S >= 1, "normalize_u8: bad arguments N=%d S=%d", N, S);
  VITED_CHECK((S * S) % 4 == 0, "normalize_u8: S*S must be a multiple of 4, got S=%d", S);
  const size_t n_quads = (size_t)N * 3 * S * S / 4;
  const int blocks = (int)std::min<size_t>((n_quads + 255) / 256, (size_t)148 * 16);
  normalize_u8_kernel<<<blocks, 256, 0, stream>>>(in, out, n_quads, S * S);
  VITED_CUDA_OK(cudaGetLastError());
  return 0;
}

int prepare_pieces(const uint8_t* lab, int H, int W, int piece_width, int side, int off, int out_size, float* dst,
                   int* n_pieces, cudaStream_t stream) {
  VITED_CHECK(lab != nullptr && dst != nullptr, "prepare_pieces: null pointer");
  VITED_CHECK(piece_width > 0 && H >= piece_width && W >= piece_width,
              "prepare_pieces: image %dx%d is smaller than one %d-pixel piece", H, W, piece_width);
  VITED_CHECK(side >= 1 && side <= piece_width && side <= kMaxSide && off >= 0 && off + side <= piece_width,
              "prepare_pieces: eroded side %d / offset %d invalid for piece width %d (side limit %d)", side, off,
              piece_width, kMaxSide);
  VITED_CHECK(out_size >= 1 && out_size <= kMaxOut, "prepare_pieces: output size %d out of range (1..%d)", out_size, kMaxOut);
  PrepArgs p;
  p.lab = lab; p.H = H; p.W = W; p.piece_width = piece_width;
  p.cols = W / piece_width; p.rows = H / piece_width;                              // puzzle_importer.py:196-199
  p.top = (H - p.rows * piece_width) / 2; p.left = (W - p.cols * piece_width) / 2;  // :205-212 centred grid
  p.side = side; p.off = off; p.out = out_size; p.dst = dst;
  if (n_pieces != nullptr) *n_pieces = p.rows * p.cols;
  std::vector<int> coef, bounds;
  int taps = 0;
  pil_bilinear_coeffs(side, out_size, &taps, &coef, &bounds);
  VITED_CHECK(taps <= kMaxTaps, "prepare_pieces: %d filter taps (downscale %d -> %d too strong)", taps, side, out_size);
  if (upload_tables()) return 1;
  // the coefficient table rides in a small stream-ordered allocation
  int* d_tab = nullptr;
  const size_t n_coef = coef.size(), n_bounds = bounds.size();
  VITED_CUDA_OK(cudaMallocAsync(&d_tab, (n_coef + n_bounds) * sizeof(int), stream));
  VITED_CUDA_OK(cudaMemcpyAsync(d_tab, coef.data(), n_coef * sizeof(int), cudaMemcpyHostToDevice, stream));
  VITED_CUDA_OK(cudaMemcpyAsync(d_tab + n_coef, bounds.data(), n_bounds * sizeof(int), cudaMemcpyHostToDevice, stream));
  VITED_CUDA_OK(cudaStreamSynchronize(stream));   // coef / bounds are stack-lifetime host vectors (pageable copies)
  p.coef = d_tab; p.bounds = d_tab + n_coef; p.taps = taps;
  const size_t smem = 4096 + (size_t)side * side * 3 + (size_t)side * out_size * 3;
  VITED_CUDA_OK(cudaFuncSetAttribute(prep_pieces_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prep_pieces_kernel<<<p.rows * p.cols, 256, smem, stream>>>(p);
  VITED_CUDA_OK(cudaGetLastError());
  VITED_CUDA_OK(cudaFreeAsync(d_tab, stream));
  return 0;
}

}  // namespace vited
