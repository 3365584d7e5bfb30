// kernels.cu -- hand-written sm_100a kernels of the baseline-JPEG encode path.
//
//   F1  colour convert + 4:2:0 downsample + integer fDCT + quantise, one 8x8 block per thread,
//       everything in registers (fully unrolled; the zig-zag reorder is a compile-time register
//       permutation).  Fast path: one warp per tile, strips of 8 pixel rows x 768 bytes staged
//       into shared memory with bulk-async copies (TMA engine, cp.async.bulk + mbarrier).
//       Generic path: clamped byte loads, handles edges / odd strides / BGRA / RGBA.
//       Planar path: the same shape for YUV420 / YUV444 / NV12 / NV21 / gray planes (sample = pixel - 128).
//   Q1  re-quantise stored raw coefficients (adaptive quantisation, methods >= 3)
//   H1  coefficient histogram (methods >= 3), shared-memory privatised
//   A1  analysis of the histograms -> adapted matrices and quantiser constants (methods >= 3): the
//       reference's float / double arithmetic in the reference's order, two small kernels
//   T1  trellis quantiser (methods 7, 8), one block per thread, working storage in shared-memory
//       columns, blocks handed out sorted by their number of non-zeros
//   S1  Huffman symbol statistics (optimised tables)
//   E   entropy stage in one pass, persistent and warp-specialised: worker warps walk the blocks
//       of a tile (bits + packed words per block) -> scan -> a dedicated warp runs the decoupled
//       look-back over the tiles of the picture while the workers walk the next tile -> bit
//       packing at the resulting offsets (replaces the serial bit writer)
//   S   0xFF byte stuffing in one pass: count -> CTA scan -> decoupled look-back -> scatter,
//       padding and EOI
// Every kernel takes a group of pictures (gridDim.y); see kernels.cuh.
// Reference behaviour each stage reproduces is cited in block_ops.cuh and at each kernel.
#include "kernels.cuh"

#include <cuda_runtime.h>
#include <atomic>

namespace sjb {

namespace {

// -------------------------------------------------------------------------------------------
// small device helpers
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t pack16(int lo, int hi) {   // one PRMT
  return __byte_perm(static_cast<uint32_t>(lo), static_cast<uint32_t>(hi), 0x5410);
}

// block index inside its MCU without a run-time division (mcu_blocks is 6, 3 or 1)
__device__ __forceinline__ int block_in_mcu(size_t g, int mcu_blocks) {
  const unsigned gg = static_cast<unsigned>(g);
  return (mcu_blocks == 6) ? static_cast<int>(gg % 6u) : (mcu_blocks == 3) ? static_cast<int>(gg % 3u) : 0;
}

// CTA-wide exclusive scan of one uint32 per thread (blockDim.x multiple of 32, <= 1024).
// Returns the exclusive prefix; *total receives the CTA sum.  scratch: 33 words of smem.
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* scratch, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < nwarps) ? scratch[lane] : 0;
    uint32_t wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    scratch[lane] = wi - w;            // exclusive warp offsets
    if (lane == 31) scratch[32] = wi;  // grand total
  }
  __syncthreads();
  const uint32_t res = scratch[warp] + incl - v;
  *total = scratch[32];
  __syncthreads();   // scratch reusable afterwards
  return res;
}

// -------------------------------------------------------------------------------------------
// block output: 64 int32 registers -> 128 bytes, four sectors kCoefSectorStride apart
// (sector-interleaved layout, block_ops.cuh); dst = base of the block (coef_block_base)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_block_natural(const int (&v)[64], int16_t* dst) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    d[coef_chunk_index(i)] = make_uint4(pack16(v[8 * i], v[8 * i + 1]), pack16(v[8 * i + 2], v[8 * i + 3]),
                                        pack16(v[8 * i + 4], v[8 * i + 5]), pack16(v[8 * i + 6], v[8 * i + 7]));
  }
}

// quantise (QuantTab) + zig-zag + non-zero chunk bitmap (bit c <=> the 16-byte chunk c != 0).
// quantize.cc:288-320 without the run/level emission, which the entropy kernel redoes from the bitmap.
// Tab supplies the constants of output pair p = zig-zag positions 2p, 2p+1.
struct ParamTab {       // kernel-parameter (constant bank) table, compile-time offsets
  const QuantTab& t;
  __device__ __forceinline__ void pair(int p, int& iq0, int& c0, int& iq1, int& c1) const {
    iq0 = t.e[2 * p][0]; c0 = t.e[2 * p][1]; iq1 = t.e[2 * p + 1][0]; c1 = t.e[2 * p + 1][1];
  }
};
struct SmemTab {        // shared-memory copy, run-time base (lets luma and chroma share code)
  uint32_t addr;
  __device__ __forceinline__ void pair(int p, int& iq0, int& c0, int& iq1, int& c1) const {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(iq0), "=r"(c0), "=r"(iq1), "=r"(c1)
                 : "r"(addr + 16 * p));
  }
};

template <bool kSparse = false, class Tab>
__device__ __forceinline__ void quantize_store_block(const int (&v)[64], const Tab& tab,
                                                     int16_t* dst, uint8_t* chunkmask) {
  constexpr int zz[64] = SJB_ZIGZAG_INIT;
  uint4* d = reinterpret_cast<uint4*>(dst);
  uint32_t mask = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s) {          // one 32-byte sector = two 16-byte chunks
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int p = 8 * s + j;
      int iq0, c0, iq1, c1;
      tab.pair(p, iq0, c0, iq1, c1);
      w[j] = pack16(quantize_coeff(v[zz[2 * p]], iq0, c0), quantize_coeff(v[zz[2 * p + 1]], iq1, c1));
    }
    const bool nz0 = (w[0] | w[1] | w[2] | w[3]) != 0, nz1 = (w[4] | w[5] | w[6] | w[7]) != 0;
    if (nz0) mask |= 1u << (2 * s);
    if (nz1) mask |= 1u << (2 * s + 1);
    // Sector 0 carries the DC and is always stored.  An all-zero sector is never read back (the
    // bitmap gates every load of the entropy stage), so the fast path does not write it; whole
    // sectors are written so that L2 never has to fill a partial one from DRAM.
    if (!kSparse || s == 0 || nz0 || nz1) {
      d[coef_chunk_index(2 * s)] = make_uint4(w[0], w[1], w[2], w[3]);
      d[coef_chunk_index(2 * s + 1)] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
  *chunkmask = static_cast<uint8_t>(mask);
}

// -------------------------------------------------------------------------------------------
// F1 generic: one thread per block, clamped byte loads
// -------------------------------------------------------------------------------------------
struct PixelReader {
  const uint8_t* base;
  long long stride;
  int w1, h1, pstep, ro, bo;   // last valid x / y, bytes per pixel, offsets of R and B
  __device__ __forceinline__ void get(int x, int y, int* r, int* g, int* b) const {
    x = min(x, w1);
    y = min(y, h1);
    const uint8_t* p = base + y * stride + static_cast<long long>(x) * pstep;
    *r = p[ro];
    *g = p[1];
    *b = p[bo];
  }
};

__device__ __forceinline__ void luma_samples(const PixelReader& px, int x0, int y0, int (&v)[64]) {
#pragma unroll
  for (int y = 0; y < 8; ++y) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      int r, g, b;
      px.get(x0 + x, y0 + y, &r, &g, &b);
      v[8 * y + x] = rgb_to_y(r, g, b);
    }
  }
}

// planar sources: sample = pixel - 128, last valid row / column replicated
// (Convert8To16b[Clipped] colors_rgb.cc:1234-1260, GetReplicatedYSamples encoders.cc:138-143)
__device__ __forceinline__ void plane_samples(const uint8_t* p, long long stride, int xstep, int pw1, int ph1,
                                              int x0, int y0, int (&v)[64]) {
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    const uint8_t* row = p + min(y0 + y, ph1) * stride;
#pragma unroll
    for (int x = 0; x < 8; ++x) v[8 * y + x] = static_cast<int>(row[min(x0 + x, pw1) * xstep]) - 128;
  }
}

template <bool kRaw>
__global__ void __launch_bounds__(128)
f1_generic_kernel(const __grid_constant__ FrameSet fs, int mx0, int my0, int mx1, int my1,
                  const __grid_constant__ QuantTabs qt, GroupBuffers gb) {
  const int rect_w = mx1 - mx0;
  const int mcu_blocks = fs.mcu_blocks;
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nb = static_cast<long long>(rect_w) * (my1 - my0) * mcu_blocks;
  if (t >= nb) return;
  const int frame = blockIdx.y;
  const int k = static_cast<int>(t % mcu_blocks);
  const long long m = t / mcu_blocks;
  const int mx = mx0 + static_cast<int>(m % rect_w), my = my0 + static_cast<int>(m / rect_w);
  const size_t g = (static_cast<size_t>(my) * fs.mcus_x + mx) * mcu_blocks + k;
  int16_t* coef = gb.coef + frame * gb.coef_pitch;
  uint8_t* nzmask = gb.nzmask + frame * gb.mask_pitch;

  PixelReader px;
  px.base = fs.pix[frame];
  px.stride = fs.stride;
  px.w1 = fs.width - 1;
  px.h1 = fs.height - 1;
  px.pstep = (fs.pix_fmt == kFmtRGB) ? 3 : 4;
  px.ro = (fs.pix_fmt == kFmtBGRA) ? 2 : 0;
  px.bo = (fs.pix_fmt == kFmtBGRA) ? 0 : 2;

  int v[64];
  int chroma = 0;
  if (fs.yuv_mode == kYuv420) {
    const int X = 16 * mx, Y = 16 * my;
    if (k < 4) {
      // AverageExtraLuma (encoders.cc:107-125): luma blocks wholly outside the picture are
      // flattened to the rounded mean of a neighbour block's samples.
      const int sub_w = fs.width - X, sub_h = fs.height - Y;
      int src = -1;
      if (k == 1 && sub_w <= 8) src = 0;
      if (k >= 2 && sub_h <= 8) src = (sub_w > 8) ? 1 : 0;
      else if (k == 3 && sub_w <= 8) src = 2;
      const int kk = (src < 0) ? k : src;
      if (fs.planar) plane_samples(px.base, px.stride, 1, px.w1, px.h1, X + 8 * (kk & 1), Y + 8 * (kk >> 1), v);
      else luma_samples(px, X + 8 * (kk & 1), Y + 8 * (kk >> 1), v);
      if (src >= 0) {
        int sum = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) sum += v[i];
        const int dc = (sum + 32) >> 6;
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = dc;
      }
    } else if (fs.planar) {
      chroma = 1;
      plane_samples(k == 4 ? fs.pix_u[frame] : fs.pix_v[frame], k == 4 ? fs.stride_u : fs.stride_v, fs.uv_step,
                    ((fs.width + 1) >> 1) - 1, ((fs.height + 1) >> 1) - 1, 8 * mx, 8 * my, v);
    } else {
      chroma = 1;
#pragma unroll
      for (int y = 0; y < 8; ++y) {
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          int sr = 0, sg = 0, sb = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            int r, gg, b;
            px.get(X + 2 * x + (q & 1), Y + 2 * y + (q >> 1), &r, &gg, &b);
            sr += r; sg += gg; sb += b;
          }
          v[8 * y + x] = (k == 4) ? rgb4_to_u(sr, sg, sb) : rgb4_to_v(sr, sg, sb);
        }
      }
    }
  } else if (fs.planar) {
    chroma = (k > 0);
    const uint8_t* base = (k == 0) ? px.base : (k == 1) ? fs.pix_u[frame] : fs.pix_v[frame];
    const long long st = (k == 0) ? px.stride : (k == 1) ? fs.stride_u : fs.stride_v;
    plane_samples(base, st, 1, px.w1, px.h1, 8 * mx, 8 * my, v);
  } else {
    const int X = 8 * mx, Y = 8 * my;
    chroma = (k > 0);
#pragma unroll
    for (int y = 0; y < 8; ++y) {
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        int r, gg, b;
        px.get(X + x, Y + y, &r, &gg, &b);
        v[8 * y + x] = (k == 0) ? rgb_to_y(r, gg, b) : (k == 1) ? rgb_to_u(r, gg, b) : rgb_to_v(r, gg, b);
      }
    }
  }
  fdct64(v);
  if (kRaw) {
    store_block_natural(v, coef + coef_block_base(g));
  } else {
    if (chroma) quantize_store_block(v, ParamTab{qt.m[1]}, coef + coef_block_base(g), nzmask + g);
    else        quantize_store_block(v, ParamTab{qt.m[0]}, coef + coef_block_base(g), nzmask + g);
  }
}

// -------------------------------------------------------------------------------------------
// F1 fast path.  One warp (= one CTA) per tile; every lane owns one 8x8 block at a time, all 64
// samples in registers.  A strip = 8 pixel rows x 256 pixels (32 block columns, 768 bytes per
// row) and is staged into shared memory by 8 bulk-async row copies (TMA engine, cp.async.bulk +
// mbarrier complete_tx).  All copies of a tile are issued up front; the many resident warps per
// SM (16) hide each other's load latency, and the hardware CTA scheduler balances the SMs.
//   4:2:0 : tile = two vertically adjacent strips = 16 MCUs; lane = 2*mcu + half; each lane
//           converts+transforms luma block `half` (top strip) and `2+half` (bottom strip); the
//           2x2-summed chroma goes through a smem exchange (aliased on the consumed top strip),
//           then lane half=0 does the U block and half=1 the V block of its MCU.
//   4:4:4 : tile = one strip = 32 MCUs, lane = MCU, three passes over the staged pixels (Y, U, V).
//   4:0:0 : as 4:4:4, Y only.
// -------------------------------------------------------------------------------------------
enum { kUvMcuBytes = 288 };
// pixel layouts of the fast path: bytes per pixel and the byte offsets of R and B
template <int kFmt> struct PixLayout;
template <> struct PixLayout<kFmtRGB>  { enum { kStep = 3, kR = 0, kB = 2 }; };
template <> struct PixLayout<kFmtRGBA> { enum { kStep = 4, kR = 0, kB = 2 }; };
template <> struct PixLayout<kFmtBGRA> { enum { kStep = 4, kR = 2, kB = 0 }; };

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// The pixels are read exactly once: mark them evict-first in L2 so that the streaming input does
// not push out the coefficients the entropy kernel is about to read back.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  unsigned long long policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

// one block row (8 pixels = 24 or 32 bytes) held in 32-bit words; one PRMT per byte:
// result = source byte (i & 3), upper three bytes zero
#define SJB_BYTE(w, i) __byte_perm((w)[(i) >> 2], 0u, 0x4440u | ((i) & 3))

template <int kStep>
__device__ __forceinline__ void load_block_row(uint32_t addr, uint32_t (&w)[2 * kStep]) {
  if (kStep == 3) {
    // three 64-bit shared loads; lane stride is 24 bytes => conflict-free per half-warp
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(addr));
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+8];" : "=r"(w[2]), "=r"(w[3]) : "r"(addr));
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+16];" : "=r"(w[4]), "=r"(w[5]) : "r"(addr));
  } else {
    // two 128-bit shared loads; lane stride 32 bytes (2-way conflict, the LSU has headroom)
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(addr));
  }
}

// 4:2:0 strip: 8x8 pixels -> 64 luma samples + 4x4 U and V (2x2 sums).  colors_rgb.cc:850-879
template <int kFmt>
__device__ __forceinline__ void convert_strip_420(uint32_t addr, uint32_t row_stride, int (&y)[64], int (&u)[16],
                                                  int (&v)[16]) {
  typedef PixLayout<kFmt> P;
#pragma unroll
  for (int rp = 0; rp < 4; ++rp) {
    uint32_t a[2 * P::kStep], b[2 * P::kStep];
    load_block_row<P::kStep>(addr + (2 * rp) * row_stride, a);
    load_block_row<P::kStep>(addr + (2 * rp + 1) * row_stride, b);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int sr = 0, sg = 0, sb = 0;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int x = 2 * q + e;
        const int r0 = SJB_BYTE(a, P::kStep * x + P::kR), g0 = SJB_BYTE(a, P::kStep * x + 1), b0 = SJB_BYTE(a, P::kStep * x + P::kB);
        const int r1 = SJB_BYTE(b, P::kStep * x + P::kR), g1 = SJB_BYTE(b, P::kStep * x + 1), b1 = SJB_BYTE(b, P::kStep * x + P::kB);
        y[16 * rp + x] = rgb_to_y(r0, g0, b0);
        y[16 * rp + 8 + x] = rgb_to_y(r1, g1, b1);
        sr += r0 + r1; sg += g0 + g1; sb += b0 + b1;
      }
      u[4 * rp + q] = rgb4_to_u(sr, sg, sb);
      v[4 * rp + q] = rgb4_to_v(sr, sg, sb);
    }
  }
}

// 4:4:4 / 4:0:0 strip: one component of 8x8 pixels, (cr*r + cg*g + cb*b + rnd) >> 16 with the
// component's coefficients in registers so that Y, U and V share the code.  colors_rgb.cc:809-848
template <int kFmt>
__device__ __forceinline__ void convert_strip_444(uint32_t addr, uint32_t row_stride, int cr, int cg, int cb, int rnd,
                                                  int (&s)[64]) {
  typedef PixLayout<kFmt> P;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    uint32_t a[2 * P::kStep];
    load_block_row<P::kStep>(addr + r * row_stride, a);
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const int rr = SJB_BYTE(a, P::kStep * x + P::kR), gg = SJB_BYTE(a, P::kStep * x + 1), bb = SJB_BYTE(a, P::kStep * x + P::kB);
      s[8 * r + x] = (cr * rr + cg * gg + cb * bb + rnd) >> 16;
    }
  }
}

template <bool kRaw>
__device__ __forceinline__ void finish_block(int (&v)[64], uint32_t tab_addr, int16_t* coef,
                                             uint8_t* nzmask, uint32_t g) {
  fdct64(v);
  if (kRaw) store_block_natural(v, coef + coef_block_base(g));
  else quantize_store_block<true>(v, SmemTab{tab_addr}, coef + coef_block_base(g), nzmask + g);
}

template <int kMode, bool kRaw, int kFmt>
__global__ void __launch_bounds__(32, (kFmt == kFmtRGB) ? 16 : 12)
f1_fast_kernel(const __grid_constant__ FrameSet fs, int mx_full, int my0,
               const __grid_constant__ QuantTabs qt, GroupBuffers gb) {
  constexpr bool k420 = (kMode == kYuv420);
  constexpr int kStep = PixLayout<kFmt>::kStep;
  constexpr int kStripRowBytes = 32 * 8 * kStep;        // 32 block columns
  constexpr int kStripBytes = 8 * kStripRowBytes;
  constexpr int kMcuBlocks = k420 ? 6 : (kMode == kYuv444 ? 3 : 1);
  constexpr int kMcusPerTile = k420 ? 16 : 32;          // MCUs per tile along x
  constexpr int kStrips = k420 ? 2 : 1;
  __shared__ __align__(128) uint8_t strips[kStrips * kStripBytes];
  __shared__ __align__(8) unsigned long long bars[kStrips];
  __shared__ __align__(16) int32_t qtab[2][64][2];      // zig-zag order {iq, cpos}
  const int lane = threadIdx.x;
  const int frame = blockIdx.y;
  const uint32_t slot0 = smem_addr(strips);
  const uint32_t bar0 = smem_addr(bars);
  const uint32_t tab0 = smem_addr(qtab);
  int16_t* coef = gb.coef + frame * gb.coef_pitch;
  uint8_t* nzmask = gb.nzmask + frame * gb.mask_pitch;

  const int chunks_x = (mx_full + kMcusPerTile - 1) / kMcusPerTile;
  const int cx = blockIdx.x % chunks_x, ry = my0 + blockIdx.x / chunks_x;
  const int mcus = min(kMcusPerTile, mx_full - cx * kMcusPerTile);
  const uint32_t row_bytes = static_cast<uint32_t>(mcus) * (k420 ? 16u : 8u) * kStep;

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kStrips; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int i = 0; i < kStrips; ++i) mbar_expect_tx(bar0 + 8 * i, 8 * row_bytes);
  }
  __syncwarp();
  if (lane < 8 * kStrips) {
    // lane -> (strip, row): pixel row = tile origin + lane
    const long long py = static_cast<long long>(k420 ? 16 : 8) * ry + lane;
    const uint8_t* src = fs.pix[frame] + py * fs.stride + static_cast<long long>(cx) * kStripRowBytes;
    bulk_g2s(slot0 + lane * kStripRowBytes, src, row_bytes, bar0 + 8 * (lane >> 3));
  }
  if (!kRaw) {
    // quantiser constants: kernel parameters -> shared memory (256 ints, 8 per lane)
    const int32_t* q = &qt.m[0].e[0][0];
#pragma unroll
    for (int i = 0; i < 8; ++i) (&qtab[0][0][0])[lane + 32 * i] = q[lane + 32 * i];
  }
  __syncwarp();
  // block indices fit 32 bits (at most 3 * 8192 * 8192 blocks per picture)
  const uint32_t mcu0 = static_cast<uint32_t>(ry) * fs.mcus_x + static_cast<uint32_t>(cx) * kMcusPerTile;
  const uint32_t src = slot0 + lane * (8 * kStep);

  if (k420) {
    const int m = lane >> 1, half = lane & 1;
    const bool active = m < mcus;
    const uint32_t uvm = slot0 + m * kUvMcuBytes;     // exchange buffer aliases the top strip
    // three blocks per lane through ONE copy of the fDCT/quantise code: top luma, bottom luma,
    // then the chroma block (U for half 0, V for half 1)
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
      int x[64];
      if (j < 2) {
        int u[16], v[16];
        mbar_wait(bar0 + 8 * j, 0);
        convert_strip_420<kFmt>(src + j * kStripBytes, kStripRowBytes, x, u, v);
        __syncwarp();                                 // top strip fully consumed before reuse
        // chroma partials: rows interleaved U,V (16 bytes each) inside a 288-byte MCU record
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint32_t a = uvm + (4 * j + r) * 32 + half * 8;
          asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(pack16(u[4 * r], u[4 * r + 1])),
                       "r"(pack16(u[4 * r + 2], u[4 * r + 3])) : "memory");
          asm volatile("st.shared.v2.u32 [%0+16], {%1,%2};" ::"r"(a), "r"(pack16(v[4 * r], v[4 * r + 1])),
                       "r"(pack16(v[4 * r + 2], v[4 * r + 3])) : "memory");
        }
      } else {
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                       : "r"(uvm + r * 32 + half * 16));
          x[8 * r + 0] = static_cast<int16_t>(w0 & 0xffff); x[8 * r + 1] = static_cast<int>(w0) >> 16;
          x[8 * r + 2] = static_cast<int16_t>(w1 & 0xffff); x[8 * r + 3] = static_cast<int>(w1) >> 16;
          x[8 * r + 4] = static_cast<int16_t>(w2 & 0xffff); x[8 * r + 5] = static_cast<int>(w2) >> 16;
          x[8 * r + 6] = static_cast<int16_t>(w3 & 0xffff); x[8 * r + 7] = static_cast<int>(w3) >> 16;
        }
      }
      const uint32_t g = (mcu0 + m) * 6u + ((j < 2) ? 2 * j : 4) + half;
      if (active) finish_block<kRaw>(x, tab0 + ((j < 2) ? 0u : 512u), coef, nzmask, g);
    }
  } else {
    const bool active = lane < mcus;
    mbar_wait(bar0, 0);
    int cr = 19595, cg = 38469, cb = 7471, rnd = 32768 - (128 << 16);
#pragma unroll 1
    for (int c = 0; c < kMcuBlocks; ++c) {
      int x[64];
      convert_strip_444<kFmt>(src, kStripRowBytes, cr, cg, cb, rnd, x);
      if (active) finish_block<kRaw>(x, tab0 + (c ? 512u : 0u), coef, nzmask, (mcu0 + lane) * kMcuBlocks + c);
      // next component: U then V (colors_rgb.cc:809-819)
      if (c == 0) { cr = -11059; cg = -21709; cb = 32768; } else { cr = 32768; cg = -27439; cb = -5329; }
      rnd = 32768;
    }
  }
}

// -------------------------------------------------------------------------------------------
// F1 fast path for PLANAR sources (EncodeYUV420 / YUV444 / NV12 / NV21 / Gray, encoders.cc:256-507,
// and the planes the sharp conversion leaves): no colour conversion, sample = pixel - 128
// (Convert8To16b, colors_rgb.cc:1234-1260).  Same shape as the packed kernel: one warp per tile,
// rows of the tile staged into shared memory by bulk-async copies (one row per lane, one mbarrier
// for the whole tile), every lane owns one 8x8 block at a time with its 64 samples in registers.
//   4:2:0 : tile = 16 MCUs: 16 luma rows x 256 bytes + 8 chroma rows (U | V side by side, or the
//           interleaved plane of NV12 / NV21); lane = 2 * mcu + half does luma blocks half and
//           2 + half, then the U (half 0) or V (half 1) block
//   4:4:4 : tile = 32 MCUs, three 8-row strips, lane = MCU        4:0:0 : the luma strip only
// Needs 16-byte aligned plane bases and strides and an even number of MCUs per tile row (chroma
// rows of 8 bytes per MCU must be multiples of 16 bytes); engine.cu sends the rest to the generic kernel.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void plane_block_from_smem(uint32_t addr, uint32_t row_stride, int (&x)[64]) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    uint32_t w[2];
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(addr + r * row_stride));
#pragma unroll
    for (int c = 0; c < 8; ++c) x[8 * r + c] = static_cast<int>(SJB_BYTE(w, c)) - 128;
  }
}
// one component of an interleaved chroma row: bytes parity, parity + 2, ... of 16 bytes
__device__ __forceinline__ void plane_block_from_smem_interleaved(uint32_t addr, uint32_t row_stride, int parity, int (&x)[64]) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    uint32_t w[4];
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr + r * row_stride));
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t even = SJB_BYTE(w, 2 * c), odd = SJB_BYTE(w, 2 * c + 1);
      x[8 * r + c] = static_cast<int>(parity ? odd : even) - 128;
    }
  }
}

template <int kMode, bool kRaw>
__global__ void __launch_bounds__(32, 16)
f1_planar_kernel(const __grid_constant__ FrameSet fs, int mx_full, int my0,
                 const __grid_constant__ QuantTabs qt, GroupBuffers gb) {
  constexpr bool k420 = (kMode == kYuv420);
  constexpr int kMcusPerTile = k420 ? 16 : 32;
  constexpr int kYRows = k420 ? 16 : 8;
  constexpr int kCBytes = (kMode == kYuv400) ? 16 : (k420 ? 8 * 256 : 2 * 8 * 256);
  __shared__ __align__(128) uint8_t ybuf[kYRows * 256];
  __shared__ __align__(128) uint8_t cbuf[kCBytes];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(16) int32_t qtab[2][64][2];      // zig-zag order {iq, cpos}
  const int lane = threadIdx.x;
  const int frame = blockIdx.y;
  const uint32_t ys = smem_addr(ybuf), cs = smem_addr(cbuf), bar0 = smem_addr(&bar), tab0 = smem_addr(qtab);
  int16_t* coef = gb.coef + frame * gb.coef_pitch;
  uint8_t* nzmask = gb.nzmask + frame * gb.mask_pitch;
  const int chunks_x = (mx_full + kMcusPerTile - 1) / kMcusPerTile;
  const int cx = blockIdx.x % chunks_x, ry = my0 + blockIdx.x / chunks_x;
  const int mcus = min(kMcusPerTile, mx_full - cx * kMcusPerTile);       // even (LaunchF1)
  const bool interleaved = k420 && fs.uv_step == 2;
  const uint32_t y_bytes = static_cast<uint32_t>(mcus) * (k420 ? 16u : 8u);
  const uint32_t c_bytes = (kMode == kYuv400) ? 0u : static_cast<uint32_t>(mcus) * (interleaved ? 16u : 8u);
  const int c_copies = (kMode == kYuv400) ? 0 : (interleaved ? 8 : 16);
  if (lane == 0) {
    mbar_init(bar0, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar0, kYRows * y_bytes + c_copies * c_bytes);
  }
  __syncwarp();
  if (lane < kYRows) {
    const long long py = static_cast<long long>(kYRows) * ry + lane;
    bulk_g2s(ys + lane * 256, fs.pix[frame] + py * fs.stride + static_cast<long long>(cx) * 256, y_bytes, bar0);
  } else if (lane < kYRows + c_copies) {
    const int i = lane - kYRows;                  // chroma copy index
    const long long py = 8LL * ry + (i & 7);
    if (interleaved) {
      const uint8_t* base = (fs.pix_u[frame] < fs.pix_v[frame]) ? fs.pix_u[frame] : fs.pix_v[frame];
      bulk_g2s(cs + (i & 7) * 256, base + py * fs.stride_u + static_cast<long long>(cx) * 256, c_bytes, bar0);
    } else if (i < 8) {
      bulk_g2s(cs + (k420 ? (i & 7) * 256 : (i & 7) * 256), fs.pix_u[frame] + py * fs.stride_u + static_cast<long long>(cx) * (k420 ? 128 : 256),
               c_bytes, bar0);
    } else {
      bulk_g2s(cs + (k420 ? (i & 7) * 256 + 128 : 2048 + (i & 7) * 256),
               fs.pix_v[frame] + py * fs.stride_v + static_cast<long long>(cx) * (k420 ? 128 : 256), c_bytes, bar0);
    }
  }
  if (!kRaw) {
    const int32_t* q = &qt.m[0].e[0][0];
#pragma unroll
    for (int i = 0; i < 8; ++i) (&qtab[0][0][0])[lane + 32 * i] = q[lane + 32 * i];
  }
  __syncwarp();
  const uint32_t mcu0 = static_cast<uint32_t>(ry) * fs.mcus_x + static_cast<uint32_t>(cx) * kMcusPerTile;
  mbar_wait(bar0, 0);
  if (k420) {
    const int m = lane >> 1, half = lane & 1;
    const bool active = m < mcus;
    const int u_parity = (fs.pix_u[frame] < fs.pix_v[frame]) ? 0 : 1;
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
      int x[64];
      if (j < 2) {
        plane_block_from_smem(ys + (8 * j) * 256 + m * 16 + half * 8, 256, x);
      } else if (interleaved) {
        plane_block_from_smem_interleaved(cs + m * 16, 256, half ? (1 - u_parity) : u_parity, x);
      } else {
        plane_block_from_smem(cs + half * 128 + m * 8, 256, x);
      }
      const uint32_t g = (mcu0 + m) * 6u + ((j < 2) ? 2 * j : 4) + half;
      if (active) finish_block<kRaw>(x, tab0 + ((j < 2) ? 0u : 512u), coef, nzmask, g);
    }
  } else {
    constexpr int kMcuBlocks = (kMode == kYuv444) ? 3 : 1;
    const bool active = lane < mcus;
#pragma unroll 1
    for (int c = 0; c < kMcuBlocks; ++c) {
      int x[64];
      plane_block_from_smem((c == 0 ? ys : cs + (c - 1) * 2048) + lane * 8, 256, x);
      if (active) finish_block<kRaw>(x, tab0 + (c ? 512u : 0u), coef, nzmask, (mcu0 + lane) * kMcuBlocks + c);
    }
  }
}

// -------------------------------------------------------------------------------------------
// Q1: quantise stored raw coefficients in place (per-picture tables from global memory)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_block_natural(const int16_t* src, int (&v)[64]) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 w = s[coef_chunk_index(i)];
    v[8 * i + 0] = static_cast<int16_t>(w.x & 0xffff); v[8 * i + 1] = static_cast<int>(w.x) >> 16;
    v[8 * i + 2] = static_cast<int16_t>(w.y & 0xffff); v[8 * i + 3] = static_cast<int>(w.y) >> 16;
    v[8 * i + 4] = static_cast<int16_t>(w.z & 0xffff); v[8 * i + 5] = static_cast<int>(w.z) >> 16;
    v[8 * i + 6] = static_cast<int16_t>(w.w & 0xffff); v[8 * i + 7] = static_cast<int>(w.w) >> 16;
  }
}

__global__ void __launch_bounds__(128)
requantize_kernel(const __grid_constant__ FrameSet fs, GroupBuffers gb, const int16_t* __restrict__ raw_src) {
  __shared__ __align__(16) int32_t qtab[2][64][2];
  const int frame = blockIdx.y;
  {
    const int32_t* q = &gb.qtabs[frame].m[0].e[0][0];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) (&qtab[0][0][0])[i] = q[i];
  }
  __syncthreads();
  const size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (g >= fs.blocks_per_frame) return;
  int16_t* blk = gb.coef + frame * gb.coef_pitch + coef_block_base(g);
  int v[64];
  load_block_natural(raw_src ? raw_src + frame * gb.coef_pitch + coef_block_base(g) : blk, v);
  const uint32_t tab = smem_addr(qtab) + ((static_cast<int>(g % fs.mcu_blocks) >= fs.luma_blocks) ? 512u : 0u);
  // all-zero sectors are not written back (as in the fast F1 path: the bitmap gates every later read; in place,
  // such a sector keeps its raw values, which nobody looks at): a photographic picture writes a quarter of the bytes
  quantize_store_block<true>(v, SmemTab{tab}, blk, gb.nzmask + frame * gb.mask_pitch + g);
}

// -------------------------------------------------------------------------------------------
// Quantisation error of stored raw coefficients for the PSNR search (quantize.cc:547-559,
// dichotomy.cc:309-323): sum over all coefficients of ((|c| >> 4) - Q * quantise(|c|))^2 with the
// matrices gb.quant[frame] / tables gb.qtabs[frame]; one 64-bit sum per picture.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
quant_error_kernel(const __grid_constant__ FrameSet fs, GroupBuffers gb, const int16_t* __restrict__ raw,
                   unsigned long long* __restrict__ err_out) {
  __shared__ int32_t qtab[2][64][2];     // zig-zag order {iq, cpos}
  __shared__ uint8_t qm[2][64];          // natural order
  __shared__ unsigned long long warp_sums[8];
  const int frame = blockIdx.y;
  {
    const int32_t* q = &gb.qtabs[frame].m[0].e[0][0];
    const uint8_t* quant = gb.quant + frame * 128;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) (&qtab[0][0][0])[i] = q[i];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) qm[i >> 6][i & 63] = quant[i];
  }
  __syncthreads();
  constexpr int zz[64] = SJB_ZIGZAG_INIT;
  unsigned long long sum = 0;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < fs.blocks_per_frame; g += stride) {
    const int c = (static_cast<int>(g % fs.mcu_blocks) >= fs.luma_blocks) ? 1 : 0;
    int v[64];
    load_block_natural(raw + frame * gb.coef_pitch + coef_block_base(g), v);
    uint32_t e = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const int j = zz[i];
      const int a = abs(v[j]);
      const uint32_t q = static_cast<uint32_t>(qm[c][j]) * static_cast<uint32_t>((a * qtab[c][i][0] + qtab[c][i][1]) >> 20);
      const uint32_t d = static_cast<uint32_t>(a >> 4) - q;
      e += d * d;
    }
    sum += e;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += warp_sums[w];
    atomicAdd(&err_out[frame], t);
  }
}

// -------------------------------------------------------------------------------------------
// H1: histogram of |coef| >> 2 per matrix and position (histogram.cc:99-108).  Each CTA keeps a
// private copy of the 2 x 64 x 128 bins in shared memory and flushes the used ones with atomicAdd.
// A warp takes one block at a time, lane l owns natural positions 2l and 2l+1 (one 32-bit load;
// the 32 lanes read the block's four 32-byte sectors), so the 32 shared-memory atomics of an
// instruction go to 32 different rows of the table; rows are 129 words apart (as in the global
// array), which spreads equal bin values of different rows over the banks.  (With 8 lanes per
// block and rows 128 words apart the bank depended on the bin value only, and the small values that
// dominate every histogram serialised the warp: 121 us for an 8K picture.)
// -------------------------------------------------------------------------------------------
#ifndef SJB_H1_THREADS
#define SJB_H1_THREADS 256
#endif
enum { kH1Threads = SJB_H1_THREADS, kH1Stride = 129, kH1SmemBytes = 2 * 64 * kH1Stride * 4 };
__global__ void __launch_bounds__(kH1Threads)
histogram_kernel(const __grid_constant__ FrameSet fs, GroupBuffers gb) {
  extern __shared__ int32_t hist[];   // [2][64][129]
  for (int i = threadIdx.x; i < 2 * 64 * kH1Stride; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int frame = blockIdx.y;
  const int16_t* raw = gb.coef + frame * gb.coef_pitch;
  const uint32_t nb_blocks = fs.blocks_per_frame;
  const int lane = threadIdx.x & 31;
  // int16 offset of this lane's word inside a block: positions 2l, 2l+1 (sector-interleaved layout)
  const int word_off = coef_pos_offset(2 * lane);
  int32_t* row = hist + (2 * lane) * kH1Stride;
  // Eight consecutive blocks per warp iteration (g0 a multiple of 8): eight independent loads in
  // flight per lane, at constant offsets from one base -- blocks g0..g0+3 and g0+4..g0+7 are two
  // groups of four whose sectors share 128-byte lines (every line fetched is used in full) -- and the
  // luma / chroma class of the eight blocks comes from one shift of a periodic bit pattern (blocks
  // are stored MCU by MCU: chroma is k >= 4 of 6, k >= 1 of 3, never of 1).  The loop is issue-bound:
  // 13 instructions per block and lane; the first version spent 50 on bounds checks and divisions.
  const uint32_t warps_total = gridDim.x * (kH1Threads / 32);
  const uint32_t pattern = (fs.mcu_blocks == 6) ? 0x30C30u : (fs.mcu_blocks == 3) ? 0xDB6DB6u : 0u;
  const uint32_t full_end = nb_blocks & ~7u;
  const uint32_t step = warps_total * 8;
  uint32_t g0 = (blockIdx.x * (kH1Threads / 32) + (threadIdx.x >> 5)) * 8;
  uint32_t k0 = g0 % static_cast<uint32_t>(fs.mcu_blocks), kstep = step % static_cast<uint32_t>(fs.mcu_blocks);
  const int16_t* base = raw + word_off;
  int32_t* row_c = row + 64 * kH1Stride;
  // What bounds the loop is the shared-memory atomic unit, per ATOMS INSTRUCTION rather than per
  // lane: 140 us per 16 4K pictures with 256, 512 or 640 threads per CTA alike, the same for the
  // sparse and the noisy picture, and SLOWER (174 us) when the lanes whose value falls into bins 0..3
  // -- half of them on a photographic picture -- were counted in registers instead and predicated off
  // the atomic: the instruction count went up and the number of ATOMS instructions stayed.  Bank
  // conflicts are not it either: with the rows reordered so that the 32 lanes of an instruction own 32
  // consecutive rows (32 different banks for equal bins; rows 2l put lanes l and l + 16 on one bank) the
  // kernel took 145.4 us against 144.4.  Nor the flavour of the instruction: ptxas turns atomicAdd(.., 1) into
  // ATOMS.POPC.INC (lanes with equal addresses are counted in one go); forced to plain ATOMS.ADD (an increment the
  // compiler cannot see is 1) the kernel took 177 us.
  for (; g0 < full_end; g0 += step) {
    const int16_t* p = base + (static_cast<size_t>(g0 >> 2) << 8);
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = *reinterpret_cast<const uint32_t*>(p + ((e >> 2) << 8) + ((e & 3) << 4));
    const uint32_t cls = pattern >> k0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int a0 = abs(static_cast<int>(static_cast<int16_t>(w[e] & 0xffffu))) >> 2;
      const int a1 = abs(static_cast<int>(w[e]) >> 16) >> 2;
      int32_t* r = ((cls >> e) & 1u) ? row_c : row;
      if (a0 < 128) atomicAdd(&r[a0], 1);
      if (a1 < 128) atomicAdd(&r[kH1Stride + a1], 1);
    }
    k0 += kstep;
    if (k0 >= static_cast<uint32_t>(fs.mcu_blocks)) k0 -= fs.mcu_blocks;
  }
  if (g0 == full_end && g0 < nb_blocks) {                    // the last, partial group of eight
    for (uint32_t g = g0; g < nb_blocks; ++g) {
      const uint32_t wv = *reinterpret_cast<const uint32_t*>(raw + coef_block_base(g) + word_off);
      const int a0 = abs(static_cast<int>(static_cast<int16_t>(wv & 0xffffu))) >> 2;
      const int a1 = abs(static_cast<int>(wv) >> 16) >> 2;
      int32_t* r = (block_in_mcu(g, fs.mcu_blocks) >= fs.luma_blocks) ? row_c : row;
      if (a0 < 128) atomicAdd(&r[a0], 1);
      if (a1 < 128) atomicAdd(&r[kH1Stride + a1], 1);
    }
  }
  __syncthreads();
  int32_t* counts = gb.hist + static_cast<size_t>(frame) * 2 * 64 * 129;
  for (int i = threadIdx.x; i < 2 * 64 * kH1Stride; i += blockDim.x) {
    const int c = hist[i];
    if (c) atomicAdd(&counts[i], c);
  }
}

// -------------------------------------------------------------------------------------------
// A1: the histogram analysis of the adaptive methods on the device (block_ops.cuh: aq_*).  Until
// round 2 the histograms travelled to the host (66 KB per picture), were analysed there (0.05-0.12 ms
// per 4K picture on a pool of threads) and the new quantiser tables travelled back, with the stream
// drained in between; now the matrices never leave the device before the headers need them.
//   analyse_fit_kernel   one CTA per (position, matrix, picture): 128 threads = 25 candidate steps x 4
//                        quarters of the bins add up count x bits and count x error^2 (integers:
//                        exact in any order); thread 0 runs the position's weighted fits in
//                        candidate order, as the reference does
//   analyse_pick_kernel  one CTA per (matrix, picture): thread 0 adds the positions' terms in
//                        ascending order (double sums: the order is part of the result), lambda,
//                        then one thread per position picks its step, clamps it and derives the
//                        quantiser constants (FinalizeQuantizer)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
analyse_fit_kernel(GroupBuffers gb, const __grid_constant__ AqParams ap, AqFit* __restrict__ fit, int* __restrict__ fail) {
  __shared__ int32_t h[kAqBins];
  __shared__ long long part_bits[4][kAqNumDelta], part_dist[4][kAqNumDelta];
  __shared__ int red_total[4], red_last[4];
  const int pos = blockIdx.x, idx = blockIdx.y, frame = blockIdx.z;
  AqFit* out = fit + (static_cast<size_t>(frame) * 2 + idx) * 64 + pos;
  if (pos == 0 || pos == 1 || pos == 8) {          // never touched (histogram.cc:139)
    if (threadIdx.x == 0) {
      out->skip = 1;
      if (pos == 0 && idx == 0) fail[frame] = 0;   // the pick kernel (next launch) raises it
    }
    return;
  }
  const int32_t* row = gb.hist + (static_cast<size_t>(frame) * 2 + idx) * 64 * 129 + pos * 129;
  const int c = row[threadIdx.x];
  h[threadIdx.x] = c;
  // total count and index of the last used bin + 1
  int total = c, last = c ? static_cast<int>(threadIdx.x) + 1 : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, d);
    last = max(last, __shfl_xor_sync(0xffffffffu, last, d));
  }
  if ((threadIdx.x & 31) == 0) {
    red_total[threadIdx.x >> 5] = total;
    red_last[threadIdx.x >> 5] = last;
  }
  __syncthreads();
  total = red_total[0] + red_total[1] + red_total[2] + red_total[3];
  last = max(max(red_last[0], red_last[1]), max(red_last[2], red_last[3]));
  if (total < 0.5 * last) {                         // too sparse to fit (kDensity)
    if (threadIdx.x == 0) out->skip = 1;
    return;
  }
  const int q0 = ap.quant0[idx][pos], qmin = ap.min_quant[idx][pos];
  const int delta_top = ap.qdelta_max[idx] - kAqDeltaMin;
  const int d = threadIdx.x & 31, quarter = threadIdx.x >> 5;
  if (d < kAqNumDelta) {
    long long bs = 0, ds = 0;
    if (aq_candidate_used(q0, qmin, delta_top, d)) {
      aq_bin_sums(h, quarter * 32, min(last, quarter * 32 + 32), q0 + d + kAqDeltaMin, &bs, &ds);
    }
    part_bits[quarter][d] = bs;
    part_dist[quarter][d] = ds;
  }
  __syncthreads();
  if (threadIdx.x < kAqNumDelta) {
    part_bits[0][threadIdx.x] += part_bits[1][threadIdx.x] + part_bits[2][threadIdx.x] + part_bits[3][threadIdx.x];
    part_dist[0][threadIdx.x] += part_dist[1][threadIdx.x] + part_dist[2][threadIdx.x] + part_dist[3][threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x == 0) aq_fit_position(part_bits[0], part_dist[0], q0, qmin, delta_top, out);
}

__global__ void __launch_bounds__(64)
analyse_pick_kernel(GroupBuffers gb, const __grid_constant__ AqParams ap, const AqFit* __restrict__ fit, int* __restrict__ fail) {
  __shared__ double lambda_sh;
  __shared__ uint8_t q_sh[64];
  const int idx = blockIdx.x, frame = blockIdx.y, pos = threadIdx.x;
  const AqFit* f = fit + (static_cast<size_t>(frame) * 2 + idx) * 64;
  // every thread fetches its position's terms (one round trip for all 64), thread 0 adds them in order
  __shared__ double cov_sh[64], den_sh[64];
  __shared__ int skip_sh[64];
  const int skip = f[pos].skip;
  skip_sh[pos] = skip;
  cov_sh[pos] = skip ? 0. : f[pos].cov;
  den_sh[pos] = skip ? 0. : f[pos].den;
  __syncthreads();
  if (pos == 0) {
    double num = 0., den = 0.;
    for (int p = 0; p < 64; ++p) {
      if (skip_sh[p]) continue;
      num += cov_sh[p];
      den += den_sh[p];
    }
    lambda_sh = aq_lambda(num, den);
  }
  __syncthreads();
  const int delta_top = ap.qdelta_max[idx] - kAqDeltaMin;
  int q = ap.quant0[idx][pos];
  if (!skip) {
    AqFit mine;                                    // all 50 loads in flight at once, then the serial pick
#pragma unroll
    for (int d = 0; d < kAqNumDelta; ++d) {
      mine.rate[d] = f[pos].rate[d];
      mine.dist[d] = f[pos].dist[d];
    }
    q = static_cast<uint8_t>(q + aq_best_delta(mine, delta_top, lambda_sh));
  }
  if (q < ap.min_quant[idx][pos]) q = ap.min_quant[idx][pos];      // FinalizeQuantizer's clamp
  q_sh[pos] = static_cast<uint8_t>(q);
  gb.quant[static_cast<size_t>(frame) * 128 + idx * 64 + pos] = static_cast<uint8_t>(q);
  __syncthreads();
  // quantiser constants by ZIG-ZAG position: thread z
  constexpr int zz[64] = SJB_ZIGZAG_INIT;
  const int z = threadIdx.x, i = zz[z];
  int32_t iq, cpos;
  const bool ok = aq_finalize_entry(q_sh[i], i == 0, ap.q_bias, &iq, &cpos);
  QuantTabs* qt = gb.qtabs + frame;
  qt->m[idx].e[z][0] = iq;
  qt->m[idx].e[z][1] = cpos;
  if (ap.nb_comps == 1) {                       // grayscale: the chroma table is never used; keep it defined
    qt->m[1].e[z][0] = iq;
    qt->m[1].e[z][1] = cpos;
    gb.quant[static_cast<size_t>(frame) * 128 + 64 + pos] = ap.quant0[1][pos];
  }
  if (!ok) fail[frame] = 1;
}

// -------------------------------------------------------------------------------------------
// Entropy stage helpers
// -------------------------------------------------------------------------------------------
// quantised DC of the previous block of the same component in scan order (enc.cc:286-305:
// MCU raster order, blocks interleaved; predictors reset once per image, entropy.cc:155-159)
__device__ __forceinline__ int dc_predictor(const int16_t* zz, size_t g, int k, int mcu_blocks, int luma_blocks,
                                            const int* init) {
  size_t prev;
  if (k < luma_blocks) {
    if (k > 0) prev = g - 1;
    else if (g == 0) return init ? init[0] : 0;
    else prev = g - mcu_blocks + luma_blocks - 1;
  } else {
    if (g < static_cast<size_t>(mcu_blocks)) return init ? init[1 + k - luma_blocks] : 0;
    prev = g - mcu_blocks;
  }
  return zz[coef_block_base(prev)];
}

struct ChunkLoader {   // chunk c of a block = zig-zag positions 8c..8c+7, one 16-byte load
  const int16_t* p;    // base of the block (coef_block_base)
  __device__ __forceinline__ Words4 operator()(int c) const {
    const uint4 q = reinterpret_cast<const uint4*>(p)[coef_chunk_index(c)];
    Words4 r;
    r.w[0] = q.x; r.w[1] = q.y; r.w[2] = q.z; r.w[3] = q.w;
    return r;
  }
  __device__ __forceinline__ int value(int pos) const { return p[coef_pos_offset(pos)]; }
};
// Same, with the first two chunks (one 32-byte sector: where the low frequencies live) fetched
// eagerly together with the bitmap and the predictor, so that a typical block needs a single
// round trip to memory instead of a dependent chain.  The block's address is held as an opaque
// 64-bit register and dereferenced with explicit ld.global: left to itself, ptxas (capped at 40
// registers here) rematerialises the whole address chain -- frame pitch, tile, block -- for every
// coefficient fetched in the walk.
struct PrefetchedChunkLoader {
  unsigned long long addr;    // global address of the block's sector 0 (coef_block_base)
  uint4 c0, c1;
  __device__ __forceinline__ explicit PrefetchedChunkLoader(const int16_t* p) {
    addr = static_cast<unsigned long long>(__cvta_generic_to_global(p));
    asm volatile("" : "+l"(addr));
    c0 = chunk(0);
    c1 = chunk(1);
  }
  __device__ __forceinline__ uint4 chunk(int c) const {
    uint4 q;
    asm("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(addr + 16ull * coef_chunk_index(c)));
    return q;
  }
  __device__ __forceinline__ Words4 operator()(int c) const {
    const uint4 q = (c == 0) ? c0 : (c == 1) ? c1 : chunk(c);
    Words4 r;
    r.w[0] = q.x; r.w[1] = q.y; r.w[2] = q.z; r.w[3] = q.w;
    return r;
  }
  // sector 0 of the block is in L1 by now (fetched up front); the later sectors sit in other lines
  __device__ __forceinline__ int value(int pos) const {
    short v;
    asm("ld.global.s16 %0, [%1];" : "=h"(v) : "l"(addr + 2ull * coef_pos_offset(pos)));
    return v;
  }
};

// The same opaque-address loads without the eager fetch (busy tiles: every chunk is visited anyway)
struct AddrChunkLoader {
  unsigned long long addr;
  __device__ __forceinline__ explicit AddrChunkLoader(const int16_t* p) {
    addr = static_cast<unsigned long long>(__cvta_generic_to_global(p));
    asm volatile("" : "+l"(addr));
  }
  __device__ __forceinline__ Words4 operator()(int c) const {
    Words4 r;
    asm("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "l"(addr + 16ull * coef_chunk_index(c)));
    return r;
  }
  __device__ __forceinline__ int value(int pos) const {
    short v;
    asm("ld.global.s16 %0, [%1];" : "=h"(v) : "l"(addr + 2ull * coef_pos_offset(pos)));
    return v;
  }
};

__device__ __forceinline__ void load_code_tables(const CodeTabs* tabs, CodeTabs* sh) {
  // 2176 bytes = 136 x 16 bytes: one vector copy per thread
  static_assert(sizeof(CodeTabs) % 16 == 0, "CodeTabs must be a multiple of 16 bytes");
  const uint4* src = reinterpret_cast<const uint4*>(tabs);
  uint4* dst = reinterpret_cast<uint4*>(sh);
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(CodeTabs) / 16); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// Decoupled look-back (single-pass chained scan).  One 64-bit descriptor per tile:
// [63:62] state (0 = not ready, 1 = tile aggregate, 2 = inclusive prefix), [61:0] value; the
// value travels in the same word as the flag, so no fence is needed.  Executed by one full warp;
// returns the exclusive prefix of `tile` in every lane and publishes the inclusive one.
// Tiles are dispatched in blockIdx order, so a predecessor is always running or finished.
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// One warp inspects kLookWindow predecessors per round trip to L2, kLookPerLane per lane (lane l
// owns the predecessors at distance l*kLookPerLane + 1 .., nearest first).  Measured on B200 (4K,
// 4 pictures per launch): windows of 128 and 256 are SLOWER than 32 (entropy+stuffing 81 / 110 us
// against 60 us) -- the polling traffic of the wider windows costs more than the shorter walk
// saves -- so one descriptor per lane is the default.
#ifndef SJB_LOOK_PER_LANE
#define SJB_LOOK_PER_LANE 1
#endif
enum { kLookPerLane = SJB_LOOK_PER_LANE, kLookWindow = 32 * kLookPerLane };
// first half of the look-back, callable early: makes the tile's aggregate visible to its successors
__device__ __forceinline__ void lookback_publish(unsigned long long* state, long long tile, unsigned long long aggregate) {
  st_volatile_u64(&state[tile], ((tile == 0 ? 2ull : 1ull) << 62) | aggregate);
}
template <bool kPublished = false>
__device__ __forceinline__ unsigned long long warp_lookback(unsigned long long* state, long long tile,
                                                            unsigned long long aggregate) {
  const int lane = threadIdx.x & 31;
  const unsigned long long kValue = (1ull << 62) - 1;
  if (tile == 0) {
    if (!kPublished && lane == 0) st_volatile_u64(&state[0], (2ull << 62) | aggregate);
    return 0;
  }
  if (!kPublished && lane == 0) st_volatile_u64(&state[tile], (1ull << 62) | aggregate);
  unsigned long long prefix = 0;
  long long base = tile - 1;
  while (true) {
    unsigned long long v[kLookPerLane];
#pragma unroll
    for (int e = 0; e < kLookPerLane; ++e) {
      const long long j = base - (lane * kLookPerLane + e);
      v[e] = (j >= 0) ? ld_volatile_u64(&state[j]) : (2ull << 62);   // virtual 0 before tile 0
    }
    // within the lane, nearest first: sum up to and including the first inclusive prefix
    unsigned long long sum = 0;
    bool ready = true, closed = false;      // closed: an inclusive prefix ends the walk here
#pragma unroll
    for (int e = 0; e < kLookPerLane; ++e) {
      const unsigned flag = static_cast<unsigned>(v[e] >> 62);
      if (!closed) {
        ready = ready && (flag != 0);
        sum += v[e] & kValue;
        closed = (flag == 2);
      }
    }
    const unsigned closed_lanes = __ballot_sync(0xffffffffu, closed);
    const int first = closed_lanes ? (__ffs(closed_lanes) - 1) : 31;     // last lane that contributes
    const unsigned need = (first == 31) ? 0xffffffffu : ((2u << first) - 1u);
    const unsigned not_ready = __ballot_sync(0xffffffffu, !ready);
    if (not_ready & need) continue;                                      // spin: re-read the window
    unsigned long long val = (lane <= first) ? sum : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
    prefix += val;
    if (closed_lanes) break;
    base -= kLookWindow;
  }
  if (lane == 0) st_volatile_u64(&state[tile], (2ull << 62) | (prefix + aggregate));
  return prefix;
}

// -------------------------------------------------------------------------------------------
// E: entropy-code lengths + prefix + bit packing, one 8x8 block per thread, 256 blocks per CTA
// (entropy.cc:161-198 CodeBlock; bit_writer.h:201-209 PutBits without the serial accumulator)
// -------------------------------------------------------------------------------------------
struct StreamOut {
  uint32_t* words;
  __device__ __forceinline__ void or_word(uint64_t i, uint32_t v) { if (v) atomicOr(&words[i], v); }
  __device__ __forceinline__ void set_word(uint64_t i, uint32_t v) { words[i] = v; }
};

// Per-thread sink of pass 1: counts the bits and, as long as they fit, keeps the packed words of
// the block (bit 0 of the block = MSB of word 0) in a private shared-memory slot.
#ifndef SJB_LOCAL_WORDS
#define SJB_LOCAL_WORDS 16
#endif
enum { kLocalWords = SJB_LOCAL_WORDS };   // 32-bit words per block kept (512 bits); longer blocks are re-walked
struct LocalSink {
  uint32_t* w;                        // slot of kLocalWords words (+ the word behind them, rewritten by the caller afterwards)
  uint32_t hi, lo;                    // pending bits, top aligned in hi:lo
  int n, nw;                          // pending bit count (< 32 between calls), words emitted
  uint32_t total;
  // 32-bit formulation (a 64-bit accumulator shifted by a variable amount cost 19 instructions per
  // symbol): the symbol is first aligned to the top of a word, then spread over hi:lo by n
  __device__ __forceinline__ void put(uint32_t bits, int len) {     // 1 <= len <= 27
    total += static_cast<uint32_t>(len);
    const uint32_t top = bits << (32 - len);
    hi |= top >> n;
    lo |= __funnelshift_r(0u, top, n);          // top << (32 - n); 0 for n == 0
    n += len;
    if (n >= 32) {
      w[min(nw, static_cast<int>(kLocalWords))] = hi;   // words past the slot land on the spare word
      ++nw;
      hi = lo;
      lo = 0;
      n -= 32;
    }
  }
  __device__ __forceinline__ void finish() {
    if (n > 0) {
      w[min(nw, static_cast<int>(kLocalWords))] = hi;
      ++nw;
    }
  }
};

// Persistent, warp-specialised form.  A CTA = 8 worker warps (one 8x8 block per thread, 256 blocks
// per tile) + 1 look-back warp, and claims tiles from a per-picture counter (so a tile is only ever
// owned by a running CTA and the look-back chain cannot starve).  Profiling the one-tile-per-CTA
// form showed 48 % of every CTA's life spent by all warps at the barrier behind the look-back,
// waiting for the slowest of the ~32 preceding tiles to publish its bit count.  Here the workers
// do not wait: after walking tile i they hand (tile id, bit total) to the look-back warp and go on
// to walk tile i+1 into the other half of the double-buffered slots; only then do they pick up
// the prefix of tile i -- long since resolved -- and copy its words to the stream.
#ifndef SJB_E_PREFETCH
#define SJB_E_PREFETCH 1
#endif
#ifndef SJB_E_CTAS_PER_SM
#define SJB_E_CTAS_PER_SM ((kTileBlocks >= 512) ? 3 : (kTileBlocks >= 256) ? 5 : 10)
#endif
enum { kEWorkers = kTileBlocks, kEThreads = kTileBlocks + 32, kECtasPerSm = SJB_E_CTAS_PER_SM };
enum { kBarWorkers = 1, kBarFull = 2 /* +buffer */, kBarReady = 4 /* +buffer */ };
enum { kBusyTileBits = 32 * kTileBlocks };   // above 32 bits per block on average a tile counts as busy
// barrier ids are immediates (a register id makes ptxas reserve all 16 hardware barriers per CTA,
// which caps the SM at 4 CTAs)
template <int kId>
__device__ __forceinline__ void bar_sync_id(int count) {
  asm volatile("bar.sync %0, %1;" ::"n"(kId), "r"(count) : "memory");
}
template <int kId>
__device__ __forceinline__ void bar_arrive_id(int count) {
  asm volatile("bar.arrive %0, %1;" ::"n"(kId), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int count) {
  switch (id) {
    case 1: bar_sync_id<1>(count); break;
    case 2: bar_sync_id<2>(count); break;
    case 3: bar_sync_id<3>(count); break;
    case 4: bar_sync_id<4>(count); break;
    default: bar_sync_id<5>(count); break;
  }
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  switch (id) {
    case 2: bar_arrive_id<2>(count); break;
    case 3: bar_arrive_id<3>(count); break;
    case 4: bar_arrive_id<4>(count); break;
    default: bar_arrive_id<5>(count); break;
  }
}
// Exclusive scan over the worker warps with ONE named barrier: every warp leaves its total in the
// half of `scratch` the caller's parity selects, and after the barrier every warp scans those few
// totals itself with shuffles.  Two calls in a row must use different parities (a fast warp may be
// writing the next scan's totals while a slow one still reads this one's); with the parity
// alternating, the barrier of the call in between orders a half's reuse.
__device__ __forceinline__ uint32_t workers_exclusive_scan(uint32_t v, uint32_t* scratch /*[2][8]*/, int parity, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarps = kEWorkers / 32;
  static_assert(nwarps <= 32, "one lane per worker warp");
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  uint32_t* mine = scratch + parity * nwarps;
  if (lane == 31) mine[warp] = incl;
  bar_sync(kBarWorkers, kEWorkers);
  const uint32_t w = (lane < nwarps) ? mine[lane] : 0u;
  uint32_t wi = w;
#pragma unroll
  for (int d = 1; d < nwarps; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
    if (lane >= d) wi += t;
  }
  *total = __shfl_sync(0xffffffffu, wi, nwarps - 1);
  return __shfl_sync(0xffffffffu, wi - w, warp) + incl - v;
}

__global__ void __launch_bounds__(kEThreads, kECtasPerSm)
entropy_pack_kernel(const __grid_constant__ FrameSet fs, GroupBuffers gb) {
  __shared__ __align__(16) CodeTabs sh;
  __shared__ uint32_t scratch[33];
  __shared__ long long tile_id[2], next_id[2];
  __shared__ uint32_t tile_total[2];
  __shared__ unsigned long long tile_prefix[2];
  __shared__ uint32_t local[2][kTileBlocks][kLocalWords + 1];   // odd stride: conflict-free; word 16 = ex | bits << 20
  __shared__ uint32_t sort_hist[64], sort_base[64];             // busy tiles: counting sort of the blocks by non-zeros
  __shared__ uint8_t sort_order[kTileBlocks];
  const int frame = blockIdx.y;
  load_code_tables(gb.tabs + frame, &sh);
  const int16_t* zz = gb.coef + frame * gb.coef_pitch;
  const uint8_t* nzmask = gb.nzmask + frame * gb.mask_pitch;
  const size_t nb_blocks = fs.blocks_per_frame;
  const long long ntiles = static_cast<long long>((nb_blocks + kTileBlocks - 1) / kTileBlocks);
  unsigned long long* state = gb.bit_state + frame * gb.bit_state_pitch;
  unsigned long long* counter = state + (gb.bit_state_pitch - 1);   // last slot: next tile to claim
  const int* dc_init = gb.dc_init ? gb.dc_init + 3 * frame : nullptr;

  if (threadIdx.x >= kEWorkers) {
    // ---- look-back warp ----
    for (int i = 0;; ++i) {
      const int b = i & 1;
      bar_sync(kBarFull + b, kEThreads);            // the workers walked their i-th tile
      const long long t = tile_id[b];
      if (t < 0) break;
      const uint32_t total = tile_total[b];
      const unsigned long long p = warp_lookback<false>(state, t, total);
      if ((threadIdx.x & 31) == 0) {
        tile_prefix[b] = p;
        if (t == ntiles - 1) {
          gb.info[frame].total_bits = p + total;
          gb.info[frame].head_byte = 0;      // stripe hand-over fields, set again by the stuffing kernel
          gb.info[frame].tail_byte = 0;
          gb.info[frame].tail_bits = 0;
          gb.info[frame].head_open = 0;
        }
      }
      __threadfence_block();
      __syncwarp();
      bar_arrive(kBarReady + b, kEThreads);
    }
    return;
  }

  // ---- worker warps ----
  uint32_t* stream = gb.words + frame * gb.words_pitch;
  auto claim_tile = [&]() -> long long {
    const unsigned long long claimed = atomicAdd(counter, 1ull);
    return (claimed < static_cast<unsigned long long>(ntiles)) ? static_cast<long long>(claimed) : -1;
  };
  // copies tile `tile`, walked into buffer pb, to the stream at its now known bit offset
  auto copy_out = [&](long long tile, int pb) {
    bar_sync(kBarReady + pb, kEThreads);
    const size_t g = static_cast<size_t>(tile) * kTileBlocks + threadIdx.x;
    if (g >= nb_blocks) return;
    const uint32_t* mine = local[pb][threadIdx.x];
    const uint32_t packed = mine[kLocalWords];
    const uint32_t bits = packed >> 20;
    const unsigned long long offset = tile_prefix[pb] + (packed & 0xfffffu);
    const int nw = static_cast<int>((bits + 31) >> 5);
    if (nw <= kLocalWords) {
      // shifted copy; the first and the last stream word may be shared with the neighbouring
      // blocks (OR), the others are owned
      const int s = static_cast<int>(offset & 31);
      uint32_t* dst = stream + (offset >> 5);
      const int last = static_cast<int>((s + bits - 1) >> 5);   // index of the last stream word touched
      uint32_t prev = 0;
      for (int kk = 0; kk <= last; ++kk) {
        const uint32_t cur = (kk < nw) ? mine[kk] : 0u;
        const uint32_t v = __funnelshift_r(cur, prev, s);     // (prev << (32-s)) | (cur >> s)
        prev = cur;
        if (kk == 0 || kk == last) { if (v) atomicOr(&dst[kk], v); }
        else dst[kk] = v;
      }
    } else {
      // more than 512 bits: walk the block again, straight into the stream
      const int k = block_in_mcu(g, fs.mcu_blocks);
      const int c = (k >= fs.luma_blocks) ? 1 : 0;
      const int16_t* blk = zz + coef_block_base(g);
      StreamOut out = {stream};
      BitPackSink<StreamOut> sink(out, offset);
      code_block(ChunkLoader{blk}, nzmask[g], blk[0], dc_predictor(zz, g, k, fs.mcu_blocks, fs.luma_blocks, dc_init),
                 sh.dc[c], sh.ac[c], sink);
      sink.finish();
    }
  };
  // bit counts of buffer pb (word kLocalWords of every slot) -> exclusive offsets inside the tile
  auto scan_tile = [&](int pb) -> uint32_t {
    uint32_t* mine = local[pb][threadIdx.x];
    const uint32_t bits = mine[kLocalWords];
    uint32_t total;
    const uint32_t ex = workers_exclusive_scan(bits, scratch, pb, &total);
    mine[kLocalWords] = ex | (bits << 20);        // ex < 256 * 1696 < 2^20, bits < 2^11
    if (threadIdx.x == 0) tile_total[pb] = total;
    return total;
  };
  if (threadIdx.x == 0) next_id[1] = claim_tile();
  bar_sync(kBarWorkers, kEWorkers);
  long long t = next_id[1];
  long long t_prev = -1;
  uint32_t prev_total = 0;
  for (int i = 0;;) {
    const int b = i & 1;
    // The previous tile's bit count decides how this one is walked (neighbouring tiles look alike);
    // it is the same in every worker.
    if (t < 0 || prev_total <= kBusyTileBits) {
      // ---- sparse tile (or the end): thread i walks block i; the copy-out of the previous tile is
      // deferred behind this tile's walk, while the look-back warp resolves its prefix ----
      if (threadIdx.x == 0) tile_id[b] = t;           // for the look-back warp, published by the Full barrier
      long long t_next = -1;
      if (t >= 0) {
        const uint32_t first = static_cast<uint32_t>(t) * kTileBlocks;
        const uint32_t count = min(static_cast<uint32_t>(kTileBlocks), static_cast<uint32_t>(nb_blocks) - first);
        if (threadIdx.x >= count) {
          local[b][threadIdx.x][kLocalWords] = 0;
        } else {
          const size_t g = static_cast<size_t>(first) + threadIdx.x;
          uint32_t* mine = local[b][threadIdx.x];
          const int k = block_in_mcu(g, fs.mcu_blocks);
          const int c = (k >= fs.luma_blocks) ? 1 : 0;
          const int16_t* blk = zz + coef_block_base(g);
          const uint32_t mask = nzmask[g];
          const PrefetchedChunkLoader loader(blk);
          const int pred = dc_predictor(zz, g, k, fs.mcu_blocks, fs.luma_blocks, dc_init);
          const int dc = static_cast<int16_t>(loader.c0.x & 0xffffu);
          LocalSink sink = {mine, 0u, 0u, 0, 0, 0u};
          code_block(loader, mask, dc, pred, sh.dc[c], sh.ac[c], sink);
          sink.finish();
          mine[kLocalWords] = sink.total;
        }
        if (threadIdx.x == 0) next_id[b] = claim_tile();   // broadcast by the barriers of the scan
        prev_total = scan_tile(b);
        t_next = next_id[b];
#if SJB_E_PREFETCH
        if (t_next >= 0) {
          // the next tile's first sector and bitmap byte start their way to L1 now and arrive while
          // this tile's prefix is resolved and the previous tile is written out
          const size_t gn = static_cast<size_t>(t_next) * kTileBlocks + threadIdx.x;
          if (gn < nb_blocks) {
            // (four consecutive blocks share the line that holds their first sectors)
            if ((threadIdx.x & 3) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(zz + coef_block_base(gn)));
            if ((threadIdx.x & 31) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nzmask + gn));
          }
        }
#endif
      }
      __threadfence_block();
      bar_arrive(kBarFull + b, kEThreads);            // t < 0 tells the look-back warp to stop
      if (t_prev >= 0) copy_out(t_prev, b ^ 1);
      if (t < 0) break;
      t_prev = t;
      t = t_next;
      ++i;
      continue;
    }
    // ---- busy tile ----
    // The trip count of a warp is the largest number of non-zeros among its 32 blocks, and blocks
    // differ a lot (luma blocks carry several times the non-zeros of chroma blocks; lane efficiency
    // was 14 of 32 on the 4K gen-A picture).  So every thread first builds the non-zero maps of ITS
    // block (words 0, 1 of the block's slot), the blocks are counting-sorted by their number of
    // non-zeros, and thread r walks the block of rank r, heaviest first.  Slots stay indexed by
    // block; scan and copy-out keep the identity mapping.
    // (Tried: two tiles sorted together, thread r walking rank r and then rank n-1-r so that all
    // warps carry the same sum -- the barrier behind the walk is where a third of the stall samples
    // sit.  The walk phase became balanced and the kernel no faster: the SM's other CTAs already use
    // the issue slots a waiting warp leaves, and the pair cost the double buffering.)
    if (threadIdx.x == 0) tile_id[b] = t;
    if (threadIdx.x < 64) sort_hist[threadIdx.x] = 0;
    // slot j of this buffer is about to be refilled by whichever thread walks block j, which is not
    // the thread that copied tile i-2 out of it: everybody must be done with that copy
    bar_sync(kBarWorkers, kEWorkers);
    const uint32_t first = static_cast<uint32_t>(t) * kTileBlocks;
    const uint32_t count = min(static_cast<uint32_t>(kTileBlocks), static_cast<uint32_t>(nb_blocks) - first);
    uint32_t my_rank = 0, my_key = 0;
    if (threadIdx.x < count) {
      const size_t g = static_cast<size_t>(first) + threadIdx.x;
      const AddrChunkLoader loader(zz + coef_block_base(g));
      uint32_t lo, hi;
      block_nz_maps(loader, nzmask[g], &lo, &hi);
      local[b][threadIdx.x][0] = lo;
      local[b][threadIdx.x][1] = hi;
      my_key = static_cast<uint32_t>(__popc(lo) + __popc(hi));
      my_rank = atomicAdd(&sort_hist[my_key], 1u);
    } else {
      local[b][threadIdx.x][kLocalWords] = 0;
    }
    bar_sync(kBarWorkers, kEWorkers);
    if (threadIdx.x < 64) {
      uint32_t base = 0;
      for (int kk = 63; kk > static_cast<int>(threadIdx.x); --kk) base += sort_hist[kk];
      sort_base[threadIdx.x] = base;
    }
    bar_sync(kBarWorkers, kEWorkers);
    if (threadIdx.x < count) sort_order[sort_base[my_key] + my_rank] = static_cast<uint8_t>(threadIdx.x);
    bar_sync(kBarWorkers, kEWorkers);
    if (threadIdx.x < count) {
      const uint32_t j = sort_order[threadIdx.x];
      const size_t g = static_cast<size_t>(first) + j;
      uint32_t* mine = local[b][j];
      const uint32_t lo = mine[0], hi = mine[1];
      const int k = block_in_mcu(g, fs.mcu_blocks);
      const int c = (k >= fs.luma_blocks) ? 1 : 0;
      const int16_t* blk = zz + coef_block_base(g);
      const AddrChunkLoader loader(blk);
      const int pred = dc_predictor(zz, g, k, fs.mcu_blocks, fs.luma_blocks, dc_init);
      LocalSink sink = {mine, 0u, 0u, 0, 0, 0u};
      code_block_mapped(loader, lo, hi, blk[0], pred, sh.dc[c], sh.ac[c], sink);
      sink.finish();
      mine[kLocalWords] = sink.total;
    }
    if (threadIdx.x == 0) next_id[b] = claim_tile();   // broadcast by the barriers below
    bar_sync(kBarWorkers, kEWorkers);                  // every slot of the tile is filled
    prev_total = scan_tile(b);
    const long long t_next = next_id[b];
    __threadfence_block();
    bar_arrive(kBarFull + b, kEThreads);
    if (t_prev >= 0) copy_out(t_prev, b ^ 1);
    t_prev = t;
    t = t_next;
    ++i;
  }
}

// S1 (entropy.cc:208-227)
struct SmemStats {
  uint32_t* f;   // [272] of the block's table class
  __device__ __forceinline__ void one(int slot) { atomicAdd(&f[slot], 1u); }
  __device__ __forceinline__ void many(int slot, int n) { atomicAdd(&f[slot], static_cast<uint32_t>(n)); }
};

__global__ void __launch_bounds__(kTileBlocks)
symbol_stats_kernel(const __grid_constant__ FrameSet fs, GroupBuffers gb) {
  __shared__ uint32_t f[2][272];
  for (int i = threadIdx.x; i < 2 * 272; i += blockDim.x) (&f[0][0])[i] = 0;
  __syncthreads();
  const int frame = blockIdx.y;
  const int16_t* zz = gb.coef + frame * gb.coef_pitch;
  const uint8_t* nzmask = gb.nzmask + frame * gb.mask_pitch;
  const size_t nb_blocks = fs.blocks_per_frame;
  const size_t stride = static_cast<size_t>(gridDim.x) * kTileBlocks;
  for (size_t g = blockIdx.x * static_cast<size_t>(kTileBlocks) + threadIdx.x; g < nb_blocks; g += stride) {
    const int k = static_cast<int>(g % fs.mcu_blocks);
    const int c = (k >= fs.luma_blocks) ? 1 : 0;
    const int16_t* b = zz + coef_block_base(g);
    SmemStats add = {f[c]};
    block_symbol_stats(ChunkLoader{b}, nzmask[g], b[0],
                       dc_predictor(zz, g, k, fs.mcu_blocks, fs.luma_blocks,
                                    gb.dc_init ? gb.dc_init + 3 * frame : nullptr), add);
  }
  __syncthreads();
  uint32_t* freq = gb.freq + static_cast<size_t>(frame) * 2 * 272;
  for (int i = threadIdx.x; i < 2 * 272; i += blockDim.x) {
    const uint32_t c = (&f[0][0])[i];
    if (c) atomicAdd(&freq[i], c);
  }
}

// -------------------------------------------------------------------------------------------
// S: byte stuffing (bit_writer.h:172-196, bit_writer.cc:107-116, headers.cc:262-268).
// The word stream holds ceil(total_bits/8) bytes, MSB-first inside each word; the last byte is
// padded with 1-bits.  Every 0xFF byte is followed by 0x00.  Tile = 4096 stream bytes per CTA
// iteration, 16 bytes per thread; persistent CTAs claim the tiles of a picture from a counter.
// -------------------------------------------------------------------------------------------
// 16 bytes of the byte-aligned stream R = (shift zero bits) ++ packed bits, starting at byte0
// (multiple of 16).  end_bits = shift + total_bits; bytes at or past ceil(end_bits/8) read as 0.
__device__ __forceinline__ uint4 load_stream16(const uint32_t* stream, unsigned long long byte0,
                                               unsigned long long end_bits, unsigned shift) {
  uint4 w = make_uint4(0, 0, 0, 0);
  if (byte0 * 8 < end_bits) {
    const unsigned long long j = byte0 >> 2;
    w = *reinterpret_cast<const uint4*>(stream + j);
    if (shift) {
      const uint32_t before = (j > 0) ? stream[j - 1] : 0u;
      const uint32_t x = w.x, y = w.y, z = w.z;
      w.x = __funnelshift_r(x, before, shift);
      w.y = __funnelshift_r(y, x, shift);
      w.z = __funnelshift_r(z, y, shift);
      w.w = __funnelshift_r(w.w, z, shift);
    }
  }
  return w;
}
__device__ __forceinline__ int count_ff(uint32_t w) {
  // a byte is 0xFF iff all its bits are set
  uint32_t t = w & (w >> 4);
  t &= t >> 2;
  t &= t >> 1;
  return __popc(t & 0x01010101u);
}
__device__ __forceinline__ uint32_t stream_byte(const uint4& w, int i) {   // byte i (0..15) in stream order
  const uint32_t v = (i < 4) ? w.x : (i < 8) ? w.y : (i < 12) ? w.z : w.w;
  return (v >> (8 * (3 - (i & 3)))) & 0xffu;
}

// what a thread keeps of its 16 stream bytes of a tile between the two halves of the tile's life
struct StuffPiece {
  uint4 w;             // the 16 bytes (padding applied)
  uint32_t ex;         // 0xFF bytes before this thread inside the tile
  uint32_t total;      // 0xFF bytes of the tile
  int lo, hi;          // bytes [lo, hi) of w are emitted
};

__global__ void __launch_bounds__(kStuffThreads)
stuff_kernel(GroupBuffers gb, const __grid_constant__ StuffArgs args) {
  __shared__ uint32_t scratch[33];
  __shared__ unsigned long long tile_prefix;
  __shared__ unsigned long long claimed;
  __shared__ __align__(16) uint8_t obuf[2 * kStuffTileBytes + 32];   // a tile's output: at most every byte stuffed, + EOI
  const int frame = blockIdx.y;
  uint32_t* stream = gb.words + frame * gb.words_pitch;
  unsigned long long* state = gb.ff_state + frame * gb.ff_state_pitch;
  // Tiles are claimed from a per-picture counter (last slot of the picture's descriptor row), as
  // in the entropy kernel: a tile is only ever owned by a running CTA, so the look-back chain makes
  // progress whatever the grid size and whatever else occupies the SMs.
  unsigned long long* counter = state + (gb.ff_state_pitch - 1);
  const unsigned shift = gb.bit_offsets ? static_cast<unsigned>(gb.bit_offsets[frame] & 7) : args.shift[frame];
  const unsigned flags = args.flags[frame];
  const bool last = (flags & kStuffLast) != 0;
  const unsigned long long end_bits = gb.info[frame].total_bits + shift;
  // bytes [b0, b1) of R are emitted here; with kStuffLast the final partial byte is padded with
  // 1-bits (bit_writer.cc:107-116) and emitted too
  const unsigned long long b0 = (shift != 0 && !(flags & kStuffFirst)) ? 1 : 0;
  const unsigned long long b1 = last ? (end_bits + 7) >> 3 : end_bits >> 3;
  const unsigned pad = last ? static_cast<unsigned>((0 - end_bits) & 7) : 0;
  const unsigned long long tiles = (((end_bits + 7) >> 3) + kStuffTileBytes - 1) / kStuffTileBytes;
  uint8_t* out = gb.out + frame * gb.out_pitch + args.header_len[frame];

  // first half of a tile: load, count, scan, PUBLISH the tile's 0xFF count.  Run for the next
  // tile BEFORE the current one is written out: the lowest unclaimed tile is the one every later
  // tile's look-back is about to wait for, so its count must be out within a microsecond of the
  // claim -- not after the byte-wise scatter of a busy tile (claiming ahead and publishing late
  // cost 15 % of the 4K gen-A pipeline).
  auto first_half = [&](unsigned long long t, StuffPiece& p) {
    const unsigned long long byte0 = t * kStuffTileBytes + threadIdx.x * 16ull;
    p.w = load_stream16(stream, byte0, end_bits, shift);
    if (pad && byte0 < b1 && b1 - byte0 <= 16) {             // the padded byte is in here
      const unsigned i = static_cast<unsigned>(b1 - 1 - byte0);
      const uint32_t m = ((1u << pad) - 1u) << (8 * (3 - (i & 3)));
      if ((i >> 2) == 0) p.w.x |= m; else if ((i >> 2) == 1) p.w.y |= m; else if ((i >> 2) == 2) p.w.z |= m; else p.w.w |= m;
    }
    p.lo = (byte0 < b0) ? static_cast<int>(min(16ull, b0 - byte0)) : 0;
    p.hi = (byte0 >= b1) ? 0 : static_cast<int>(min(16ull, b1 - byte0));
    uint32_t ff = 0;
    if (p.lo == 0 && p.hi == 16) {
      ff = static_cast<uint32_t>(count_ff(p.w.x) + count_ff(p.w.y) + count_ff(p.w.z) + count_ff(p.w.w));
    } else {
      for (int i = p.lo; i < p.hi; ++i) ff += (stream_byte(p.w, i) == 0xffu) ? 1u : 0u;
    }
    p.ex = cta_exclusive_scan(ff, scratch, &p.total);
    if (threadIdx.x == 0) lookback_publish(state, static_cast<long long>(t), p.total);
    // shared bytes of a stripe: reported, not emitted
    if (b0 == 1 && byte0 == 0) {
      gb.info[frame].head_byte = static_cast<unsigned char>(stream_byte(p.w, 0));
      gb.info[frame].head_open = (!last && end_bits < 8) ? 1 : 0;
    }
    if (!last && byte0 <= b1 && b1 < byte0 + 16) {
      gb.info[frame].tail_bits = static_cast<unsigned char>(end_bits & 7);
      gb.info[frame].tail_byte = static_cast<unsigned char>((end_bits & 7) ? stream_byte(p.w, static_cast<int>(b1 - byte0)) : 0u);
    }
    if (byte0 * 8 < end_bits && !(flags & kStuffKeepWords)) {
      // self-cleaning: the stream buffer must be all zero for the next encode's atomicOr
      // (only without a shift: shifted reads look one word back into the neighbour's words)
      *reinterpret_cast<uint4*>(stream + (byte0 >> 2)) = make_uint4(0, 0, 0, 0);
    }
  };

  if (threadIdx.x == 0) claimed = atomicAdd(counter, 1ull);
  __syncthreads();
  unsigned long long t = claimed;
  if (t >= tiles) return;
  StuffPiece cur;
  first_half(t, cur);
  for (;;) {
    // resolve this tile's prefix (its count has long been published) and claim the next tile
    __syncthreads();                       // everybody has read `claimed`; tile_prefix is free
    if (threadIdx.x < 32) {
      const unsigned long long p = warp_lookback<true>(state, static_cast<long long>(t), cur.total);
      if (threadIdx.x == 0) {
        tile_prefix = p;
        claimed = atomicAdd(counter, 1ull);
        if (t == tiles - 1) {
          gb.info[frame].stuffed_bytes = p + cur.total;
          gb.info[frame].out_size = args.header_len[frame] + (b1 > b0 ? b1 - b0 : 0) + p + cur.total + (last ? 2 : 0);
        }
      }
    }
    __syncthreads();
    const unsigned long long prefix = tile_prefix;
    const unsigned long long t_next = claimed;
    StuffPiece nxt;
    if (t_next < tiles) first_half(t_next, nxt);     // uniform branch: the scan's barriers are safe
    // second half: the tile's output bytes (0x00 inserted, EOI appended) are put together in shared
    // memory and leave as aligned 16-byte stores.  (Written straight to global memory one byte at a
    // time -- 17 to 33 scattered single-byte stores per thread -- the scatter was the whole cost of
    // this kernel on busy pictures: 121 us per 16 gen-A 4K pictures, 25 % issue-slot use, warps
    // waiting on the store path.)
    const unsigned long long tile_lo = max(t * static_cast<unsigned long long>(kStuffTileBytes), b0);
    const unsigned long long tile_hi = min((t + 1) * static_cast<unsigned long long>(kStuffTileBytes), b1);
    uint32_t n_out = (tile_hi > tile_lo) ? static_cast<uint32_t>(tile_hi - tile_lo) + cur.total : 0u;
    if (cur.hi > cur.lo) {
      const unsigned long long byte0 = t * kStuffTileBytes + threadIdx.x * 16ull;
      uint8_t* o = obuf + static_cast<uint32_t>(byte0 + cur.lo - tile_lo) + cur.ex;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i >= cur.lo && i < cur.hi) {
          const uint32_t b = stream_byte(cur.w, i);
          *o++ = static_cast<uint8_t>(b);
          if (b == 0xffu) *o++ = 0;
        }
      }
      if (last && byte0 + cur.hi == b1) {   // EOI (headers.cc:262-268)
        o[0] = 0xff;
        o[1] = 0xd9;
      }
    }
    if (last && tile_hi == b1 && tile_hi > tile_lo) n_out += 2;
    __syncthreads();
    if (n_out != 0) {
      uint8_t* dst = out + (tile_lo - b0) + prefix;
      const uint32_t head = min(n_out, static_cast<uint32_t>((16u - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
      if (threadIdx.x < head) dst[threadIdx.x] = obuf[threadIdx.x];
      const uint32_t body = (n_out - head) >> 4;          // aligned 16-byte pieces
      const uint32_t* ow = reinterpret_cast<const uint32_t*>(obuf);
      const uint32_t wsh = head >> 2, bsh = head & 3u;    // obuf is word aligned: uniform funnel shift
      for (uint32_t k = threadIdx.x; k < body; k += kStuffThreads) {
        const uint32_t* q = ow + wsh + 4 * k;
        uint4 v;
        if (bsh == 0) {
          v = make_uint4(q[0], q[1], q[2], q[3]);
        } else {
          const uint32_t sel = 0x3210u + 0x1111u * bsh;   // bytes bsh..bsh+3 of the pair {lo, hi}
          v = make_uint4(__byte_perm(q[0], q[1], sel), __byte_perm(q[1], q[2], sel), __byte_perm(q[2], q[3], sel),
                         __byte_perm(q[3], q[4], sel));
        }
        *reinterpret_cast<uint4*>(dst + head + 16 * k) = v;
      }
      const uint32_t done = head + 16 * body;
      if (threadIdx.x < n_out - done) dst[done + threadIdx.x] = obuf[done + threadIdx.x];
    }
    if (last && b1 <= b0 && t == 0 && threadIdx.x == 0) {   // a last stripe without a byte of its own: EOI only
      out[0] = 0xff;
      out[1] = 0xd9;
    }
    if (t_next >= tiles) break;
    t = t_next;
    cur = nxt;
  }
}

// -------------------------------------------------------------------------------------------
// T1: trellis quantisation (quantize.cc:325-457), one block per thread; the dynamic programme is
// block_ops.cuh::trellis_block_v2 (shared with the CPU emulation).  A thread's working storage --
// the block, the prefix sums of V^2, two nodes per position -- is a private 1 KB column of SHARED
// memory laid out [word][lane] (bank = lane: conflict-free whatever each lane indexes); nothing is
// in local memory.  That is 32 KB per warp, so a CTA is 7 warps with 224 KB of dynamic shared
// memory, one CTA per SM.  Thread t of the grid takes the block of rank t in the order left by the
// counting sort below (heaviest first).
// -------------------------------------------------------------------------------------------
enum { kTrWarps = 7, kTrThreads = 32 * kTrWarps };
struct TrellisSmemMem {
  uint32_t* coef;    // [32][32 lanes], this lane's column
  uint32_t* disto;   // [64][32]
  uint2* score;      // [64][32]  {candidate 0, candidate 1}
  uint32_t* prev;    // [32][32] words = 128 bytes per lane
  __device__ __forceinline__ uint32_t coef_ld(int w) const { return coef[w * 32]; }
  __device__ __forceinline__ void coef_st(int w, uint32_t v) { coef[w * 32] = v; }
  __device__ __forceinline__ uint32_t disto_ld(int i) const { return disto[i * 32]; }
  __device__ __forceinline__ void disto_st(int i, uint32_t v) { disto[i * 32] = v; }
  __device__ __forceinline__ void score_ld(int pos, uint32_t& s0, uint32_t& s1) const {
    const uint2 v = score[pos * 32];
    s0 = v.x;
    s1 = v.y;
  }
  __device__ __forceinline__ void score_st(int pos, int k, uint32_t v) {
    reinterpret_cast<uint32_t*>(&score[pos * 32])[k] = v;
  }
  __device__ __forceinline__ uint32_t prev_ld(int node) const {
    return reinterpret_cast<const uint8_t*>(&prev[(node >> 2) * 32])[node & 3];
  }
  __device__ __forceinline__ void prev_st(int node, uint32_t v) {
    reinterpret_cast<uint8_t*>(&prev[(node >> 2) * 32])[node & 3] = static_cast<uint8_t>(v);
  }
  // the quantised block goes where the prefix sums were (they are dead by then)
  __device__ __forceinline__ void out_zero() {
#pragma unroll
    for (int w = 0; w < 32; ++w) disto[w * 32] = 0;
  }
  __device__ __forceinline__ void out_st(int pos, int v) {
    reinterpret_cast<uint16_t*>(&disto[(pos >> 1) * 32])[pos & 1] = static_cast<uint16_t>(v);
  }
  __device__ __forceinline__ uint32_t out_ld(int w) const { return disto[w * 32]; }
};
struct TrellisSmemTab {
  const int32_t* qtab;     // [64][2] of this block's matrix, zig-zag order
  const uint16_t* q16s;    // [64] matrix entry << 4, zig-zag order
  const uint8_t* ac_len;   // [256]
  __device__ __forceinline__ void qt(int i, int& iq, int& cpos) const {
    const int2 v = *reinterpret_cast<const int2*>(qtab + 2 * i);
    iq = v.x;
    cpos = v.y;
  }
  __device__ __forceinline__ int q16(int i) const { return q16s[i]; }
  __device__ __forceinline__ uint32_t len(int sym) const { return ac_len[sym]; }
};
enum { kTrTableBytes = 2 * 64 * 2 * 4 + 2 * 64 * 2 + 2 * 256, kTrSmemBytes = kTrThreads * 1024 + kTrTableBytes };

// Blocks are handed to the trellis threads SORTED by their number of non-zero AC positions,
// heaviest first: the trip count of a warp is the largest count among its 32 blocks, and a busy
// picture mixes luma blocks with 40 non-zeros and chroma blocks with 5 (measured on the 8K gen-A
// picture: with blocks in scan order a CTA ran for as long as its heaviest block, 3.8 x the mean).
// Counting sort in two small kernels: keys (0..63) into the bitmap array (free until T1 writes it)
// plus a per-picture histogram; then every CTA ranks its blocks inside the buckets.
// trellis_sort[frame]: [0,64) bucket counts, [64,128) bucket fill cursors; zero on entry.
__global__ void __launch_bounds__(256)
trellis_keys_kernel(const __grid_constant__ FrameSet fs, GroupBuffers gb, const int16_t* __restrict__ raw_src,
                    uint32_t* __restrict__ sort_state) {
  __shared__ __align__(16) int32_t qtab[2][64][2];
  __shared__ uint32_t hist[64];
  const int frame = blockIdx.y;
  {
    const int32_t* q = &gb.qtabs[frame].m[0].e[0][0];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) (&qtab[0][0][0])[i] = q[i];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
  }
  __syncthreads();
  const size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (g < fs.blocks_per_frame) {
    constexpr int zz[64] = SJB_ZIGZAG_INIT;
    const int c = (block_in_mcu(g, fs.mcu_blocks) >= fs.luma_blocks) ? 1 : 0;
    int v[64];
    load_block_natural((raw_src ? raw_src : gb.coef) + frame * gb.coef_pitch + coef_block_base(g), v);
    int n = 0;
#pragma unroll
    for (int i = 1; i < 64; ++i) {
      const int a = abs(v[zz[i]]);
      n += (a * qtab[c][i][0] + qtab[c][i][1] >= (1 << 20)) ? 1 : 0;
    }
    gb.nzmask[frame * gb.mask_pitch + g] = static_cast<uint8_t>(n);
    atomicAdd(&hist[n], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 64 && hist[threadIdx.x]) atomicAdd(&sort_state[frame * 128 + threadIdx.x], hist[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
trellis_rank_kernel(const __grid_constant__ FrameSet fs, GroupBuffers gb, uint32_t* __restrict__ sort_state,
                    uint32_t* __restrict__ perm, size_t perm_pitch) {
  __shared__ uint32_t base[64], cnt[64], cta_base[64];
  const int frame = blockIdx.y;
  uint32_t* state = sort_state + frame * 128;
  if (threadIdx.x < 64) {
    // heaviest bucket first: base[k] = number of blocks with a larger key
    uint32_t b = 0;
    for (int k = 63; k > static_cast<int>(threadIdx.x); --k) b += state[k];
    base[threadIdx.x] = b;
    cnt[threadIdx.x] = 0;
  }
  __syncthreads();
  const size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  int key = 0;
  uint32_t local = 0;
  if (g < fs.blocks_per_frame) {
    key = gb.nzmask[frame * gb.mask_pitch + g];
    local = atomicAdd(&cnt[key], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 64 && cnt[threadIdx.x]) cta_base[threadIdx.x] = atomicAdd(&state[64 + threadIdx.x], cnt[threadIdx.x]);
  __syncthreads();
  if (g < fs.blocks_per_frame) perm[frame * perm_pitch + base[key] + cta_base[key] + local] = static_cast<uint32_t>(g);
}

__global__ void __launch_bounds__(kTrThreads, 1)
trellis_kernel(const __grid_constant__ FrameSet fs, GroupBuffers gb, const int16_t* __restrict__ raw_src,
               const uint32_t* __restrict__ perm, size_t perm_pitch) {
  extern __shared__ __align__(16) uint8_t trellis_smem[];
  int32_t* qtab = reinterpret_cast<int32_t*>(trellis_smem);                   // [2][64][2]
  uint16_t* q16s = reinterpret_cast<uint16_t*>(trellis_smem + 1024);         // [2][64]
  uint8_t* ac_len = trellis_smem + 1024 + 256;                               // [2][256]
  uint8_t* columns = trellis_smem + kTrTableBytes;                           // 32 KB per warp
  const int frame = blockIdx.y;
  {
    constexpr int zz[64] = SJB_ZIGZAG_INIT;
    const CodeTabs* tabs = gb.tabs + frame;
    const uint8_t* quant = gb.quant + frame * 128;
    const int32_t* q = &gb.qtabs[frame].m[0].e[0][0];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) ac_len[i] = static_cast<uint8_t>(tabs->ac[i >> 8][i & 255] & 0xff);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) q16s[i] = static_cast<uint16_t>(quant[(i & 64) + zz[i & 63]]) << 4;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) qtab[i] = q[i];
  }
  __syncthreads();
  const uint32_t rank = blockIdx.x * static_cast<uint32_t>(kTrThreads) + threadIdx.x;
  if (rank >= fs.blocks_per_frame) return;
  const size_t g = perm[frame * perm_pitch + rank];
  const int c = (block_in_mcu(g, fs.mcu_blocks) >= fs.luma_blocks) ? 1 : 0;
  int16_t* blk = gb.coef + frame * gb.coef_pitch + coef_block_base(g);
  uint32_t raw[32];
  {
    const uint4* s = reinterpret_cast<const uint4*>(raw_src ? raw_src + frame * gb.coef_pitch + coef_block_base(g) : blk);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 w = s[coef_chunk_index(i)];
      raw[4 * i] = w.x; raw[4 * i + 1] = w.y; raw[4 * i + 2] = w.z; raw[4 * i + 3] = w.w;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* col = columns + warp * 32768;
  TrellisSmemMem mem;
  mem.coef = reinterpret_cast<uint32_t*>(col) + lane;                        // 4 KB
  mem.disto = reinterpret_cast<uint32_t*>(col + 4096) + lane;                // 8 KB
  mem.score = reinterpret_cast<uint2*>(col + 12288) + lane;                  // 16 KB
  mem.prev = reinterpret_cast<uint32_t*>(col + 28672) + lane;                // 4 KB
  const TrellisSmemTab tab = {qtab + c * 128, q16s + c * 64, ac_len + c * 256};
  const uint32_t mask = trellis_block_v2(raw, mem, tab);
  {
    uint4* d = reinterpret_cast<uint4*>(blk);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      d[coef_chunk_index(i)] = make_uint4(mem.out_ld(4 * i), mem.out_ld(4 * i + 1), mem.out_ld(4 * i + 2), mem.out_ld(4 * i + 3));
    }
  }
  gb.nzmask[frame * gb.mask_pitch + g] = static_cast<uint8_t>(mask);
}

// quantised DC of the last block of each component (what the next stripe predicts from)
__global__ void last_dc_kernel(const __grid_constant__ FrameSet fs, GroupBuffers gb, int* out /*[frames][3]*/) {
  const int t = threadIdx.x;
  if (t >= fs.frames * 3) return;
  const int frame = t / 3, comp = t % 3;
  const int nb_comps = (fs.mcu_blocks == 1) ? 1 : 3;
  int v = 0;
  if (comp < nb_comps) {
    const size_t last_mcu = static_cast<size_t>(fs.blocks_per_frame) - fs.mcu_blocks;
    const size_t g = last_mcu + ((comp == 0) ? fs.luma_blocks - 1 : fs.luma_blocks + comp - 1);
    v = gb.coef[frame * gb.coef_pitch + coef_block_base(g)];
  }
  out[t] = v;
}

// ---- row stripes: glue between the collectives ------------------------------------------------
__global__ void stripe_bits_kernel(GroupBuffers gb, int frames, unsigned long long* bits) {
  const int f = threadIdx.x;
  if (f < frames) bits[f] = gb.info[f].total_bits;
}
__global__ void stripe_offsets_kernel(const unsigned long long* __restrict__ all_bits, int n, int rank,
                                      unsigned long long* __restrict__ offsets) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long o = 0;
  for (int r = 0; r < rank; ++r) o += all_bits[static_cast<size_t>(r) * n + i];
  offsets[i] = o;
}
__global__ void stripe_meta_kernel(GroupBuffers gb, int frames, unsigned long long* meta) {
  const int f = threadIdx.x;
  if (f >= frames) return;
  const StreamInfo& in = gb.info[f];
  meta[2 * f] = in.out_size;
  meta[2 * f + 1] = static_cast<unsigned long long>(in.head_byte) | (static_cast<unsigned long long>(in.tail_byte) << 8) |
                    (static_cast<unsigned long long>(in.tail_bits) << 16) | (static_cast<unsigned long long>(in.head_open) << 24);
}
// Small copies between PINNED HOST and device memory done by a kernel (pinned memory is device-
// accessible under unified addressing) instead of the copy engine.  The copy engine serves a
// direction in FIFO order across streams: a 2 KB table upload enqueued after 50 MB of pixel uploads
// of a LATER group waits for all of them (measured in the stripe exchange: the entropy stage of a
// chunk started 0.86 ms late, exactly the next chunk's upload time).  A load/store from an SM does
// not queue behind the engine.  Either pointer may be pinned host or device memory.
__global__ void copy_small_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t bytes) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 7) == 0) {
    const size_t words = bytes >> 3;
    for (size_t i = t; i < words; i += stride) {
      reinterpret_cast<unsigned long long*>(dst)[i] = reinterpret_cast<const unsigned long long*>(src)[i];
    }
    for (size_t i = (words << 3) + t; i < bytes; i += stride) dst[i] = src[i];
  } else {
    for (size_t i = t; i < bytes; i += stride) dst[i] = src[i];
  }
  __threadfence_system();
}
// gridDim.y = stripe of the group; CTAs stride over its bytes
__global__ void __launch_bounds__(256)
stripe_compact_kernel(const uint8_t* __restrict__ group_out, size_t out_pitch, int first,
                      const unsigned long long* __restrict__ meta, uint8_t* __restrict__ dst) {
  const int f = blockIdx.y;
  __shared__ unsigned long long base;
  if (threadIdx.x == 0) {
    unsigned long long b = 0;
    for (int i = 0; i < first + f; ++i) b += meta[2 * i];
    base = b;
  }
  __syncthreads();
  const unsigned long long size = meta[2 * (first + f)];
  const uint8_t* src = group_out + f * out_pitch;
  uint8_t* d = dst + base;
  for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < size; i += gridDim.x * 256ull) d[i] = src[i];
}

unsigned cdiv(size_t a, size_t b) { return static_cast<unsigned>((a + b - 1) / b); }

// SMs of the current device (grids of the persistent kernels are sized from it), cached per device;
// concurrent host threads may race to fill a slot with the same value.
unsigned SmCount() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return static_cast<unsigned>(n);
}

}  // namespace

// -------------------------------------------------------------------------------------------
// launch wrappers
// -------------------------------------------------------------------------------------------
void LaunchF1Generic(const FrameSet& fs, int mx0, int my0, int mx1, int my1, bool raw,
                     const QuantTabs& qt, const GroupBuffers& gb, cudaStream_t s) {
  if (mx1 <= mx0 || my1 <= my0) return;
  const size_t nb = static_cast<size_t>(mx1 - mx0) * (my1 - my0) * fs.mcu_blocks;
  const dim3 grid(cdiv(nb, 128), fs.frames);
  if (raw) f1_generic_kernel<true><<<grid, 128, 0, s>>>(fs, mx0, my0, mx1, my1, qt, gb);
  else     f1_generic_kernel<false><<<grid, 128, 0, s>>>(fs, mx0, my0, mx1, my1, qt, gb);
}

bool F1FastEligible(const FrameSet& fs) {
  if ((fs.stride & 15) != 0) return false;
  for (int f = 0; f < fs.frames; ++f) {
    if ((reinterpret_cast<uintptr_t>(fs.pix[f]) & 15) != 0) return false;
  }
  if (!fs.planar) return true;
  // planar sources: every plane 16-byte aligned; semi-planar: U and V one byte apart in one plane
  if (fs.yuv_mode == kYuv400) return true;
  if ((fs.stride_u & 15) != 0 || (fs.stride_v & 15) != 0) return false;
  for (int f = 0; f < fs.frames; ++f) {
    const uintptr_t u = reinterpret_cast<uintptr_t>(fs.pix_u[f]), v = reinterpret_cast<uintptr_t>(fs.pix_v[f]);
    if (fs.uv_step == 2) {
      if (fs.stride_u != fs.stride_v || (u > v ? u - v : v - u) != 1 || ((u < v ? u : v) & 15) != 0) return false;
    } else if (fs.uv_step != 1 || (u & 15) != 0 || (v & 15) != 0) {
      return false;
    }
  }
  return true;
}

template <int kMode, bool kRaw, int kFmt>
static void LaunchF1FastT(const FrameSet& fs, int mx_full, int my0, int my1, const QuantTabs& qt,
                          const GroupBuffers& gb, cudaStream_t s) {
  const int per_tile = (kMode == kYuv420) ? 16 : 32;
  const long long tiles = static_cast<long long>((mx_full + per_tile - 1) / per_tile) * (my1 - my0);
  const dim3 grid(static_cast<unsigned>(tiles), fs.frames);
  f1_fast_kernel<kMode, kRaw, kFmt><<<grid, 32, 0, s>>>(fs, mx_full, my0, qt, gb);
}

template <int kMode, bool kRaw>
static void LaunchF1FastF(const FrameSet& fs, int mx_full, int my0, int my1, const QuantTabs& qt,
                          const GroupBuffers& gb, cudaStream_t s) {
  if (fs.pix_fmt == kFmtRGB) LaunchF1FastT<kMode, kRaw, kFmtRGB>(fs, mx_full, my0, my1, qt, gb, s);
  else if (fs.pix_fmt == kFmtRGBA) LaunchF1FastT<kMode, kRaw, kFmtRGBA>(fs, mx_full, my0, my1, qt, gb, s);
  else LaunchF1FastT<kMode, kRaw, kFmtBGRA>(fs, mx_full, my0, my1, qt, gb, s);
}

template <int kMode, bool kRaw>
static void LaunchF1PlanarT(const FrameSet& fs, int mx_full, int my0, int my1, const QuantTabs& qt,
                            const GroupBuffers& gb, cudaStream_t s) {
  const int per_tile = (kMode == kYuv420) ? 16 : 32;
  const long long tiles = static_cast<long long>((mx_full + per_tile - 1) / per_tile) * (my1 - my0);
  f1_planar_kernel<kMode, kRaw><<<dim3(static_cast<unsigned>(tiles), fs.frames), 32, 0, s>>>(fs, mx_full, my0, qt, gb);
}

void LaunchF1Fast(const FrameSet& fs, int mx_full, int my0, int my1, bool raw, const QuantTabs& qt,
                  const GroupBuffers& gb, cudaStream_t s) {
  if (mx_full <= 0 || my1 <= my0) return;
  if (fs.planar) {
    switch (fs.yuv_mode) {
      case kYuv420:
        if (raw) LaunchF1PlanarT<kYuv420, true>(fs, mx_full, my0, my1, qt, gb, s);
        else     LaunchF1PlanarT<kYuv420, false>(fs, mx_full, my0, my1, qt, gb, s);
        break;
      case kYuv444:
        if (raw) LaunchF1PlanarT<kYuv444, true>(fs, mx_full, my0, my1, qt, gb, s);
        else     LaunchF1PlanarT<kYuv444, false>(fs, mx_full, my0, my1, qt, gb, s);
        break;
      default:
        if (raw) LaunchF1PlanarT<kYuv400, true>(fs, mx_full, my0, my1, qt, gb, s);
        else     LaunchF1PlanarT<kYuv400, false>(fs, mx_full, my0, my1, qt, gb, s);
        break;
    }
    return;
  }
  switch (fs.yuv_mode) {
    case kYuv420:
      if (raw) LaunchF1FastF<kYuv420, true>(fs, mx_full, my0, my1, qt, gb, s);
      else     LaunchF1FastF<kYuv420, false>(fs, mx_full, my0, my1, qt, gb, s);
      break;
    case kYuv444:
      if (raw) LaunchF1FastF<kYuv444, true>(fs, mx_full, my0, my1, qt, gb, s);
      else     LaunchF1FastF<kYuv444, false>(fs, mx_full, my0, my1, qt, gb, s);
      break;
    default:
      if (raw) LaunchF1FastF<kYuv400, true>(fs, mx_full, my0, my1, qt, gb, s);
      else     LaunchF1FastF<kYuv400, false>(fs, mx_full, my0, my1, qt, gb, s);
      break;
  }
}

void LaunchRequantize(const FrameSet& fs, const GroupBuffers& gb, const int16_t* raw_src, cudaStream_t s) {
  requantize_kernel<<<dim3(cdiv(fs.blocks_per_frame, 128), fs.frames), 128, 0, s>>>(fs, gb, raw_src);
}

void LaunchQuantError(const FrameSet& fs, const GroupBuffers& gb, const int16_t* raw, unsigned long long* err,
                      cudaStream_t s) {
  unsigned grid = cdiv(fs.blocks_per_frame, 256);
  if (grid > SmCount() * 8) grid = SmCount() * 8;
  quant_error_kernel<<<dim3(grid, fs.frames), 256, 0, s>>>(fs, gb, raw, err);
}

void LaunchHistogram(const FrameSet& fs, const GroupBuffers& gb, cudaStream_t s) {
  // the attribute is per device; concurrent host threads may get here together (any of them may set it)
  static std::atomic<bool> init[64];
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!init[dev].load(std::memory_order_acquire)) {
    cudaFuncSetAttribute(histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kH1SmemBytes));
    init[dev].store(true, std::memory_order_release);
  }
  // 3 CTAs (66 KB of counters each) per SM over all pictures of the group, at least 64 blocks per warp
  unsigned grid = cdiv(fs.blocks_per_frame, (kH1Threads / 32) * 64);
  unsigned cap = SmCount() * 3 / (fs.frames > 0 ? fs.frames : 1);   // rounded DOWN: a 445th CTA would run alone in a second wave
  if (cap < 1) cap = 1;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  histogram_kernel<<<dim3(grid, fs.frames), kH1Threads, kH1SmemBytes, s>>>(fs, gb);
}

void LaunchAnalyseHistograms(int frames, const GroupBuffers& gb, const AqParams& ap, AqFit* fit, int* fail, cudaStream_t s) {
  const int comps = ap.nb_comps > 1 ? 2 : 1;
  analyse_fit_kernel<<<dim3(64, comps, frames), 128, 0, s>>>(gb, ap, fit, fail);
  analyse_pick_kernel<<<dim3(comps, frames), 64, 0, s>>>(gb, ap, fit, fail);
}

void LaunchTrellis(const FrameSet& fs, const GroupBuffers& gb, const int16_t* raw_src, uint32_t* sort_state, uint32_t* perm,
                   size_t perm_pitch, cudaStream_t s) {
  static std::atomic<bool> init[64];
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!init[dev].load(std::memory_order_acquire)) {
    cudaFuncSetAttribute(trellis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kTrSmemBytes));
    init[dev].store(true, std::memory_order_release);
  }
  // counting sort of the blocks by their number of non-zero AC positions, then the dynamic programme
  cudaMemsetAsync(sort_state, 0, static_cast<size_t>(fs.frames) * 128 * sizeof(uint32_t), s);
  const dim3 grid256(cdiv(fs.blocks_per_frame, 256), fs.frames);
  trellis_keys_kernel<<<grid256, 256, 0, s>>>(fs, gb, raw_src, sort_state);
  trellis_rank_kernel<<<grid256, 256, 0, s>>>(fs, gb, sort_state, perm, perm_pitch);
  trellis_kernel<<<dim3(cdiv(fs.blocks_per_frame, kTrThreads), fs.frames), kTrThreads, kTrSmemBytes, s>>>(fs, gb, raw_src, perm,
                                                                                                       perm_pitch);
}

void LaunchSymbolStats(const FrameSet& fs, const GroupBuffers& gb, cudaStream_t s) {
  unsigned grid = cdiv(fs.blocks_per_frame, kTileBlocks);
  unsigned cap = SmCount() * 8 / (fs.frames > 0 ? fs.frames : 1);   // rounded down: one wave
  if (cap < 1) cap = 1;
  if (grid > cap) grid = cap;
  symbol_stats_kernel<<<dim3(grid, fs.frames), kTileBlocks, 0, s>>>(fs, gb);
}

void LaunchEntropyPack(const FrameSet& fs, const GroupBuffers& gb, cudaStream_t s) {
  // persistent CTAs claiming tiles from a counter: about 5 CTAs per SM over all pictures of the group
  const unsigned tiles = cdiv(fs.blocks_per_frame, kTileBlocks);
  unsigned grid = SmCount() * kECtasPerSm / (fs.frames > 0 ? fs.frames : 1);
  if (grid < 1) grid = 1;
  if (grid > tiles) grid = tiles;
  entropy_pack_kernel<<<dim3(grid, fs.frames), kEThreads, 0, s>>>(fs, gb);
}

void LaunchLastDc(const FrameSet& fs, const GroupBuffers& gb, int* out, cudaStream_t s) {
  last_dc_kernel<<<1, (3 * kMaxGroup + 31) / 32 * 32, 0, s>>>(fs, gb, out);
}

void LaunchStripeBits(const GroupBuffers& gb, int frames, unsigned long long* bits, cudaStream_t s) {
  stripe_bits_kernel<<<1, 32, 0, s>>>(gb, frames, bits);
}
void LaunchStripeOffsets(const unsigned long long* all_bits, int n, int rank, unsigned long long* offsets, cudaStream_t s) {
  stripe_offsets_kernel<<<cdiv(n, 128), 128, 0, s>>>(all_bits, n, rank, offsets);
}
void LaunchStripeMeta(const GroupBuffers& gb, int frames, unsigned long long* meta, cudaStream_t s) {
  stripe_meta_kernel<<<1, 32, 0, s>>>(gb, frames, meta);
}
void LaunchCopySmall(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return;
  unsigned grid = cdiv(bytes, 256 * 8);
  if (grid > 32) grid = 32;
  copy_small_kernel<<<grid, 256, 0, s>>>(static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), bytes);
}
void LaunchStripeCompact(const uint8_t* group_out, size_t out_pitch, int first, int frames, const unsigned long long* meta_all_local,
                         uint8_t* dst, cudaStream_t s) {
  stripe_compact_kernel<<<dim3(16, frames), 256, 0, s>>>(group_out, out_pitch, first, meta_all_local, dst);
}

void LaunchStuff(const FrameSet& fs, const GroupBuffers& gb, const StuffArgs& args, cudaStream_t s) {
  // persistent CTAs claiming tiles from a counter (no co-residency requirement): about one SM-full
  // of threads per SM over all pictures of the group
  const size_t max_tiles = (gb.words_pitch * 4 + kStuffTileBytes - 1) / kStuffTileBytes;
  unsigned grid = SmCount() * (1024 / kStuffThreads) / (fs.frames > 0 ? fs.frames : 1);
  if (grid < 1) grid = 1;
  if (grid > max_tiles) grid = static_cast<unsigned>(max_tiles);
  stuff_kernel<<<dim3(grid, fs.frames), kStuffThreads, 0, s>>>(gb, args);
}

}  // namespace sjb
