// block_ops.cuh -- per-8x8-block integer arithmetic of the encode path, written once as
// __host__ __device__ inline functions so that the exact same code runs inside the sm_100a
// kernels (kernels.cu) and, compiled by g++, inside the CPU emulation used by the no-GPU tests
// (tests/emul/emul_main.cc).  Everything here is int32 arithmetic that reproduces the reference's
// scalar C path bit for bit:
//   colour conversion   /root/reference/src/colors_rgb.cc:785-879
//   integer fDCT        /root/reference/src/fdct.cc:67-209, 596-609
//   quantiser           /root/reference/src/quantize.cc:116-148, 288-320
//   DC diff / run-level /root/reference/src/entropy.cc:133-198
//   bit accumulator     /root/reference/src/bit_writer.h:172-209
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SJB_HD __host__ __device__ __forceinline__
#else
#define SJB_HD inline
#endif

namespace sjb {

enum { kYuv420 = 1, kYuv444 = 3, kYuv400 = 4 };   // numbering of sjpeg.h:54-60
enum { kFmtRGB = 0, kFmtBGRA = 1, kFmtRGBA = 2 };

// Per-matrix quantiser constants.  For a coefficient x, with
// s = x >> 31 (0 or -1):
//   q = ((x * iq + (cpos ^ s)) >> 20) - s             (arithmetic shift)
// equals sign(x) * (((|x| + bias) * iquant) >> 16 >> 4) of quantize.cc:116-121, where
// cpos = bias * iquant: for x < 0 the addend becomes -cpos - 1 and the "+1" after the shift turns
// the floor into the ceiling that negating the magnitude quotient needs.  q != 0 <=> |x| >= qthresh
// (quantize.cc:144-147), so the reference's threshold test needs no separate compare.
// Entries are stored by ZIG-ZAG position (entry i belongs to natural index zigzag[i]) as
// {iq, cpos} pairs, so that two consecutive output coefficients share one 16-byte load.
struct QuantTab {
  int32_t e[64][2];
};
struct QuantTabs {
  QuantTab m[2];   // 0 = luma, 1 = chroma
};

// Huffman code tables, packed as the reference does: (code << 16) | length  (entropy.cc:98-112)
struct CodeTabs {
  uint32_t dc[2][16];    // 12 used
  uint32_t ac[2][256];
};

// zig-zag index -> natural index (T.81 figure 5; quantize.cc:32-41)
#define SJB_ZIGZAG_INIT                                                                       \
  { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, \
    20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,  \
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 }

// ---------------------------------------------------------------------------------------------
// Coefficient storage in HBM: sector-interleaved.  A block is 64 int16 = four 32-byte sectors
// (sector s = positions 16s..16s+15, zig-zag order once quantised).  Four consecutive blocks form
// a group of four 128-byte lines, and line s of the group holds sector s of each of its blocks:
//   int16 offset of (block g, position p) = (g / 4) * 256 + (p / 16) * 64 + (g % 4) * 16 + p % 16
// Photographic pictures quantise to non-zeros in sector 0 only (99.97 % of the blocks of the 4K
// benchmark picture), and with one 128-byte line per block every such block cost the entropy
// stage a whole line of DRAM / L2 / L1 for 32 useful bytes (ncu: 128 bytes of DRAM reads per
// block).  Interleaved, the lines that hold the sector-0s are dense and the others are never
// touched.  Kernels address a block through coef_block_base() and step kCoefSectorStride int16
// from one sector to the next; a picture's array is padded to a multiple of four blocks.
// ---------------------------------------------------------------------------------------------
enum { kCoefSectorStride = 64 };   // int16 units between two sectors of the same block
SJB_HD size_t coef_block_base(size_t g) {   // block indices fit 32 bits (<= 3 * 8192 * 8192 blocks)
  const uint32_t b = static_cast<uint32_t>(g);
  return (static_cast<size_t>(b >> 2) << 8) | ((b & 3u) << 4);
}
SJB_HD int coef_pos_offset(int pos) { return pos + (pos >> 4) * 48; }   // = (pos / 16) * 64 + pos % 16
// index, in 16-byte units from the block base, of chunk c (positions 8c..8c+7)
SJB_HD int coef_chunk_index(int c) { return ((c >> 1) << 3) + (c & 1); }
SJB_HD size_t coef_padded_blocks(size_t nb) { return (nb + 3) & ~static_cast<size_t>(3); }

// ---------------------------------------------------------------------------------------------
// colour conversion.  16-bit fixed point BT.601, colors_rgb.cc:17-32.
// ---------------------------------------------------------------------------------------------
SJB_HD int rgb_to_y(int r, int g, int b) {          // colors_rgb.cc:785-795 ; range [-128,127]
  return (19595 * r + 38469 * g + 7471 * b + (32768 - (128 << 16))) >> 16;
}
SJB_HD int rgb_to_u(int r, int g, int b) {          // colors_rgb.cc:809-819 (4:4:4)
  return (-11059 * r - 21709 * g + 32768 * b + 32768) >> 16;
}
SJB_HD int rgb_to_v(int r, int g, int b) {
  return (32768 * r - 27439 * g - 5329 * b + 32768) >> 16;
}
SJB_HD int rgb4_to_u(int sr, int sg, int sb) {      // colors_rgb.cc:797-806 : sums of a 2x2 quad
  return (-11059 * sr - 21709 * sg + 32768 * sb + 131072) >> 18;
}
SJB_HD int rgb4_to_v(int sr, int sg, int sb) {
  return (32768 * sr - 27439 * sg - 5329 * sb + 131072) >> 18;
}

// ---------------------------------------------------------------------------------------------
// integer fDCT, output = 16 x the JPEG-normalised DCT.  Column pass first (fdct.cc:67-144 with
// the C macros of :150-157), then the row pass with four constant tables (fdct.cc:174-209).
// All intermediate stores of the reference are int16; for samples in [-128,128] every stored
// value stays inside int16 (|DC| <= 16384), so keeping int32 registers is exact.
// ---------------------------------------------------------------------------------------------
SJB_HD int mulhi16(int a, int c) { return (a * c) >> 16; }

SJB_HD void column_dct8(int& x0, int& x1, int& x2, int& x3, int& x4, int& x5, int& x6, int& x7) {
  int m0 = x0 - x7, m7 = x0 + x7;
  int m2 = x2 - x5, m5 = x2 + x5;
  int m3 = x3 - x4, m4 = x3 + x4;
  int m1 = x1 - x6, m6 = x1 + x6;
  { const int t = m7 - m4; m4 = m7 + m4; m7 = t; }
  { const int t = m6 - m5; m5 = m6 + m5; m6 = t; }
  m4 <<= 3; m5 <<= 3;
  x0 = m4 + m5;
  x4 = m4 - m5;
  m7 <<= 3; m6 <<= 3; m3 <<= 3; m0 <<= 3;
  x2 = mulhi16(27146, m6) + m7;                   // kTan2
  x6 = mulhi16(27146, m7) - m6;
  m2 <<= 4; m1 <<= 4;
  { const int t = m1 - m2; m2 = m1 + m2; m1 = t; }
  m2 = mulhi16(m2, 23170);                        // k2Sqrt2
  m1 = mulhi16(m1, 23170);
  { const int t = m3 - m1; m1 = m3 + m1; m3 = t; }
  { const int t = m0 - m2; m2 = m0 + m2; m0 = t; }
  const int s3 = m3, s1 = m1;
  m3 = mulhi16(m3, -21746) + s3 + 1;              // kTan3m1, CORRECT_LSB
  m1 = mulhi16(m1, 13036) + m2 + 1;               // kTan1,   CORRECT_LSB
  const int t4 = mulhi16(-21746, m0) + m0;
  const int t5 = mulhi16(13036, m2);
  x1 = m1;
  x3 = m0 - m3;
  x5 = s3 + t4;
  x7 = t5 - s1;
}

// row tables of fdct.cc:28-35, selected at compile time
template <int T> struct RowTab;
template <> struct RowTab<0> { enum { C1 = 22725, C2 = 21407, C3 = 19266, C4 = 16384, C5 = 12873, C6 = 8867, C7 = 4520 }; };
template <> struct RowTab<1> { enum { C1 = 31521, C2 = 29692, C3 = 26722, C4 = 22725, C5 = 17855, C6 = 12299, C7 = 6270 }; };
template <> struct RowTab<2> { enum { C1 = 29692, C2 = 27969, C3 = 25172, C4 = 21407, C5 = 16819, C6 = 11585, C7 = 5906 }; };
template <> struct RowTab<3> { enum { C1 = 26722, C2 = 25172, C3 = 22654, C4 = 19266, C5 = 15137, C6 = 10426, C7 = 5315 }; };

template <int T>
SJB_HD void row_dct8(int& x0, int& x1, int& x2, int& x3, int& x4, int& x5, int& x6, int& x7) {
  typedef RowTab<T> K;
  const int a0 = x0 + x7, b0 = x0 - x7;
  const int a1 = x1 + x6, b1 = x1 - x6;
  const int a2 = x2 + x5, b2 = x2 - x5;
  const int a3 = x3 + x4, b3 = x3 - x4;
  const int c0 = a0 + a3, c1 = a0 - a3, c2 = a1 + a2, c3 = a1 - a2;
  x0 = (K::C4 * (c0 + c2)) >> 16;
  x4 = (K::C4 * (c0 - c2)) >> 16;
  x2 = (K::C2 * c1 + K::C6 * c3) >> 16;
  x6 = (K::C6 * c1 - K::C2 * c3) >> 16;
  x1 = (K::C1 * b0 + K::C3 * b1 + K::C5 * b2 + K::C7 * b3) >> 16;
  x3 = (K::C3 * b0 - K::C7 * b1 - K::C1 * b2 - K::C5 * b3) >> 16;
  x5 = (K::C5 * b0 - K::C1 * b1 + K::C7 * b2 + K::C3 * b3) >> 16;
  x7 = (K::C7 * b0 - K::C5 * b1 + K::C3 * b2 - K::C1 * b3) >> 16;
}

// whole block, natural order v[8*row + col], in place.  fdct.cc:596-609
SJB_HD void fdct64(int (&v)[64]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int c = 0; c < 8; ++c) {
    column_dct8(v[c], v[8 + c], v[16 + c], v[24 + c], v[32 + c], v[40 + c], v[48 + c], v[56 + c]);
  }
  row_dct8<0>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
  row_dct8<1>(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
  row_dct8<2>(v[16], v[17], v[18], v[19], v[20], v[21], v[22], v[23]);
  row_dct8<3>(v[24], v[25], v[26], v[27], v[28], v[29], v[30], v[31]);
  row_dct8<0>(v[32], v[33], v[34], v[35], v[36], v[37], v[38], v[39]);
  row_dct8<3>(v[40], v[41], v[42], v[43], v[44], v[45], v[46], v[47]);
  row_dct8<2>(v[48], v[49], v[50], v[51], v[52], v[53], v[54], v[55]);
  row_dct8<1>(v[56], v[57], v[58], v[59], v[60], v[61], v[62], v[63]);
}

// ---------------------------------------------------------------------------------------------
// quantiser (see QuantTab)
// ---------------------------------------------------------------------------------------------
SJB_HD int quantize_coeff(int x, int iq, int cpos) {
  const int s = x >> 31;
  return ((x * iq + (cpos ^ s)) >> 20) - s;
}

// number of bits of v >= 0 (0 for v == 0); sjpegi.h:186-198
SJB_HD int bit_length(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return 32 - __clz((int)v);
#else
  return v ? 32 - __builtin_clz(v) : 0;
#endif
}
SJB_HD int find_first_set64(uint64_t m) {   // index of lowest set bit, m != 0
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)m) - 1;
#else
  return __builtin_ctzll(m);
#endif
}

// JPEG "size + amplitude bits" of a non-zero value: negative numbers are sent as value-1 in n
// bits (one's complement).  quantize.cc:298-303, entropy.cc:133-150.
SJB_HD void size_and_bits(int v, int* n, uint32_t* bits) {
  const int m = v >> 31;
  const uint32_t a = (uint32_t)((v ^ m) - m);
  *n = bit_length(a);
  *bits = (uint32_t)(v + m) & ((1u << *n) - 1u);
}

// ---------------------------------------------------------------------------------------------
// Block -> Huffman symbols.  The block is 64 quantised int16 in zig-zag order, seen as 8 chunks
// of 16 bytes (chunk c = positions 8c..8c+7 = four little-endian words of two values each);
// chunkmask bit c is set when chunk c holds a non-zero value (chunk 0 also holds the DC), so a
// coder only loads the chunks that matter.  dc_pred = quantised DC of the previous block of the
// same component.  Sink receives (code_bits, length) in stream order; the two uses are bit counting
// and packing.  Follows entropy.cc:161-198 (CodeBlock) with run/levels recomputed on the fly as
// quantize.cc:288-320 emits them.
// ---------------------------------------------------------------------------------------------
SJB_HD int find_first_set32(uint32_t m) {   // index of lowest set bit, m != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)m) - 1;
#else
  return __builtin_ctz(m);
#endif
}

struct Words4 {
  uint32_t w[4];
};

// 8 flag bits of a chunk: bit k <=> value k of the chunk (zig-zag position 8c + k) is non-zero
SJB_HD uint32_t chunk_nonzero_bits(const Words4& q) {
  uint32_t m = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < 4; ++j) {
    const uint32_t w = q.w[j];
    const uint32_t lo = (w & 0xffffu) ? 1u : 0u, hi = (w >> 16) ? 2u : 0u;
    m |= (lo | hi) << (2 * j);
  }
  return m;
}

// The walk has two phases so that the lanes of a warp stay converged: first the map of non-zero
// positions (two 32-bit words) is assembled from the chunks the bitmap names (a short loop,
// usually one or two chunks), then ONE loop iteration per non-zero coefficient emits its run/size
// symbol -- the trip count of a warp is the largest number of non-zeros among its 32 blocks (per
// half of the block), not the number of (chunk, slot) pairs any of them touches.  Loader: operator()(chunk) -> Words4, value(pos) -> the
// quantised value at zig-zag position pos.
// map of the non-zero AC positions of a block (bits 1..31 of *lo = positions 1..31, *hi = 32..63),
// assembled from the chunks the bitmap names (a short loop, usually one or two chunks)
template <class Loader>
SJB_HD void block_nz_maps(const Loader& load, uint32_t chunkmask, uint32_t* lo, uint32_t* hi) {
  uint32_t nz_lo = 0, nz_hi = 0;
  for (uint32_t m = chunkmask; m; m &= m - 1) {
    const int c = find_first_set32(m);
    const uint32_t b8 = chunk_nonzero_bits(load(c)) << (8 * (c & 3));
    if (c < 4) nz_lo |= b8; else nz_hi |= b8;
  }
  *lo = nz_lo & ~1u;                       // position 0 is the DC
  *hi = nz_hi;
}

// DC symbol, then ONE loop iteration per non-zero coefficient, then EOB
template <class Loader, class Sink>
SJB_HD void code_block_mapped(const Loader& load, uint32_t nz_lo, uint32_t nz_hi, int dc, int dc_pred,
                              const uint32_t* dc_codes, const uint32_t* ac_codes, Sink& sink) {
  {
    const int diff = dc - dc_pred;
    int n = 0;
    uint32_t bits = 0;
    if (diff != 0) size_and_bits(diff, &n, &bits);
    const uint32_t c = dc_codes[n];
    // code then n suffix bits; at most 16 + 11 bits
    sink.put(((c >> 16) << n) | bits, (int)(c & 0xff) + n);
  }
  const uint32_t zrl = ac_codes[0xf0];
  int prev = 0;                            // zig-zag position of the previous non-zero (0 = DC slot)
  // ONE loop over both halves of the map (positions 1..31, then 32..63): its trip count is the
  // block's number of non-zeros, so a warp whose blocks carry similar numbers stays converged even
  // when they split differently between the halves (two loops, one per half, ran max + max).
  uint32_t m = nz_lo, rest = nz_hi;
  int base = 0;
  if (m == 0) {
    m = rest;
    rest = 0;
    base = 32;
  }
  if (m) {
    // software pipeline: the value of the NEXT non-zero is requested before this one is coded
    int next_pos = base + find_first_set32(m);
    int next_v = load.value(next_pos);
    for (;;) {
      const int pos = next_pos;
      const int v = next_v;
      m &= m - 1;
      if (m == 0) {                        // first half exhausted: go on with the second (if any)
        m = rest;
        rest = 0;
        base = 32;
      }
      const bool more = m != 0;
      if (more) {
        next_pos = base + find_first_set32(m);
        next_v = load.value(next_pos);
      }
      int run = pos - prev - 1;
      prev = pos;
      while (run >= 16) {                  // ZRL escapes, entropy.cc:176-179
        sink.put(zrl >> 16, (int)(zrl & 0xff));
        run -= 16;
      }
      int n;
      uint32_t bits;
      size_and_bits(v, &n, &bits);
      const uint32_t c = ac_codes[(run << 4) | n];
      sink.put(((c >> 16) << n) | bits, (int)(c & 0xff) + n);
      if (!more) break;
    }
  }
  if (prev < 63) {                         // EOB, entropy.cc:195-197
    const uint32_t c = ac_codes[0x00];
    sink.put(c >> 16, (int)(c & 0xff));
  }
}

template <class Loader, class Sink>
SJB_HD void code_block(Loader load, uint32_t chunkmask, int dc, int dc_pred, const uint32_t* dc_codes,
                       const uint32_t* ac_codes, Sink& sink) {
  // positions 0..31 and 32..63 as two 32-bit maps: cheaper to scan than one 64-bit word
  uint32_t nz_lo, nz_hi;
  block_nz_maps(load, chunkmask, &nz_lo, &nz_hi);
  code_block_mapped(load, nz_lo, nz_hi, dc, dc_pred, dc_codes, ac_codes, sink);
}

struct BitCountSink {
  uint32_t total;
  SJB_HD void put(uint32_t, int len) { total += (uint32_t)len; }
};

// symbol statistics of a block (entropy.cc:208-227); Add(table_slot) where slot < 256 is an AC
// symbol and 256 + n a DC size.  (Walking by non-zero like the coder -- map first, one value load
// per non-zero -- was measured SLOWER here: 4K gen A 217 -> 257 us per 16 pictures, 4:4:4 q90 220 ->
// 305; the chunk walk below keeps its values in registers and the counters' atomics dominate.)
template <class LoadChunk, class Add>
SJB_HD void block_symbol_stats(LoadChunk load_chunk, uint32_t chunkmask, int dc, int dc_pred, Add& add) {
  {
    const int diff = dc - dc_pred;
    const int m = diff >> 31;
    add.one(256 + bit_length((uint32_t)((diff ^ m) - m)));
  }
  uint32_t m = chunkmask;
  int prev = 0;
  while (m) {
    const int c = find_first_set32(m);
    m &= m - 1;
    const Words4 q = load_chunk(c);
    for (int j = 0; j < 4; ++j) {
      const uint32_t w = q.w[j];
      const int p = 4 * c + j;
      const int v2[2] = {(int16_t)(w & 0xffffu), (int32_t)w >> 16};
      for (int e = (p == 0) ? 1 : 0; e < 2; ++e) {
        const int v = v2[e];
        if (v == 0) continue;
        const int pos = 2 * p + e;
        const int run = pos - prev - 1;
        prev = pos;
        if (run >> 4) add.many(0xf0, run >> 4);
        const int s = v >> 31;
        add.one(((run & 15) << 4) | bit_length((uint32_t)((v ^ s) - s)));
      }
    }
  }
  if (prev < 63) add.one(0x00);
}

// ---------------------------------------------------------------------------------------------
// MSB-first bit packing into 32-bit words (word w holds stream bits 32w..32w+31, bit 32w in the
// MSB).  A block starts at an arbitrary bit offset: its first and last words are shared with the
// neighbouring blocks and are merged with OR; words in between are owned and stored plainly.
// Out must provide or_word(index, value) and set_word(index, value).
// ---------------------------------------------------------------------------------------------
template <class Out>
struct BitPackSink {
  Out& out;
  uint64_t acc;      // pending bits, top aligned
  int n;             // number of pending bits (< 32 between calls), includes the lead-in gap
  uint64_t word;     // index of the word the top of acc belongs to
  bool shared;       // true while the next word to emit is shared with the previous block
  SJB_HD BitPackSink(Out& o, uint64_t bit_offset)
      : out(o), acc(0), n((int)(bit_offset & 31)), word(bit_offset >> 5), shared((bit_offset & 31) != 0) {}
  SJB_HD void put(uint32_t bits, int len) {     // len <= 27, bits < 2^len
    acc |= (uint64_t)bits << (64 - n - len);
    n += len;
    if (n >= 32) {
      const uint32_t w = (uint32_t)(acc >> 32);
      if (shared) out.or_word(word, w); else out.set_word(word, w);
      shared = false;
      acc <<= 32;
      n -= 32;
      ++word;
    }
  }
  SJB_HD void finish() {
    if (n > 0) out.or_word(word, (uint32_t)(acc >> 32));
  }
};

// ---------------------------------------------------------------------------------------------
// Trellis quantiser of one block (quantize.cc:325-457): rate-distortion
// dynamic programme over the non-zero zig-zag positions, two candidate levels per position,
// uint32 scores that wrap exactly like the reference's score_t.
//   in      raw x16 coefficients, natural order      qm     8-bit matrix, natural order
//   qtab    {iq, cpos} by zig-zag position           ac_len code lengths of the DEFAULT AC table
//   out     quantised values, zig-zag order (all 64 written); returns the chunk bitmap
// Working storage is passed in (local memory on the device): structure-of-arrays nodes, 10 bytes
// each -- score | packed {pos:6, nbits:4, prev:7, rank:7, run:6} | amplitude code.
// ---------------------------------------------------------------------------------------------
struct TrellisScratch {
  uint32_t score[1 + 2 * 63];
  uint32_t meta[1 + 2 * 63];
  uint16_t code[1 + 2 * 63];
  uint32_t disto0[64];
};
SJB_HD uint32_t trellis_meta(int pos, int nbits, int prev, int rank, int run) {
  return (uint32_t)pos | ((uint32_t)nbits << 6) | ((uint32_t)prev << 10) | ((uint32_t)rank << 17) | ((uint32_t)run << 24);
}
SJB_HD uint32_t trellis_block(const int16_t* in, const uint8_t* qm, const int32_t (*qtab)[2], const uint8_t* ac_len,
                              int16_t* out, TrellisScratch& S) {
  const int zz[64] = SJB_ZIGZAG_INIT;
  S.score[0] = 0;
  S.meta[0] = 0;                 // the sink: position 0, rank 0
  S.code[0] = 0;
  S.disto0[0] = 0;
  int cur = 1;
  for (int i = 1; i < 64; ++i) {
    const int j = zz[i];
    const uint32_t q = (uint32_t)qm[j] << 4;
    const uint32_t lambda = q * q / 32u;
    const int x = in[j];
    const int sign = x >> 31;
    const int V = (x ^ sign) - sign;
    S.disto0[i] = (uint32_t)(V * V) + S.disto0[i - 1];
    int v = (V * qtab[i][0] + qtab[i][1]) >> 20;   // V >= 0
    if (v == 0) continue;
    int nbits = bit_length((uint32_t)v);
    for (int k = 0; k < 2; ++k) {
      const int err = V - v * (int)q;
      const uint32_t base = (uint32_t)(err * err) + S.disto0[i - 1];
      // SearchBestPrev, quantize.cc:350-383: walk back from the newest kept node to the sink
      uint32_t best_score = 0xffffffffu;
      int best_prev = -1, best_rank = 0, best_run = 0;
      for (int p = cur - 1; p >= 0; --p) {
        const uint32_t pm = S.meta[p];
        const int ppos = (int)(pm & 63u);
        const int run = i - 1 - ppos;
        if (run < 0) continue;                       // the other candidate of this same position
        uint32_t bits = (uint32_t)nbits + (uint32_t)(run >> 4) * ac_len[0xf0];
        const uint32_t d = base - S.disto0[ppos];
        if (d + lambda * bits >= best_score) break;   // exact early-out of the reference
        bits += ac_len[((run & 15) << 4) | nbits];
        const uint32_t score = d + lambda * bits + S.score[p];
        if (score < best_score) {
          best_score = score;
          best_prev = p;
          best_rank = (int)((pm >> 17) & 127u) + 1;
          best_run = run;
        }
      }
      if (best_prev >= 0) {                           // a candidate without predecessor is dropped
        S.score[cur] = best_score;
        S.meta[cur] = trellis_meta(i, nbits, best_prev, best_rank, best_run);
        S.code[cur] = (uint16_t)((v ^ sign) & ((1 << nbits) - 1));
        ++cur;
      }
      --nbits;
      if (nbits <= 0) break;
      v = (1 << nbits) - 1;
    }
  }
  int best = 0;
  if (cur != 1) {
    uint32_t best_score = 0xffffffffu;
    for (int p = cur - 1; p >= 0; --p) {         // the sink takes part (quantize.cc:430: all-zero block)
      const uint32_t s = S.score[p] + (S.disto0[63] - S.disto0[S.meta[p] & 63u]);
      if (s < best_score) { best = p; best_score = s; }
    }
  }
  for (int i = 0; i < 64; ++i) out[i] = 0;
  out[0] = (int16_t)quantize_coeff(in[0], qtab[0][0], qtab[0][1]);      // DC: plain quantiser
  uint32_t mask = (out[0] != 0) ? 1u : 0u;
  for (int p = best; p > 0; p = (int)((S.meta[p] >> 10) & 127u)) {
    const uint32_t pm = S.meta[p];
    const int pos = (int)(pm & 63u), n = (int)((pm >> 6) & 15u);
    const int amp = S.code[p];
    out[pos] = (int16_t)((amp >> (n - 1)) ? amp : amp - ((1 << n) - 1));
    mask |= 1u << (pos >> 3);
  }
  return mask;
}

// ---------------------------------------------------------------------------------------------
// Trellis quantiser, second form: the same dynamic programme laid out for a GPU thread whose
// working storage is a 1 KB private column of SHARED memory (no local memory, no data-dependent
// register indexing).  What changes against trellis_block() is bookkeeping only:
//   * nodes are addressed by (zig-zag position, candidate) instead of a compacted list: at most
//     two candidates exist per position (quantize.cc:386), so node id = 2 * pos + k; which nodes
//     were kept is two 64-bit masks in registers.  The reference walks its list from the newest
//     kept node down to the sink (quantize.cc:357): that is descending position, candidate 1
//     before candidate 0, the sink (position 0) last -- the order the masks are scanned in.
//   * a first, branch-free pass over the 63 AC positions leaves the prefix sums of V^2 (disto0)
//     and the map of positions that quantise to non-zero; the second pass visits only those.
//   * the two candidates of a position see exactly the same predecessors (the first candidate is
//     skipped by the second: run < 0, quantize.cc:362), so both walks share one loop and one set
//     of loads; each keeps its own running best and its own early-out (quantize.cc:367).
//   * a node stores its score and its predecessor only; level and sign are re-derived from the
//     coefficient when the best path is read back.
// Scores are uint32 and wrap as the reference's score_t does.
// Mem: private storage.  Tab: per-matrix constants by ZIG-ZAG position.
//   Mem: coef_ld(w) / coef_st(w, v)        32 words: the block, two int16 per word
//        disto_ld(i) / disto_st(i, v)      64 words
//        score_ld(pos, &s0, &s1) / score_st(pos, k, v)
//        prev_ld(node) / prev_st(node, v)  128 bytes
//        out_zero() / out_st(pos, v) / out_ld(w)   32 words (may alias disto)
//   Tab: qt(i, &iq, &cpos), q16(i) = matrix entry << 4, len(sym) = AC code length
// raw: the 64 unquantised x16 coefficients in natural order, two per word (low half first).
// Result: quantised block in zig-zag order in Mem::out (two per word); returns the chunk bitmap.
// ---------------------------------------------------------------------------------------------
SJB_HD int sjb_half(uint32_t w, int hi) { return hi ? ((int32_t)w >> 16) : ((int32_t)(w << 16) >> 16); }
SJB_HD uint32_t sjb_clz32(uint32_t m) {     // leading zeros, m != 0
#if defined(__CUDA_ARCH__)
  return (uint32_t)__clz((int)m);
#else
  return (uint32_t)__builtin_clz(m);
#endif
}
SJB_HD int find_last_set64(uint64_t m) {   // index of highest set bit, m != 0
#if defined(__CUDA_ARCH__)
  return 63 - __clzll((long long)m);
#else
  return 63 - __builtin_clzll(m);
#endif
}

template <class Mem, class Tab>
SJB_HD uint32_t trellis_block_v2(const uint32_t (&raw)[32], Mem& M, const Tab& T) {
#if defined(__CUDA_ARCH__)
  constexpr int zz[64] = SJB_ZIGZAG_INIT;
#else
  const int zz[64] = SJB_ZIGZAG_INIT;
#endif
  // ---- pass 1: prefix sums of V^2, map of non-zero positions, block re-stored in zig-zag order ----
  uint64_t nz = 0;
  {
    uint32_t acc = 0;
    M.disto_st(0, 0u);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 1; i < 64; ++i) {
      const int x = sjb_half(raw[zz[i] >> 1], zz[i] & 1);
      const int sg = x >> 31;
      const int V = (x ^ sg) - sg;
      acc += (uint32_t)(V * V);
      M.disto_st(i, acc);
      int iq, cpos;
      T.qt(i, iq, cpos);
      if (V * iq + cpos >= (1 << 20)) nz |= 1ull << i;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int w = 0; w < 32; ++w) {
      const uint32_t lo = (uint32_t)sjb_half(raw[zz[2 * w] >> 1], zz[2 * w] & 1) & 0xffffu;
      const uint32_t hi = (uint32_t)sjb_half(raw[zz[2 * w + 1] >> 1], zz[2 * w + 1] & 1);
      M.coef_st(w, lo | (hi << 16));
    }
  }
  // ---- pass 2: the dynamic programme over the non-zero positions ----
  uint64_t kept0 = 1, kept1 = 0;          // bit pos <=> node (pos, k) exists; (0, 0) is the sink
  M.score_st(0, 0, 0u);
  const uint32_t zrl_len = T.len(0xf0);
  for (uint64_t rest = nz; rest; rest &= rest - 1) {
    const int i = find_first_set64(rest);
    const int x = sjb_half(M.coef_ld(i >> 1), i & 1);
    const int sg = x >> 31;
    const int V = (x ^ sg) - sg;
    int iq, cpos;
    T.qt(i, iq, cpos);
    const int q = T.q16(i);
    const uint32_t lambda = (uint32_t)(q * q) / 32u;
    const int vA = (V * iq + cpos) >> 20;                    // >= 1
    const int nbA = bit_length((uint32_t)vA);
    const uint32_t dprev = M.disto_ld(i - 1);
    const int eA = V - vA * q;
    const uint32_t baseA = (uint32_t)(eA * eA) + dprev;
    const bool hasB = nbA > 1;                               // second candidate: (1 << (nbits-1)) - 1
    const int nbB = nbA - 1;
    const int eB = V - ((1 << nbB) - 1) * q;
    const uint32_t baseB = (uint32_t)(eB * eB) + dprev;
    uint32_t bestA = 0xffffffffu, bestB = 0xffffffffu;
    int bpA = -1, bpB = -1;
    bool actA = true, actB = hasB;
    // Predecessors: the kept nodes at positions below i, newest first; the sink (bit 0) is always
    // among them.  The loop is software-pipelined: the loads of the NEXT predecessor (prefix sum,
    // both scores, both code lengths) are issued before the current one is evaluated, which takes
    // the shared-memory latency off the dependent chain of compares (few warps are resident, so
    // that chain is what bounds the kernel).  A load past the end re-reads the sink: harmless.
    // The position masks are walked as two 32-bit halves, and the evaluation of a node is written
    // without branches (selects on the two early-out / improvement conditions): on the device the
    // loop body is straight-line code for the 32 blocks of a warp.
    const uint32_t below_hi = (i > 32) ? ((1u << (i - 32)) - 1u) : 0u;
    const uint32_t below_lo = (i >= 32) ? 0xffffffffu : ((1u << i) - 1u);
    uint32_t mh = (static_cast<uint32_t>(kept0 >> 32) | static_cast<uint32_t>(kept1 >> 32)) & below_hi;
    uint32_t ml = (static_cast<uint32_t>(kept0) | static_cast<uint32_t>(kept1)) & below_lo;
    // next predecessor position (and whether its candidates 0 / 1 exist); (mh | ml) != 0 on entry
    int pp;
    uint32_t hbits;                        // bit 0: candidate 0 kept, bit 1: candidate 1 kept
    {
      const bool in_hi = mh != 0;
      const uint32_t mm = in_hi ? mh : ml;
      const int bpos = 31 - (int)sjb_clz32(mm);
      if (in_hi) mh ^= 1u << bpos; else ml ^= 1u << bpos;
      pp = bpos + (in_hi ? 32 : 0);
      const uint32_t c0 = in_hi ? static_cast<uint32_t>(kept0 >> 32) : static_cast<uint32_t>(kept0);
      const uint32_t c1 = in_hi ? static_cast<uint32_t>(kept1 >> 32) : static_cast<uint32_t>(kept1);
      hbits = ((c0 >> bpos) & 1u) | (((c1 >> bpos) & 1u) << 1);
    }
    uint32_t d0 = M.disto_ld(pp), s0, s1;
    M.score_ld(pp, s0, s1);
    uint32_t lenA = T.len((((i - 1 - pp) & 15) << 4) | nbA), lenB = T.len((((i - 1 - pp) & 15) << 4) | nbB);
    bool more = true;
    while (more && (actA || actB)) {
      more = (mh | ml) != 0;
      int np = 0;
      uint32_t nhbits = 0;
      {
        const bool in_hi = mh != 0;
        const uint32_t mm = in_hi ? mh : ml;
        const int bpos = more ? 31 - (int)sjb_clz32(mm) : 0;
        const uint32_t bit = more ? (1u << bpos) : 0u;
        if (in_hi) mh ^= bit; else ml ^= bit;
        np = bpos + (in_hi ? 32 : 0);
        const uint32_t c0 = in_hi ? static_cast<uint32_t>(kept0 >> 32) : static_cast<uint32_t>(kept0);
        const uint32_t c1 = in_hi ? static_cast<uint32_t>(kept1 >> 32) : static_cast<uint32_t>(kept1);
        nhbits = ((c0 >> bpos) & 1u) | (((c1 >> bpos) & 1u) << 1);
      }
      const uint32_t nd0 = M.disto_ld(np);
      uint32_t ns0, ns1;
      M.score_ld(np, ns0, ns1);
      const uint32_t nlenA = T.len((((i - 1 - np) & 15) << 4) | nbA), nlenB = T.len((((i - 1 - np) & 15) << 4) | nbB);
      const bool h0 = (hbits & 1u) != 0, h1 = (hbits & 2u) != 0;
      const uint32_t zr = (uint32_t)((i - 1 - pp) >> 4) * zrl_len;
      {
        // candidate A against node (pp, 1), then (pp, 0): quantize.cc:357-382 for two consecutive nodes
        const uint32_t thr = baseA - d0 + lambda * ((uint32_t)nbA + zr);
        const uint32_t full = thr + lambda * lenA;
        const bool stop1 = actA && h1 && thr >= bestA;
        const uint32_t sc1 = full + s1;
        const bool take1 = actA && h1 && !stop1 && sc1 < bestA;
        bestA = take1 ? sc1 : bestA;
        bpA = take1 ? 2 * pp + 1 : bpA;
        actA = actA && !stop1;
        const bool stop0 = actA && h0 && thr >= bestA;
        const uint32_t sc0 = full + s0;
        const bool take0 = actA && h0 && !stop0 && sc0 < bestA;
        bestA = take0 ? sc0 : bestA;
        bpA = take0 ? 2 * pp : bpA;
        actA = actA && !stop0;
      }
      {
        const uint32_t thr = baseB - d0 + lambda * ((uint32_t)nbB + zr);
        const uint32_t full = thr + lambda * lenB;
        const bool stop1 = actB && h1 && thr >= bestB;
        const uint32_t sc1 = full + s1;
        const bool take1 = actB && h1 && !stop1 && sc1 < bestB;
        bestB = take1 ? sc1 : bestB;
        bpB = take1 ? 2 * pp + 1 : bpB;
        actB = actB && !stop1;
        const bool stop0 = actB && h0 && thr >= bestB;
        const uint32_t sc0 = full + s0;
        const bool take0 = actB && h0 && !stop0 && sc0 < bestB;
        bestB = take0 ? sc0 : bestB;
        bpB = take0 ? 2 * pp : bpB;
        actB = actB && !stop0;
      }
      pp = np; hbits = nhbits; d0 = nd0; s0 = ns0; s1 = ns1; lenA = nlenA; lenB = nlenB;
    }
    if (bpA >= 0) {                                          // a candidate without predecessor is dropped
      M.score_st(i, 0, bestA);
      M.prev_st(2 * i, (uint32_t)bpA);
      kept0 |= 1ull << i;
    }
    if (hasB && bpB >= 0) {
      M.score_st(i, 1, bestB);
      M.prev_st(2 * i + 1, (uint32_t)bpB);
      kept1 |= 1ull << i;
    }
  }
  // ---- best end node (quantize.cc:427-441): newest first, strict <, the sink takes part ----
  int best = 0;
  if ((kept0 | kept1) != 1ull) {
    uint32_t best_score = 0xffffffffu;
    const uint32_t d63 = M.disto_ld(63);
    for (uint64_t m = kept0 | kept1; m != 0;) {
      const int pp = find_last_set64(m);
      m ^= 1ull << pp;
      const uint32_t tail = d63 - M.disto_ld(pp);
      uint32_t s0, s1;
      M.score_ld(pp, s0, s1);
      if ((kept1 >> pp) & 1) { const uint32_t sc = s1 + tail; if (sc < best_score) { best_score = sc; best = 2 * pp + 1; } }
      if ((kept0 >> pp) & 1) { const uint32_t sc = s0 + tail; if (sc < best_score) { best_score = sc; best = 2 * pp; } }
    }
  }
  // ---- read the path back; levels re-derived from the coefficients ----
  int dcq;
  {
    int iq, cpos;
    T.qt(0, iq, cpos);
    dcq = quantize_coeff(sjb_half(raw[0], 0), iq, cpos);     // DC: plain quantiser
  }
  uint32_t mask = (dcq != 0) ? 1u : 0u;
  // positions on the path, collected as a map; their candidate bits alongside
  uint64_t on_path = 0, path_k = 0;
  for (int node = best; node != 0; node = (int)M.prev_ld(node)) {
    on_path |= 1ull << (node >> 1);
    path_k |= (uint64_t)(node & 1) << (node >> 1);
  }
  M.out_zero();                                              // may overwrite disto: not needed any more
  M.out_st(0, dcq);
  for (uint64_t m = on_path; m; m &= m - 1) {
    const int pos = find_first_set64(m);
    const int x = sjb_half(M.coef_ld(pos >> 1), pos & 1);
    const int sg = x >> 31;
    const int V = (x ^ sg) - sg;
    int iq, cpos;
    T.qt(pos, iq, cpos);
    int v = (V * iq + cpos) >> 20;
    if ((path_k >> pos) & 1) v = (1 << (bit_length((uint32_t)v) - 1)) - 1;
    M.out_st(pos, (v ^ sg) - sg);
    mask |= 1u << (pos >> 3);
  }
  return mask;
}

// plain-array storage and tables for host code (CPU emulation of the kernel, tests/emul)
struct TrellisHostMem {
  uint32_t coef[32], disto[64], score[64][2], out[32];
  uint8_t prev[128];
  uint32_t coef_ld(int w) const { return coef[w]; }
  void coef_st(int w, uint32_t v) { coef[w] = v; }
  uint32_t disto_ld(int i) const { return disto[i]; }
  void disto_st(int i, uint32_t v) { disto[i] = v; }
  void score_ld(int pos, uint32_t& s0, uint32_t& s1) const { s0 = score[pos][0]; s1 = score[pos][1]; }
  void score_st(int pos, int k, uint32_t v) { score[pos][k] = v; }
  uint32_t prev_ld(int node) const { return prev[node]; }
  void prev_st(int node, uint32_t v) { prev[node] = (uint8_t)v; }
  void out_zero() { for (int i = 0; i < 32; ++i) out[i] = 0; }
  void out_st(int pos, int v) {
    const uint32_t h = (uint32_t)v & 0xffffu;
    out[pos >> 1] = (pos & 1) ? ((out[pos >> 1] & 0xffffu) | (h << 16)) : ((out[pos >> 1] & 0xffff0000u) | h);
  }
};
struct TrellisHostTab {
  const int32_t (*qtab)[2];   // {iq, cpos} by zig-zag position
  const uint8_t* qm;          // 8-bit matrix, natural order
  const uint8_t* ac_len;
  void qt(int i, int& iq, int& cpos) const { iq = qtab[i][0]; cpos = qtab[i][1]; }
  int q16(int i) const { const int zz[64] = SJB_ZIGZAG_INIT; return (int)qm[zz[i]] << 4; }
  uint32_t len(int sym) const { return ac_len[sym]; }
};

SJB_HD uint32_t sjb_minu(uint32_t a, uint32_t b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------
// Adaptive quantisation: analysis of the coefficient histograms (histogram.cc:126-315) in a form
// that runs on the device (kernels.cu: analyse_fit_kernel + analyse_pick_kernel) and, compiled by
// g++, in the CPU emulation.  The arithmetic is the reference's: 32-bit wrapping integer products,
// sums of integers far below 2^53 (exact in double whatever the order), then the SAME sequence of
// double / float operations per position and over the positions in ascending order -- IEEE
// operations without contraction give the same bits on the device as on the host (the library is
// compiled with --fmad=false / -ffp-contract=off).
//   histogram row h: int32[128] counts of |coef| >> 2 for one matrix and position
//   candidate steps: q0 + d - 12 for d = 0..24; weights of the linear fit of distortion and rate
//   against the step; lambda = -cov(dist) / cov(rate) summed over the positions that pass the
//   density and correlation tests; per position the step minimising dist + lambda * rate
// ---------------------------------------------------------------------------------------------
enum { kAqDeltaMin = -12, kAqNumDelta = 25, kAqShift = 2, kAqBins = 128 };
#define SJB_AQ_FLT_MAX 3.402823466e+38f

struct AqFit {                 // what the fit of one position leaves for the pick
  float rate[kAqNumDelta];
  float dist[kAqNumDelta];
  double cov, den;             // this position's terms of the two sums lambda is made of
  int skip;                    // position keeps its step (never touched, too sparse, or uncorrelated)
  int pad;
};

SJB_HD float aq_weight(int d) {
  // histogram.cc:119-124
  return d < 5 || d > 19 ? 0.f
       : d == 5 || d == 19 ? 1.f : d == 6 || d == 18 ? 5.f : d == 7 || d == 17 ? 16.f : d == 8 || d == 16 ? 43.f
       : d == 9 || d == 15 ? 94.f : d == 10 || d == 14 ? 164.f : d == 11 || d == 13 ? 228.f : 255.f;
}
// does candidate d of a position take part (as a fitted point or as a choice)?
SJB_HD bool aq_candidate_used(int q0, int qmin, int delta_top, int d) {
  const int q = q0 + d + kAqDeltaMin;
  if (q < qmin || q > 255) return false;
  if (d > delta_top && aq_weight(d) == 0.f) return false;
  return true;
}
// sums over the bins [i0, i1) of count * bit length of the quantised level and count * squared
// error, for step q (histogram.cc:236-247; products wrap at 32 bits like the compiled reference's)
SJB_HD void aq_bin_sums(const int32_t* h, int i0, int i1, int q, long long* bits_sum, long long* dist_sum) {
  const int recip = ((1 << 16) + q - 1) / q;
  long long bs = 0, ds = 0;
  for (int i = i0; i < i1; ++i) {
    const int v = (i << kAqShift) + (1 << (kAqShift - 1));   // bin centre
    const int level = (v * recip + 32768) >> 16;
    const int e = v - level * q;
    bs += (int32_t)((uint32_t)h[i] * (uint32_t)bit_length((uint32_t)level));
    ds += (int32_t)((uint32_t)h[i] * (uint32_t)(e * e));
  }
  *bits_sum = bs;
  *dist_sum = ds;
}
// the weighted linear fits of one position over its candidates, in candidate order
SJB_HD void aq_fit_position(const long long* bits_sum, const long long* dist_sum, int q0, int qmin, int delta_top,
                            AqFit* out) {
  const double kCorrelation = 0.5;
  double sw = 0., sx = 0., sxx = 0., syy1 = 0., sy1 = 0., sxy1 = 0., sy2 = 0., sxy2 = 0.;
  for (int d = 0; d < kAqNumDelta; ++d) {
    if (!aq_candidate_used(q0, qmin, delta_top, d)) {
      out->dist[d] = SJB_AQ_FLT_MAX;
      out->rate[d] = 0;
      continue;
    }
    const double bsum = (double)bits_sum[d], dsum = (double)dist_sum[d];
    out->dist[d] = (float)dsum;
    out->rate[d] = (float)bsum;
    const double w = aq_weight(d);
    if (w > 0.) {
      const double x = (double)(d + kAqDeltaMin);
      sw += w;
      sx += w * x;
      sxx += w * x * x;
      sy1 += w * dsum;
      syy1 += w * dsum * dsum;
      sy2 += w * bsum;
      sxy1 += w * dsum * x;
      sxy2 += w * bsum * x;
    }
  }
  const double cov = sw * sxy1 - sx * sy1;
  out->skip = (cov * cov < kCorrelation * (sw * sxx - sx * sx) * (sw * syy1 - sy1 * sy1)) ? 1 : 0;
  out->cov = cov;
  out->den = sw * sxy2 - sx * sy2;
}
SJB_HD double aq_lambda(double num, double den) {
  double lambda = 128.;                    // fallback
  if (num > 1000. && den < -10.) {
    lambda = -num / den;
    if (lambda < 1.) lambda = 1.;
  }
  return lambda;
}
SJB_HD int aq_best_delta(const AqFit& f, int delta_top, double lambda) {
  float best = SJB_AQ_FLT_MAX;
  int best_delta = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int d = 0; d < kAqNumDelta; ++d) {            // fixed trip count: candidates stay in registers on the device
    if (d <= delta_top && f.dist[d] < SJB_AQ_FLT_MAX) {
      const float score = (float)((double)f.dist[d] + lambda * (double)f.rate[d]);
      if (score < best) {
        best = score;
        best_delta = d + kAqDeltaMin;
      }
    }
  }
  return best_delta;
}
// reciprocal / bias constants of one matrix entry (quantize.cc:116-148 in the fused form of the top
// of this file); returns false when the entry cannot be expressed (the caller refuses the encode)
SJB_HD bool aq_finalize_entry(uint32_t q, bool is_dc, int q_bias, int32_t* iq, int32_t* cpos_out) {
  const uint32_t recip = (q == 1) ? 0xffffu : (((1u << 16) + q / 2) / q) & 0xffffu;
  const uint32_t bias8 = (q == 1 || is_dc) ? 0x80u : (uint32_t)q_bias;
  const uint32_t bias = ((((bias8 * q) << 4) + 128) >> 8) & 0xffffu;
  const int thresh = (int)(((1u << 20) + recip - 1) / recip) - (int)bias;
  const long long cpos = (long long)bias * recip;
  *iq = (int32_t)recip;
  *cpos_out = (int32_t)cpos;
  return !(thresh < 0 || thresh > 0xffff || cpos + 17000LL * recip >= (1LL << 31));
}

// Position inside a tile of `count` consecutive blocks starting at global block `first` of the
// block that worker `i` walks: the tile's luma blocks in order, then its chroma blocks.  Blocks
// are stored MCU by MCU, `lb` luma blocks then `mb - lb` chroma blocks (6/4 for 4:2:0, 3/1 for
// 4:4:4; 4:0:0 has no chroma and keeps the identity).
SJB_HD uint32_t walk_order(uint32_t first, uint32_t count, uint32_t i, int mcu_blocks) {
  if (mcu_blocks == 1) return i;
  uint32_t g;
  if (mcu_blocks == 6) {
    const uint32_t l0 = (first / 6u) * 4u + sjb_minu(first % 6u, 4u);                     // luma blocks before the tile
    const uint32_t end = first + count;
    const uint32_t nl = (end / 6u) * 4u + sjb_minu(end % 6u, 4u) - l0;                    // luma blocks in the tile
    if (i < nl) { const uint32_t k = l0 + i; g = (k / 4u) * 6u + (k % 4u); }
    else { const uint32_t k = (first - l0) + (i - nl); g = (k / 2u) * 6u + 4u + (k % 2u); }
  } else {
    const uint32_t l0 = (first / 3u) + sjb_minu(first % 3u, 1u);
    const uint32_t end = first + count;
    const uint32_t nl = (end / 3u) + sjb_minu(end % 3u, 1u) - l0;
    if (i < nl) { g = (l0 + i) * 3u; }
    else { const uint32_t k = (first - l0) + (i - nl); g = (k / 2u) * 3u + 1u + (k % 2u); }
  }
  return g - first;
}

}  // namespace sjb
