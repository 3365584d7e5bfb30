/*
 * sjpeg_oracle.c -- TEST INFRASTRUCTURE ONLY (see sjpeg_oracle.h for the rules and the parity pin).
 *
 * CPU restatement, in plain C99, of the reference's scalar ("plain-C") code path for
 *   RGB -> YCbCr (+4:2:0) -> integer fDCT -> quantise (plain / trellis) -> run/level ->
 *   Huffman (default or optimised tables) -> bit-pack -> 0xFF stuffing, with the adaptive
 *   quantisation histogram analysis.
 * All citations are file:line inside /root/reference/src.
 */
#include "sjpeg_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Tables (ITU-T T.81 standard data; quantize.cc:32-75, entropy.cc:31-82)
 * ---------------------------------------------------------------------------------------- */
static const uint8_t kZZ[64] = { /* zig-zag index -> natural index, quantize.cc:32-41 */
  0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20,
  13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52,
  45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

static const uint8_t kK1[2][64] = { /* T.81 K.1, quantize.cc:57-75 */
  { 16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113,
    92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99 },
  { 17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
    47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99 } };

/* T.81 K.3 default Huffman tables, entropy.cc:31-82 */
static const uint8_t kStdDCBits[2][16] = {
  { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 },
  { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 } };
static const uint8_t kStdDCSyms[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
static const uint8_t kStdACBits[2][16] = {
  { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125 },
  { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119 } };
static const uint8_t kStdACSyms[2][162] = {
  { 0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07,
    0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0,
    0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49,
    0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7,
    0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5,
    0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa },
  { 0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71,
    0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0,
    0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5,
    0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa } };

/* sjpegi.h:165, 268-276 */
enum { FP_BITS = 16, AC_BITS = 4, HSHIFT = 2, HHALF = 1 << (HSHIFT - 1), HBINS = 1 << 7,
       HLAMBDA = 0x80, QDELTA_MIN = -12, QDELTA_MAX = 12, QSIZE = QDELTA_MAX + 1 - QDELTA_MIN };

/* sjpegi.h:186-198 : number of bits of v > 0 */
static int bitlen(int v) {
  int n = 0;
  while (v > 0) { ++n; v >>= 1; }
  return n;
}

/* ------------------------------------------------------------------------------------------
 * Quantiser tables
 * ---------------------------------------------------------------------------------------- */

/* quantize.cc:77-82 (GetQFactor) + :88-96 (SetQuantMatrix). float arithmetic on purpose. */
static void scale_matrix(const uint8_t in[64], float q_factor, uint8_t out[64]) {
  q_factor /= 100.f;
  for (int i = 0; i < 64; ++i) {
    const int v = (int)(in[i] * q_factor + .5f);
    out[i] = (uint8_t)((v < 1) ? 1 : (v > 255) ? 255 : v);
  }
}

void sjo_quality_to_matrices(float q, uint8_t out[2][64]) {
  q = (q <= 0) ? 5000 : (q < 50) ? 5000 / q : (q < 100) ? 2 * (100 - q) : 0;
  q = floorf(q);
  scale_matrix(kK1[0], q, out[0]);
  scale_matrix(kK1[1], q, out[1]);
}

/* quantize.cc:123-148 */
void sjo_finalize_quantizer(uint8_t quant[64], const uint8_t min_quant[64], int q_bias,
                            uint16_t iquant[64], uint16_t bias[64], uint16_t qthresh[64]) {
  for (int i = 0; i < 64; ++i) {
    if (quant[i] < min_quant[i]) quant[i] = min_quant[i];
  }
  for (int i = 0; i < 64; ++i) {
    const uint16_t v = quant[i];
    const uint16_t iq = (v == 1) ? 0xffffu : (uint16_t)(((1u << FP_BITS) + v / 2) / v);
    const uint16_t b = (v == 1) ? 0x80 : (i == 0) ? 0x80 : (uint16_t)q_bias;
    const uint16_t ib = (uint16_t)((((b * v) << AC_BITS) + 128) >> 8);
    const uint16_t th = (uint16_t)(((1 << (FP_BITS + AC_BITS)) + iq - 1) / iq - ib);
    iquant[i] = iq;
    bias[i] = ib;
    qthresh[i] = th;
  }
}

typedef struct {
  uint8_t quant[64];
  uint8_t min_quant[64];
  uint16_t iquant[64], qthresh[64], bias[64];
  const uint32_t* codes;   /* AC codes for the trellis rate (quantize.cc:150-152) */
} Quantizer;

static void finalize(Quantizer* q, int q_bias) {
  sjo_finalize_quantizer(q->quant, q->min_quant, q_bias, q->iquant, q->bias, q->qthresh);
}

#define QUANTIZE(A, M, B) (((((A) + (B)) * (M)) >> FP_BITS) >> AC_BITS)  /* quantize.cc:116-121 */

/* ------------------------------------------------------------------------------------------
 * Geometry
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int nb_comps, mcu_blocks, block_w, block_h;
  int nb_blocks[3], quant_idx[3], block_dims[3];
} Geom;

/* encoders.cc:32-88 */
static int init_geom(int yuv_mode, Geom* g) {
  memset(g, 0, sizeof(*g));
  if (yuv_mode == SJO_YUV_444) {
    g->nb_comps = 3; g->mcu_blocks = 3; g->block_w = g->block_h = 8;
    for (int c = 0; c < 3; ++c) { g->nb_blocks[c] = 1; g->quant_idx[c] = (c > 0); g->block_dims[c] = 0x11; }
  } else if (yuv_mode == SJO_YUV_420) {
    g->nb_comps = 3; g->mcu_blocks = 6; g->block_w = g->block_h = 16;
    g->nb_blocks[0] = 4; g->nb_blocks[1] = g->nb_blocks[2] = 1;
    g->quant_idx[0] = 0; g->quant_idx[1] = g->quant_idx[2] = 1;
    g->block_dims[0] = 0x22; g->block_dims[1] = g->block_dims[2] = 0x11;
  } else if (yuv_mode == SJO_YUV_400) {
    g->nb_comps = 1; g->mcu_blocks = 1; g->block_w = g->block_h = 8;
    g->nb_blocks[0] = 1; g->quant_idx[0] = 0; g->block_dims[0] = 0x11;
  } else {
    return 0;
  }
  return 1;
}

void sjo_geometry(int yuv_mode, int w, int h, int* mcu_w, int* mcu_h, int* mcu_blocks,
                  int* mb_w, int* mb_h) {
  Geom g;
  if (!init_geom(yuv_mode, &g)) { *mcu_w = *mcu_h = *mcu_blocks = *mb_w = *mb_h = 0; return; }
  *mcu_w = g.block_w; *mcu_h = g.block_h; *mcu_blocks = g.mcu_blocks;
  *mb_w = (w + g.block_w - 1) / g.block_w;   /* enc.cc:410-411 */
  *mb_h = (h + g.block_h - 1) / g.block_h;
}

/* component (0 luma, else chroma) -> quantiser index of block k inside an MCU */
static int qidx_of_block(const Geom* g, int k) {
  int c = 0, acc = 0;
  for (c = 0; c < g->nb_comps; ++c) {
    acc += g->nb_blocks[c];
    if (k < acc) break;
  }
  return g->quant_idx[c];
}
static int comp_of_block(const Geom* g, int k) {
  int c = 0, acc = 0;
  for (c = 0; c < g->nb_comps; ++c) {
    acc += g->nb_blocks[c];
    if (k < acc) break;
  }
  return c;
}

/* ------------------------------------------------------------------------------------------
 * Colour conversion (colors_rgb.cc:31-32, 785-879; BGRA/RGBA :881-1025 read the same triplet)
 * ---------------------------------------------------------------------------------------- */
enum { FRAC = 16, HALF = 1 << FRAC >> 1, ROUND_UV = HALF << 2, ROUND_Y = HALF - (128 << FRAC) };

static void load_rgb(const uint8_t* p, int fmt, int* r, int* g, int* b) {
  if (fmt == SJO_BGRA) { *r = p[2]; *g = p[1]; *b = p[0]; }
  else { *r = p[0]; *g = p[1]; *b = p[2]; }
}
static int16_t to_y(int r, int g, int b) {   /* colors_rgb.cc:785-795 */
  return (int16_t)((19595 * r + 38469 * g + 7471 * b + ROUND_Y) >> FRAC);
}

/* colors_rgb.cc:830-848 : 8x8 pixels -> Y[,U,V] blocks (4:4:4 and 4:0:0) */
static void get_8x8(const uint8_t* data, int step, int pstep, int fmt, int with_uv, int16_t* out) {
  for (int y = 0; y < 8; ++y) {
    for (int x = 0; x < 8; ++x) {
      int r, g, b;
      load_rgb(data + pstep * x, fmt, &r, &g, &b);
      out[8 * y + x] = to_y(r, g, b);
      if (with_uv) {   /* colors_rgb.cc:809-819 */
        out[64 + 8 * y + x] = (int16_t)((-11059 * r - 21709 * g + 32768 * b + HALF) >> FRAC);
        out[128 + 8 * y + x] = (int16_t)((32768 * r - 27439 * g - 5329 * b + HALF) >> FRAC);
      }
    }
    data += step;
  }
}

/* colors_rgb.cc:850-879 : 16x16 pixels -> Y00 Y01 Y10 Y11 U V */
static void get_16x16(const uint8_t* data, int step, int pstep, int fmt, int16_t* out) {
  for (int y = 0; y < 16; ++y) {
    for (int x = 0; x < 16; ++x) {
      int r, g, b;
      load_rgb(data + y * step + pstep * x, fmt, &r, &g, &b);
      out[64 * (2 * (y >> 3) + (x >> 3)) + 8 * (y & 7) + (x & 7)] = to_y(r, g, b);
    }
  }
  for (int y = 0; y < 8; ++y) {
    for (int x = 0; x < 8; ++x) {
      int sr = 0, sg = 0, sb = 0;
      for (int k = 0; k < 4; ++k) {
        int r, g, b;
        load_rgb(data + (2 * y + (k >> 1)) * step + pstep * (2 * x + (k & 1)), fmt, &r, &g, &b);
        sr += r; sg += g; sb += b;
      }
      /* ToU / ToV, colors_rgb.cc:797-806 (arithmetic shift = floor) */
      out[4 * 64 + 8 * y + x] = (int16_t)((-11059 * sr - 21709 * sg + 32768 * sb + ROUND_UV) >> (FRAC + 2));
      out[5 * 64 + 8 * y + x] = (int16_t)((32768 * sr - 27439 * sg - 5329 * sb + ROUND_UV) >> (FRAC + 2));
    }
  }
}

/* colors_rgb.cc:1212-1232 */
static void replicate(const uint8_t* src, int src_stride, uint8_t* dst, int dst_stride,
                      int sub_w, int sub_h, int w, int h, int pstep) {
  if (sub_w > w) sub_w = w;
  if (sub_h > h) sub_h = h;
  for (int y = 0; y < h; ++y) {
    const int sy = (y < sub_h) ? y : sub_h - 1;
    for (int x = 0; x < w; ++x) {
      const int sx = (x < sub_w) ? x : sub_w - 1;
      memcpy(dst + y * dst_stride + x * pstep, src + (ptrdiff_t)sy * src_stride + sx * pstep, pstep);
    }
  }
}

static int block_average(const int16_t* b) {   /* encoders.cc:95-99 */
  int dc = 0;
  for (int i = 0; i < 64; ++i) dc += b[i];
  return (dc + 32) >> 6;
}
static void block_fill(int dc, int16_t* b) {
  for (int i = 0; i < 64; ++i) b[i] = (int16_t)dc;
}
/* encoders.cc:107-125 */
static void average_extra_luma(int sub_w, int sub_h, int16_t* out) {
  int dc = block_average(out);
  if (sub_w <= 8) block_fill(dc, out + 64);
  if (sub_h <= 8) {
    if (sub_w > 8) dc = block_average(out + 64);
    block_fill(dc, out + 128);
    block_fill(dc, out + 192);
  } else if (sub_w <= 8) {
    dc = block_average(out + 128);
    block_fill(dc, out + 192);
  }
}

/* Encoder420/444/400::GetSamples, encoders.cc:170-182, 206-215, 239-248 */
static void get_samples(const uint8_t* pix, int W, int H, int stride, const Geom* g, int fmt,
                        int mb_x, int mb_y, int16_t* out) {
  const int pstep = (fmt == SJO_RGB) ? 3 : 4;
  const int bs = g->block_w;   /* 16 or 8 */
  const int clipped = (mb_x == W / bs) || (mb_y == H / bs);   /* enc.cc:279-292 */
  const uint8_t* src = pix + ((ptrdiff_t)pstep * mb_x + (ptrdiff_t)mb_y * stride) * bs;
  int step = stride;
  uint8_t tmp[16 * 16 * 4];
  if (clipped) {
    replicate(src, stride, tmp, pstep * bs, W - mb_x * bs, H - mb_y * bs, bs, bs, pstep);
    src = tmp;
    step = pstep * bs;
  }
  if (g->mcu_blocks == 6) {
    get_16x16(src, step, pstep, fmt, out);
    if (clipped) average_extra_luma(W - mb_x * 16, H - mb_y * 16, out);
  } else {
    get_8x8(src, step, pstep, fmt, g->mcu_blocks == 3, out);
  }
}

/* Planar / semi-planar sources (encoders.cc:256-507): samples are pixel - 128
 * (colors_rgb.cc:1234-1260 Convert8To16b[Clipped]), clipped MCUs replicate the last valid
 * row/column of each plane (GetReplicatedYSamples encoders.cc:138-143, Replicate8b with x_step 2
 * for NV12/NV21 :307-314).  uv_step = 1 for planar chroma, 2 for interleaved. */
typedef struct {
  const uint8_t* y; const uint8_t* u; const uint8_t* v;
  int ys, us, vs, uv_step;
} Planes;

static void plane_block(const uint8_t* p, int stride, int xstep, int pw, int ph, int x0, int y0, int16_t* out) {
  for (int y = 0; y < 8; ++y) {
    const int sy = (y0 + y < ph) ? y0 + y : ph - 1;
    for (int x = 0; x < 8; ++x) {
      const int sx = (x0 + x < pw) ? x0 + x : pw - 1;
      out[8 * y + x] = (int16_t)(p[(ptrdiff_t)sy * stride + sx * xstep] - 128);
    }
  }
}

static void get_samples_planar(const Planes* P, int W, int H, const Geom* g, int mb_x, int mb_y, int16_t* out) {
  if (g->mcu_blocks == 6) {
    const int clipped = (mb_x == W / 16) || (mb_y == H / 16);
    for (int k = 0; k < 4; ++k) plane_block(P->y, P->ys, 1, W, H, 16 * mb_x + 8 * (k & 1), 16 * mb_y + 8 * (k >> 1), out + 64 * k);
    if (clipped) average_extra_luma(W - mb_x * 16, H - mb_y * 16, out);
    plane_block(P->u, P->us, P->uv_step, (W + 1) >> 1, (H + 1) >> 1, 8 * mb_x, 8 * mb_y, out + 4 * 64);
    plane_block(P->v, P->vs, P->uv_step, (W + 1) >> 1, (H + 1) >> 1, 8 * mb_x, 8 * mb_y, out + 5 * 64);
  } else {
    plane_block(P->y, P->ys, 1, W, H, 8 * mb_x, 8 * mb_y, out);
    if (g->mcu_blocks == 3) {
      plane_block(P->u, P->us, 1, W, H, 8 * mb_x, 8 * mb_y, out + 64);
      plane_block(P->v, P->vs, 1, W, H, 8 * mb_x, 8 * mb_y, out + 128);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Integer fDCT (fdct.cc:28-43, 67-144, 150-209, 596-609).  Output = 16 x JPEG-normalised DCT.
 * ---------------------------------------------------------------------------------------- */
static const int16_t kT04[7] = { 22725, 21407, 19266, 16384, 12873, 8867, 4520 };
static const int16_t kT17[7] = { 31521, 29692, 26722, 22725, 17855, 12299, 6270 };
static const int16_t kT26[7] = { 29692, 27969, 25172, 21407, 16819, 11585, 5906 };
static const int16_t kT35[7] = { 26722, 25172, 22654, 19266, 15137, 10426, 5315 };

#define MUL16(a, c) (((a) * (c)) >> 16)   /* fdct.cc:153 */
/* {a,b} <- {a-b, a+b}  fdct.cc:45-50 */
#define BFLY(a, b) do { (a) -= (b); (b) += (b); (b) += (a); } while (0)

/* one column (stride 8), fdct.cc:67-144 with the C macros of :150-157 */
static void column_dct(int16_t* in) {
  int32_t m0 = in[0 * 8], m2 = in[2 * 8], m7 = in[7 * 8], m5 = in[5 * 8];
  int32_t m1, m3, m4, m6;
  BFLY(m0, m7);
  BFLY(m2, m5);
  m3 = in[3 * 8]; m4 = in[4 * 8];
  BFLY(m3, m4);
  m6 = in[6 * 8]; m1 = in[1 * 8];
  BFLY(m1, m6);
  BFLY(m7, m4);
  BFLY(m6, m5);
  m4 <<= 3; m5 <<= 3;
  BFLY(m4, m5);
  in[0 * 8] = (int16_t)m5;
  in[4 * 8] = (int16_t)m4;
  m7 <<= 3; m6 <<= 3; m3 <<= 3; m0 <<= 3;
  m4 = 27146; m5 = m4;                 /* kTan2 */
  m4 = MUL16(m4, m7);
  m5 = MUL16(m5, m6);
  m4 -= m6;
  m5 += m7;
  in[2 * 8] = (int16_t)m5;
  in[6 * 8] = (int16_t)m4;
  m6 = 23170;                          /* k2Sqrt2 */
  m2 <<= 4; m1 <<= 4;
  BFLY(m1, m2);
  m2 = MUL16(m2, m6);
  m1 = MUL16(m1, m6);
  BFLY(m3, m1);
  BFLY(m0, m2);
  m4 = -21746;                         /* kTan3m1 */
  m5 = 13036;                          /* kTan1 */
  m7 = m3;
  m6 = m1;
  m3 = MUL16(m3, m4);
  m1 = MUL16(m1, m5);
  m3 += m7;
  m1 += m2;
  m1 += 1;                             /* CORRECT_LSB */
  m3 += 1;
  m4 = MUL16(m4, m0);
  m5 = MUL16(m5, m2);
  m4 += m0;
  m0 -= m3;
  m7 += m4;
  m5 -= m6;
  in[1 * 8] = (int16_t)m1;
  in[3 * 8] = (int16_t)m0;
  in[5 * 8] = (int16_t)m7;
  in[7 * 8] = (int16_t)m5;
}

/* fdct.cc:174-209 */
static void row_dct(int16_t* in, const int16_t* t) {
  const int a0 = in[0] + in[7], b0 = in[0] - in[7];
  const int a1 = in[1] + in[6], b1 = in[1] - in[6];
  const int a2 = in[2] + in[5], b2 = in[2] - in[5];
  const int a3 = in[3] + in[4], b3 = in[3] - in[4];
  const int C1 = t[0], C2 = t[1], C3 = t[2], C4 = t[3], C5 = t[4], C6 = t[5], C7 = t[6];
  const int c0 = a0 + a3, c1 = a0 - a3, c2 = a1 + a2, c3 = a1 - a2;
  in[0] = (int16_t)((C4 * (c0 + c2)) >> 16);
  in[4] = (int16_t)((C4 * (c0 - c2)) >> 16);
  in[2] = (int16_t)((C2 * c1 + C6 * c3) >> 16);
  in[6] = (int16_t)((C6 * c1 - C2 * c3) >> 16);
  in[1] = (int16_t)((C1 * b0 + C3 * b1 + C5 * b2 + C7 * b3) >> 16);
  in[3] = (int16_t)((C3 * b0 - C7 * b1 - C1 * b2 - C5 * b3) >> 16);
  in[5] = (int16_t)((C5 * b0 - C1 * b1 + C7 * b2 + C3 * b3) >> 16);
  in[7] = (int16_t)((C7 * b0 - C5 * b1 + C3 * b2 - C1 * b3) >> 16);
}

void sjo_fdct(int16_t* c, int nb) {   /* fdct.cc:596-609 */
  for (; nb > 0; --nb, c += 64) {
    for (int i = 0; i < 8; ++i) column_dct(c + i);
    row_dct(c + 0 * 8, kT04); row_dct(c + 1 * 8, kT17); row_dct(c + 2 * 8, kT26);
    row_dct(c + 3 * 8, kT35); row_dct(c + 4 * 8, kT04); row_dct(c + 5 * 8, kT35);
    row_dct(c + 6 * 8, kT26); row_dct(c + 7 * 8, kT17);
  }
}

static void image_to_blocks_ex(const uint8_t* pix, const Planes* planes, int w, int h, int stride, int yuv_mode,
                               int fmt, int do_dct, int16_t* out) {
  Geom g;
  if (!init_geom(yuv_mode, &g)) return;
  const int mb_w = (w + g.block_w - 1) / g.block_w, mb_h = (h + g.block_h - 1) / g.block_h;
  for (int mb_y = 0; mb_y < mb_h; ++mb_y) {     /* enc.cc:257-271 raster order */
    for (int mb_x = 0; mb_x < mb_w; ++mb_x) {
      if (planes) get_samples_planar(planes, w, h, &g, mb_x, mb_y, out);
      else get_samples(pix, w, h, stride, &g, fmt, mb_x, mb_y, out);
      if (do_dct) sjo_fdct(out, g.mcu_blocks);
      out += 64 * g.mcu_blocks;
    }
  }
}
static void image_to_blocks(const uint8_t* pix, int w, int h, int stride, int yuv_mode, int fmt,
                            int do_dct, int16_t* out) {
  image_to_blocks_ex(pix, NULL, w, h, stride, yuv_mode, fmt, do_dct, out);
}
void sjo_image_to_coeffs(const uint8_t* pix, int w, int h, int stride, int yuv_mode, int fmt,
                         int16_t* out) {
  image_to_blocks(pix, w, h, stride, yuv_mode, fmt, 1, out);
}
void sjo_image_to_samples(const uint8_t* pix, int w, int h, int stride, int yuv_mode, int fmt,
                          int16_t* out) {
  image_to_blocks(pix, w, h, stride, yuv_mode, fmt, 0, out);
}

/* ------------------------------------------------------------------------------------------
 * Quantisation -> run/levels
 * ---------------------------------------------------------------------------------------- */
typedef struct { int16_t run; uint16_t level; } RunLevel;                     /* sjpegi.h:239 */
typedef struct { int16_t last, nb_coeffs; uint16_t dc_code; int8_t idx; } DCTCoeffs;  /* :245 */

/* quantize.cc:288-320 */
static int quantize_block(const int16_t in[64], int idx, const Quantizer* Q, DCTCoeffs* out,
                          RunLevel* rl) {
  int prev = 1, nb = 0;
  for (int i = 1; i < 64; ++i) {
    const int j = kZZ[i];
    int v = in[j];
    const int32_t mask = v >> 31;
    v = (v ^ mask) - mask;
    if (v >= Q->qthresh[j]) {
      v = QUANTIZE(v, Q->iquant[j], Q->bias[j]);
      const int n = bitlen(v);
      const uint16_t code = (uint16_t)((v ^ mask) & ((1 << n) - 1));
      rl[nb].level = (uint16_t)((code << 4) | n);
      rl[nb].run = (int16_t)(i - prev);
      prev = i + 1;
      ++nb;
    }
  }
  const int dc = (in[0] < 0) ? -QUANTIZE(-in[0], Q->iquant[0], Q->bias[0])
                             : QUANTIZE(in[0], Q->iquant[0], Q->bias[0]);
  out->idx = (int8_t)idx;
  out->last = (int16_t)(prev - 1);
  out->nb_coeffs = (int16_t)nb;
  return dc;
}

/* quantize.cc:325-457.  score_t is uint32 and wraps. */
typedef struct {
  uint32_t code; int nbits; uint32_t score, disto, bits, run; int best_prev; int pos, rank;
} TNode;

static int search_best_prev(TNode* nodes, int cur_idx, const uint32_t disto0[],
                            const uint32_t codes[], uint32_t lambda) {   /* :350-383 */
  TNode* node = &nodes[cur_idx];
  int found = 0;
  const uint32_t base_disto = node->disto + disto0[node->pos - 1];
  for (int c = cur_idx - 1; c >= 0; --c) {
    const TNode* cur = &nodes[c];
    const int run = node->pos - 1 - cur->pos;
    if (run < 0) continue;
    uint32_t bits = (uint32_t)node->nbits;
    bits += (uint32_t)(run >> 4) * (codes[0xf0] & 0xff);
    const uint32_t disto = base_disto - disto0[cur->pos];
    if (disto + lambda * bits >= node->score) break;
    const uint32_t sym = (uint32_t)(((run & 15) << 4) | node->nbits);
    bits += codes[sym] & 0xff;
    const uint32_t score = disto + lambda * bits + cur->score;
    if (score < node->score) {
      node->score = score; node->disto = disto; node->bits = bits;
      node->best_prev = c; node->rank = cur->rank + 1; node->run = (uint32_t)run;
      found = 1;
    }
  }
  return found;
}

static int trellis_quantize_block(const int16_t in[64], int idx, const Quantizer* Q,
                                  DCTCoeffs* out, RunLevel* rl) {   /* :388-457 */
  TNode nodes[1 + 2 * 63];
  uint32_t disto0[64];
  memset(&nodes[0], 0, sizeof(nodes[0]));   /* InitSink :340-347 */
  nodes[0].best_prev = -1;
  int cur = 1;
  disto0[0] = 0;
  for (int i = 1; i < 64; ++i) {
    const int j = kZZ[i];
    const uint32_t q = (uint32_t)Q->quant[j] << AC_BITS;
    const uint32_t lambda = q * q / 32u;
    int V = in[j];
    const int32_t mask = V >> 31;
    V = (V ^ mask) - mask;
    disto0[i] = (uint32_t)(V * V) + disto0[i - 1];
    int v = QUANTIZE(V, Q->iquant[j], Q->bias[j]);
    if (v == 0) continue;
    int nbits = bitlen(v);
    for (int k = 0; k < 2; ++k) {
      const int err = V - v * (int)q;
      nodes[cur].code = (uint32_t)((v ^ mask) & ((1 << nbits) - 1));
      nodes[cur].pos = i;
      nodes[cur].disto = (uint32_t)(err * err);
      nodes[cur].nbits = nbits;
      nodes[cur].score = 0xffffffffu;
      nodes[cur].best_prev = -1;
      if (search_best_prev(nodes, cur, disto0, Q->codes, lambda)) ++cur;
      --nbits;
      if (nbits <= 0) break;
      v = (1 << nbits) - 1;
    }
  }
  int nz = 0;
  if (cur != 1) {
    uint32_t best = 0xffffffffu;
    for (int c = cur - 1; c >= 0; --c) {     /* "while (cur_node-- != &nodes[0])": sink included */
      const uint32_t d = disto0[63] - disto0[nodes[c].pos];
      nodes[c].disto += d;
      nodes[c].score += d;
      if (nodes[c].score < best) { nz = c; best = nodes[c].score; }
    }
  }
  int nb = nodes[nz].rank;
  out->idx = (int8_t)idx;
  out->last = (int16_t)nodes[nz].pos;
  out->nb_coeffs = (int16_t)nb;
  while (nb-- > 0) {
    rl[nb].level = (uint16_t)((nodes[nz].code << 4) | (uint32_t)nodes[nz].nbits);
    rl[nb].run = (int16_t)nodes[nz].run;
    nz = nodes[nz].best_prev;
  }
  return (in[0] < 0) ? -QUANTIZE(-in[0], Q->iquant[0], Q->bias[0])
                     : QUANTIZE(in[0], Q->iquant[0], Q->bias[0]);
}

/* run/levels -> signed quantised coefficients in zig-zag order (inverse of what the block coder
 * consumes; used for the stage-level parity dumps) */
static void runlevels_to_zigzag(int dc, const DCTCoeffs* c, const RunLevel* rl, int16_t zz[64]) {
  memset(zz, 0, 64 * sizeof(zz[0]));
  zz[0] = (int16_t)dc;
  int pos = 1;
  for (int i = 0; i < c->nb_coeffs; ++i) {
    pos += rl[i].run;
    const int n = rl[i].level & 15, amp = rl[i].level >> 4;
    /* amplitude bits: positive as-is, negative as one's complement (quantize.cc:302-303) */
    const int v = (amp >> (n - 1)) ? amp : amp - ((1 << n) - 1);
    zz[pos++] = (int16_t)v;
  }
}

/* ------------------------------------------------------------------------------------------
 * Huffman codes
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint8_t bits[16]; uint8_t syms[256]; int nb_syms; } HTable;

/* entropy.cc:98-112 : tab[sym] = code<<16 | len */
static void build_codes(const HTable* h, uint32_t* tab) {
  uint32_t code = 0;
  const uint8_t* s = h->syms;
  for (int nb_bits = 1; nb_bits <= 16; ++nb_bits, code <<= 1) {
    for (int n = h->bits[nb_bits - 1]; n > 0; --n) {
      tab[*s++] = (code << 16) | (uint32_t)nb_bits;
      ++code;
    }
  }
}

static void std_table(int is_ac, int c, HTable* h) {
  memset(h, 0, sizeof(*h));
  if (is_ac) { memcpy(h->bits, kStdACBits[c], 16); memcpy(h->syms, kStdACSyms[c], 162); h->nb_syms = 162; }
  else       { memcpy(h->bits, kStdDCBits[c], 16); memcpy(h->syms, kStdDCSyms, 12); h->nb_syms = 12; }
}

/* entropy.cc:133-150 */
static uint16_t dc_diff_code(int dc, int* pred) {
  const int diff = dc - *pred;
  *pred = dc;
  if (diff == 0) return 0;
  int n, suff;
  if (diff < 0) { n = bitlen(-diff); suff = (diff - 1) & ((1 << n) - 1); }
  else          { n = bitlen(diff);  suff = diff; }
  return (uint16_t)(n | (suff << 4));
}

static int cmp_desc(const void* pa, const void* pb) {   /* entropy.cc:247-252 */
  const uint64_t a = *(const uint64_t*)pa, b = *(const uint64_t*)pb;
  return (a < b) ? 1 : -1;
}

/* entropy.cc:254-430 */
int sjo_build_optimal_table(const uint32_t* freq, int size, uint8_t out_bits[16], uint8_t* syms) {
  enum { MAX_BITS = 32, MAX_CODE_SIZE = 16 };
  int codesizes[257], chain[257], chain_end[257];   /* chain_end holds an index, not a pointer */
  uint64_t sorted[257];
  int nb_syms = 0;
  for (int i = 0; i < size; ++i) {
    const uint64_t v = freq[i];
    if (v > 0) sorted[nb_syms++] = (v << 9) | (uint64_t)i;
    codesizes[i] = 0; chain[i] = -1; chain_end[i] = i;
  }
  const int real_syms = nb_syms;
  qsort(sorted, (size_t)nb_syms, sizeof(sorted[0]), cmp_desc);
  sorted[nb_syms++] = (1ULL << 9) | (uint64_t)size;   /* the reserved all-ones code */
  codesizes[size] = 0; chain[size] = -1; chain_end[size] = size;

  int nb = nb_syms;
  while (nb-- > 1) {
    const uint64_t s1 = sorted[nb - 1], s2 = sorted[nb];
    int i = (int)(s1 & 0x1ff);
    const int j = (int)(s2 & 0x1ff);
    chain[chain_end[i]] = j;            /* *chain_end[i] = j */
    chain_end[i] = chain_end[j];
    do { ++codesizes[i]; i = chain[i]; } while (i >= 0);
    const uint64_t merged = s1 + (s2 & ~(uint64_t)0x1ff);
    int k = nb - 1;
    while (k > 0 && sorted[k - 1] < merged) { sorted[k] = sorted[k - 1]; --k; }
    sorted[k] = merged;
  }

  uint8_t bits[MAX_BITS];
  memset(bits, 0, sizeof(bits));
  int max_bit_size = 0;
  for (int i = 0; i <= size; ++i) {
    int s = codesizes[i];
    if (s > 0) {
      if (s > MAX_BITS) { s = MAX_BITS; codesizes[i] = MAX_BITS; }
      ++bits[s - 1];
      if (s > max_bit_size) max_bit_size = s;
    }
  }
  int start[MAX_BITS], position = 0;
  for (int i = 0; i < max_bit_size; ++i) { start[i] = position; position += bits[i]; }
  for (int symbol = 0; symbol < size; ++symbol) {
    const int s = codesizes[symbol];
    if (s > 0) syms[start[s - 1]++] = (uint8_t)symbol;
  }
  for (int l = max_bit_size - 1; l >= MAX_CODE_SIZE; --l) {
    while (bits[l] > 0) {
      int k = l - 2;
      while (bits[k] == 0) --k;
      bits[l] -= 2; bits[l - 1] += 1; bits[k] -= 1; bits[k + 1] += 2;
    }
  }
  max_bit_size = MAX_CODE_SIZE;
  while (bits[--max_bit_size] == 0) { }
  --bits[max_bit_size];
  for (int i = 0; i < MAX_CODE_SIZE; ++i) out_bits[i] = bits[i];
  return real_syms;
}

/* ------------------------------------------------------------------------------------------
 * Bit writer (bit_writer.h:172-209, bit_writer.cc:107-116): MSB first, 0xFF -> 0xFF 0x00
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint8_t* buf; size_t pos, cap; uint64_t acc; int nb; int ok; int raw; uint64_t nbits; } BW;

static void bw_need(BW* w, size_t extra) {
  if (w->pos + extra <= w->cap) return;
  size_t ncap = w->cap ? w->cap * 2 : 4096;
  while (ncap < w->pos + extra) ncap *= 2;
  uint8_t* nb = (uint8_t*)realloc(w->buf, ncap);
  if (nb == NULL) { w->ok = 0; return; }
  w->buf = nb; w->cap = ncap;
}
static void bw_byte(BW* w, int v) { bw_need(w, 1); if (w->ok) w->buf[w->pos++] = (uint8_t)v; }
static void bw_bytes(BW* w, const uint8_t* p, size_t n) {
  bw_need(w, n); if (w->ok) { memcpy(w->buf + w->pos, p, n); w->pos += n; }
}
static void bw_flush_bits(BW* w) {
  while (w->nb >= 8) {
    const int b = (int)(w->acc >> 56);
    bw_byte(w, b);
    if (b == 0xff && !w->raw) bw_byte(w, 0x00);
    w->acc <<= 8; w->nb -= 8;
  }
}
static void bw_put(BW* w, uint32_t bits, int nb) {
  w->nbits += (uint64_t)nb;
  if (w->nb + nb > 56) bw_flush_bits(w);
  w->nb += nb;
  w->acc |= (uint64_t)bits << (64 - w->nb);
}
static void bw_put_code(BW* w, uint32_t packed) { bw_put(w, packed >> 16, (int)(packed & 0xff)); }
static void bw_pad_flush(BW* w) {
  bw_flush_bits(w);
  const int pad = (-w->nb) & 7;
  if (pad) bw_put(w, (1u << pad) - 1, pad);
  bw_flush_bits(w);
}

/* ------------------------------------------------------------------------------------------
 * Encoder state + headers
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  Geom g;
  int W, H, mb_w, mb_h, yuv_mode;
  Quantizer quants[2];
  HTable huff[4];                  /* dc0 dc1 ac0 ac1 : index type*2 + c (entropy.cc:121) */
  uint32_t dc_codes[2][12], ac_codes[2][256];
  uint32_t freq_ac[2][257], freq_dc[2][13];
  int DCs[3];
  BW bw;
} Enc;

static void init_codes(Enc* e, int only_ac) {   /* entropy.cc:116-128 */
  const int nt = (e->g.nb_comps == 1) ? 1 : 2;
  for (int c = 0; c < nt; ++c) {
    if (!only_ac) build_codes(&e->huff[c], e->dc_codes[c]);
    build_codes(&e->huff[2 + c], e->ac_codes[c]);
  }
}

static void write_app0(Enc* e) {   /* headers.cc:48-61 */
  static const uint8_t k[] = { 0xff, 0xd8, 0xff, 0xe0, 0x00, 0x10, 0x4a, 0x46, 0x49, 0x46, 0x00,
                               0x01, 0x01, 0x00, 0x00, 0x01, 0x00, 0x01, 0x00, 0x00 };
  bw_bytes(&e->bw, k, sizeof(k));
}
static void write_dqt(Enc* e) {   /* headers.cc:182-196 */
  const int n = (e->yuv_mode == SJO_YUV_400) ? 1 : 2;
  const int size = n * 65 + 2;
  bw_byte(&e->bw, 0xff); bw_byte(&e->bw, 0xdb); bw_byte(&e->bw, 0x00); bw_byte(&e->bw, size);
  for (int m = 0; m < n; ++m) {
    bw_byte(&e->bw, m);
    for (int i = 0; i < 64; ++i) bw_byte(&e->bw, e->quants[m].quant[kZZ[i]]);
  }
}
static void write_sof(Enc* e) {   /* headers.cc:202-219 */
  const int size = 3 * e->g.nb_comps + 8;
  const uint8_t k[] = { 0xff, 0xc0, (uint8_t)(size >> 8), (uint8_t)size, 0x08,
                        (uint8_t)(e->H >> 8), (uint8_t)e->H, (uint8_t)(e->W >> 8), (uint8_t)e->W,
                        (uint8_t)e->g.nb_comps };
  bw_bytes(&e->bw, k, sizeof(k));
  for (int c = 0; c < e->g.nb_comps; ++c) {
    bw_byte(&e->bw, c + 1); bw_byte(&e->bw, e->g.block_dims[c]); bw_byte(&e->bw, e->g.quant_idx[c]);
  }
}
static void write_dht(Enc* e) {   /* headers.cc:221-238 */
  init_codes(e, 0);
  const int nt = (e->g.nb_comps == 1) ? 1 : 2;
  for (int c = 0; c < nt; ++c) {
    for (int type = 0; type <= 1; ++type) {
      const HTable* h = &e->huff[type * 2 + c];
      const int size = 3 + 16 + h->nb_syms;
      bw_byte(&e->bw, 0xff); bw_byte(&e->bw, 0xc4); bw_byte(&e->bw, size >> 8); bw_byte(&e->bw, size);
      bw_byte(&e->bw, (type << 4) | c);
      bw_bytes(&e->bw, h->bits, 16);
      bw_bytes(&e->bw, h->syms, (size_t)h->nb_syms);
    }
  }
}
static void write_sos(Enc* e) {   /* headers.cc:242-258 */
  const int size = 3 + e->g.nb_comps * 2 + 3;
  bw_byte(&e->bw, 0xff); bw_byte(&e->bw, 0xda); bw_byte(&e->bw, size >> 8); bw_byte(&e->bw, size);
  bw_byte(&e->bw, e->g.nb_comps);
  for (int c = 0; c < e->g.nb_comps; ++c) { bw_byte(&e->bw, c + 1); bw_byte(&e->bw, e->g.quant_idx[c] * 0x11); }
  bw_byte(&e->bw, 0x00); bw_byte(&e->bw, 0x3f); bw_byte(&e->bw, 0x00);
}

/* entropy.cc:161-198 */
static void code_block(Enc* e, const DCTCoeffs* c, const RunLevel* rl) {
  const int q = e->g.quant_idx[c->idx];
  const int dc_len = c->dc_code & 0x0f;
  bw_put_code(&e->bw, e->dc_codes[q][dc_len]);
  if (dc_len > 0) bw_put(&e->bw, c->dc_code >> 4, dc_len);
  const uint32_t* codes = e->ac_codes[q];
  for (int i = 0; i < c->nb_coeffs; ++i) {
    int run = rl[i].run;
    while (run & ~15) { bw_put_code(&e->bw, codes[0xf0]); run -= 16; }
    const uint32_t suffix = rl[i].level;
    const int n = (int)(suffix & 0x0f);
    bw_put_code(&e->bw, codes[(run << 4) | n]);
    bw_put(&e->bw, suffix >> 4, n);
  }
  if (c->last < 63) bw_put_code(&e->bw, codes[0x00]);
}

/* entropy.cc:208-227 */
static void add_stats(Enc* e, const DCTCoeffs* c, const RunLevel* rl) {
  const int q = e->g.quant_idx[c->idx];
  for (int i = 0; i < c->nb_coeffs; ++i) {
    const int run = rl[i].run;
    if (run >> 4) e->freq_ac[q][0xf0] += (uint32_t)(run >> 4);
    ++e->freq_ac[q][((run & 0x0f) << 4) | (rl[i].level & 0x0f)];
  }
  if (c->last < 63) ++e->freq_ac[q][0x00];
  ++e->freq_dc[q][c->dc_code & 0x0f];
}

/* entropy.cc:432-444 */
static void compile_stats(Enc* e) {
  const int nt = (e->g.nb_comps == 1) ? 1 : 2;
  for (int q = 0; q < nt; ++q) {
    memset(&e->huff[q], 0, sizeof(HTable));
    e->huff[q].nb_syms = sjo_build_optimal_table(e->freq_dc[q], 12, e->huff[q].bits, e->huff[q].syms);
    memset(&e->huff[2 + q], 0, sizeof(HTable));
    e->huff[2 + q].nb_syms =
        sjo_build_optimal_table(e->freq_ac[q], 256, e->huff[2 + q].bits, e->huff[2 + q].syms);
  }
}

/* ------------------------------------------------------------------------------------------
 * Adaptive quantisation (histogram.cc)
 * ---------------------------------------------------------------------------------------- */
void sjo_collect_histograms(const int16_t* coeffs, int nb_mcus, int yuv_mode, int32_t* counts) {
  Geom g;
  if (!init_geom(yuv_mode, &g)) return;
  memset(counts, 0, sizeof(int32_t) * 2 * 64 * (HBINS + 1));
  for (int m = 0; m < nb_mcus; ++m) {
    for (int k = 0; k < g.mcu_blocks; ++k, coeffs += 64) {
      int32_t* h = counts + qidx_of_block(&g, k) * 64 * (HBINS + 1);
      for (int i = 0; i < 64; ++i) {   /* histogram.cc:99-108 */
        const int a = (coeffs[i] < 0 ? -coeffs[i] : coeffs[i]) >> HSHIFT;
        if (a < HBINS) ++h[i * (HBINS + 1) + a];
      }
    }
  }
}

static const float kHistoWeight[QSIZE] = {   /* histogram.cc:119-124 */
  0, 0, 0, 0, 0, 1, 5, 16, 43, 94, 164, 228, 255, 228, 164, 94, 43, 16, 5, 1, 0, 0, 0, 0, 0 };

/* int multiply with the two's-complement wrap the compiled reference exhibits */
static int wrap_mul(int a, int b) { return (int)((uint32_t)a * (uint32_t)b); }

/* histogram.cc:126-315.  Types (int / float / double) and evaluation order kept as is. */
void sjo_analyse_histo(const int32_t* counts, int nb_comps, uint8_t quant[2][64],
                       const uint8_t min_quant[2][64], int qdelta_max_luma, int qdelta_max_chroma) {
  const double r_limit = 0.5;             /* kCorrelationThreshold */
  for (int c = (nb_comps > 1 ? 1 : 0); c >= 0; --c) {
    const int idx = c;                    /* quant_idx_[c] for c in {0,1} */
    const int32_t* histo = counts + idx * 64 * (HBINS + 1);
    const int delta_max = ((idx == 0) ? qdelta_max_luma : qdelta_max_chroma) - QDELTA_MIN;
    float sizes[64][QSIZE], distortions[64][QSIZE];
    double num = 0., den = 0.;
    uint64_t omit = 0x103ULL;             /* kOmittedChannels */
    for (int pos = 0; pos < 64; ++pos) {
      if (omit & (1ULL << pos)) continue;
      const int dq0 = quant[idx][pos];
      const int min_dq0 = min_quant[idx][pos];
      const int bias = 1 << FP_BITS >> 1;
      const int32_t* h = histo + pos * (HBINS + 1);
      int total = 0, last = 0;
      for (int i = 0; i < HBINS; ++i) { total += h[i]; if (h[i]) last = i + 1; }
      if (total < 0.5 * last) { omit |= 1ULL << pos; continue; }   /* kDensityThreshold */
      double sw = 0., sx = 0., sxx = 0., syy1 = 0., sy1 = 0., sxy1 = 0., sy2 = 0., sxy2 = 0.;
      for (int delta = 0; delta < QSIZE; ++delta) {
        double bsum = 0., dsum = 0.;
        const int dq = dq0 + (delta + QDELTA_MIN);
        if (dq >= min_dq0 && dq <= 255) {
          const int idq = ((1 << FP_BITS) + dq - 1) / dq;
          for (int i = 0; i < last; ++i) {
            if (h[i]) {
              const int v = (i << HSHIFT) + HHALF;
              const int qv = (v * idq + bias) >> FP_BITS;
              if (qv) {
                const int bits = bitlen(qv);
                const int dqv = qv * dq;
                const int error = (v - dqv) * (v - dqv);
                bsum += wrap_mul(h[i], bits);
                dsum += wrap_mul(h[i], error);
              } else {
                dsum += wrap_mul(wrap_mul(h[i], v), v);
              }
            }
          }
          distortions[pos][delta] = (float)dsum;
          sizes[pos][delta] = (float)bsum;
          const double w = kHistoWeight[delta];
          if (w > 0.) {
            const double x = (double)(delta + QDELTA_MIN);
            sw += w; sx += w * x; sxx += w * x * x;
            sy1 += w * dsum; syy1 += w * dsum * dsum;
            sy2 += w * bsum; sxy1 += w * dsum * x; sxy2 += w * bsum * x;
          }
        } else {
          distortions[pos][delta] = FLT_MAX;
          sizes[pos][delta] = 0;
        }
      }
      const double cov_xy1 = sw * sxy1 - sx * sy1;
      if (cov_xy1 * cov_xy1 < r_limit * (sw * sxx - sx * sx) * (sw * syy1 - sy1 * sy1)) {
        omit |= 1ULL << pos;
        continue;
      }
      num += cov_xy1;
      den += sw * sxy2 - sx * sy2;
    }
    double lambda = HLAMBDA;
    if (num > 1000. && den < -10.) {
      lambda = -num / den;
      if (lambda < 1.) lambda = 1.;
    }
    for (int pos = 0; pos < 64; ++pos) {
      if (omit & (1ULL << pos)) continue;
      float best_score = FLT_MAX;
      int best_dq = 0;
      for (int delta = 0; delta <= delta_max; ++delta) {
        if (distortions[pos][delta] < FLT_MAX) {
          const float score = (float)(distortions[pos][delta] + lambda * sizes[pos][delta]);
          if (score < best_score) { best_score = score; best_dq = delta + QDELTA_MIN; }
        }
      }
      quant[idx][pos] = (uint8_t)(quant[idx][pos] + best_dq);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Stage-level dumps
 * ---------------------------------------------------------------------------------------- */
static void quantize_image(const int16_t* coeffs, int nb_mcus, int yuv_mode,
                           const uint8_t quant[2][64], const uint8_t min_quant[2][64], int q_bias,
                           int trellis, int16_t* out) {
  Geom g;
  if (!init_geom(yuv_mode, &g)) return;
  Quantizer Q[2];
  uint32_t ac_codes[2][256];
  for (int i = 0; i < 2; ++i) {
    memcpy(Q[i].quant, quant[i], 64);
    memcpy(Q[i].min_quant, min_quant[i], 64);
    finalize(&Q[i], q_bias);
    HTable h;
    std_table(1, i, &h);
    memset(ac_codes[i], 0, sizeof(ac_codes[i]));
    build_codes(&h, ac_codes[i]);
    Q[i].codes = ac_codes[i];
  }
  RunLevel rl[64];
  DCTCoeffs dc;
  for (int m = 0; m < nb_mcus; ++m) {
    for (int k = 0; k < g.mcu_blocks; ++k, coeffs += 64, out += 64) {
      const Quantizer* q = &Q[qidx_of_block(&g, k)];
      const int d = trellis ? trellis_quantize_block(coeffs, 0, q, &dc, rl)
                            : quantize_block(coeffs, 0, q, &dc, rl);
      runlevels_to_zigzag(d, &dc, rl, out);
    }
  }
}
void sjo_quantize_image(const int16_t* coeffs, int nb_mcus, int yuv_mode,
                        const uint8_t quant[2][64], const uint8_t min_quant[2][64], int q_bias,
                        int16_t* out) {
  quantize_image(coeffs, nb_mcus, yuv_mode, quant, min_quant, q_bias, 0, out);
}
void sjo_trellis_quantize_image(const int16_t* coeffs, int nb_mcus, int yuv_mode,
                                const uint8_t quant[2][64], const uint8_t min_quant[2][64],
                                int q_bias, int16_t* out) {
  quantize_image(coeffs, nb_mcus, yuv_mode, quant, min_quant, q_bias, 1, out);
}

/* zig-zag quantised block -> run/levels as quantize.cc:288-320 would have emitted them */
static int zigzag_to_runlevels(const int16_t zz[64], DCTCoeffs* c, RunLevel* rl) {
  int prev = 1, nb = 0;
  for (int i = 1; i < 64; ++i) {
    int v = zz[i];
    if (v == 0) continue;
    const int32_t mask = v >> 31;
    v = (v ^ mask) - mask;
    const int n = bitlen(v);
    rl[nb].level = (uint16_t)((((v ^ mask) & ((1 << n) - 1)) << 4) | n);
    rl[nb].run = (int16_t)(i - prev);
    prev = i + 1;
    ++nb;
  }
  c->last = (int16_t)(prev - 1);
  c->nb_coeffs = (int16_t)nb;
  return zz[0];
}

void sjo_symbol_stats(const int16_t* zz, int nb_mcus, int yuv_mode, uint32_t* freq_ac,
                      uint32_t* freq_dc) {
  Enc* e = (Enc*)calloc(1, sizeof(Enc));
  if (e == NULL || !init_geom(yuv_mode, &e->g)) { free(e); return; }
  RunLevel rl[64];
  DCTCoeffs c;
  for (int m = 0; m < nb_mcus; ++m) {
    for (int k = 0; k < e->g.mcu_blocks; ++k, zz += 64) {
      const int comp = comp_of_block(&e->g, k);
      const int dc = zigzag_to_runlevels(zz, &c, rl);
      c.idx = (int8_t)comp;
      c.dc_code = dc_diff_code(dc, &e->DCs[comp]);
      add_stats(e, &c, rl);
    }
  }
  for (int q = 0; q < 2; ++q) {
    memcpy(freq_ac + 256 * q, e->freq_ac[q], 256 * sizeof(uint32_t));
    memcpy(freq_dc + 12 * q, e->freq_dc[q], 12 * sizeof(uint32_t));
  }
  free(e);
}

/* ------------------------------------------------------------------------------------------
 * Whole encode (enc.cc:391-448 with the scan loops of :276-386).  The reference's methods differ
 * only in what is kept in RAM between passes (enc.cc:121-129); here all coefficients and
 * run/levels are always kept, which gives the same bytes.
 * ---------------------------------------------------------------------------------------- */
void sjo_default_params(sjo_params* p, float quality, int method, int yuv_mode) {
  memset(p, 0, sizeof(*p));
  p->yuv_mode = yuv_mode;
  p->method = method;
  p->pix_fmt = SJO_RGB;
  sjo_quality_to_matrices(quality, p->quant);            /* enc.cc:100-104 */
  memset(p->min_quant, 1, sizeof(p->min_quant));         /* quantize.cc:106-109 */
  p->q_bias = 0x78;
  p->qdelta_max_luma = 12;
  p->qdelta_max_chroma = 1;
}

static size_t encode_impl(const uint8_t* pix, const Planes* planes, int w, int h, int stride, const sjo_params* p,
                          uint8_t** out);

size_t sjo_encode(const uint8_t* pix, int w, int h, int stride, const sjo_params* p,
                  uint8_t** out) {
  if (pix == NULL || out == NULL || p == NULL) return 0;
  const int pstep = (p->pix_fmt == SJO_RGB) ? 3 : 4;
  if (w <= 0 || h <= 0 || abs(stride) < pstep * w) return 0;     /* api.cc:35-36 */
  if (p->yuv_mode == SJO_YUV_SHARP) {                            /* EncoderSharp420, encoders.cc:512-541 */
    if (p->pix_fmt != SJO_RGB) return 0;
    const int cw = (w + 1) / 2, ch = (h + 1) / 2;
    uint8_t* yuv = (uint8_t*)malloc((size_t)w * h + 2 * (size_t)cw * ch);
    if (yuv == NULL) return 0;
    sjo_params q = *p;
    q.yuv_mode = SJO_YUV_420;
    size_t n = 0;
    if (sjo_sharp_yuv(pix, w, h, stride, yuv, yuv + (size_t)w * h, yuv + (size_t)w * h + (size_t)cw * ch)) {
      n = sjo_encode_planar(yuv, w, yuv + (size_t)w * h, cw, yuv + (size_t)w * h + (size_t)cw * ch, cw, 1, w, h, &q, out);
    }
    free(yuv);
    return n;
  }
  return encode_impl(pix, NULL, w, h, stride, p, out);
}

/* Planar / semi-planar input (sjpeg.h:313-349; argument checks of encoders.cc:346-358,421-440,
 * 492-507, api.cc:283-292).  yuv_mode selects the layout: 420 (u,v planes of ((w+1)/2, (h+1)/2);
 * uv_step 2 = one interleaved plane, u and v pointing at its first U and V byte), 444, or 400
 * (y only). */
size_t sjo_encode_planar(const uint8_t* y, int y_stride, const uint8_t* u, int u_stride, const uint8_t* v,
                         int v_stride, int uv_step, int w, int h, const sjo_params* p, uint8_t** out) {
  if (y == NULL || out == NULL || p == NULL) return 0;
  if (w <= 0 || h <= 0 || abs(y_stride) < w) return 0;
  if (p->yuv_mode != SJO_YUV_400) {
    if (u == NULL || v == NULL) return 0;
    const int cw = (p->yuv_mode == SJO_YUV_420) ? uv_step * ((w + 1) / 2) : w;
    if (abs(u_stride) < cw || abs(v_stride) < cw) return 0;
  }
  Planes P = { y, u, v, y_stride, u_stride, v_stride, (p->yuv_mode == SJO_YUV_420) ? uv_step : 1 };
  return encode_impl(NULL, &P, w, h, 0, p, out);
}

static size_t encode_impl(const uint8_t* pix, const Planes* planes, int w, int h, int stride, const sjo_params* p,
                          uint8_t** out) {
  *out = NULL;
  if (w > 65535 || h > 65535) return 0;                          /* enc.cc:406-408 */
  Enc* e = (Enc*)calloc(1, sizeof(Enc));
  if (e == NULL) return 0;
  if (!init_geom(p->yuv_mode, &e->g)) { free(e); return 0; }
  int method = p->method;
  method = (method < 0) ? 0 : (method > 8) ? 8 : method;         /* enc.cc:121-129 */
  const int adaptive = (method >= 3);
  const int optimize = (method != 0) && (method != 3);
  const int trellis = (method >= 7);
  e->W = w; e->H = h; e->yuv_mode = p->yuv_mode;
  e->mb_w = (w + e->g.block_w - 1) / e->g.block_w;
  e->mb_h = (h + e->g.block_h - 1) / e->g.block_h;
  e->bw.ok = 1;
  for (int i = 0; i < 2; ++i) {
    memcpy(e->quants[i].quant, p->quant[i], 64);
    memcpy(e->quants[i].min_quant, p->min_quant[i], 64);
    finalize(&e->quants[i], p->q_bias);                          /* enc.cc:394-395 */
    e->quants[i].codes = e->ac_codes[i];
    std_table(0, i, &e->huff[i]);                                /* entropy.cc:84-86 */
    std_table(1, i, &e->huff[2 + i]);
  }
  const size_t nb_mcus = (size_t)e->mb_w * e->mb_h;
  const size_t nb_blocks = nb_mcus * e->g.mcu_blocks;
  int16_t* coeffs = (int16_t*)malloc(nb_blocks * 64 * sizeof(int16_t));
  DCTCoeffs* infos = (DCTCoeffs*)malloc(nb_blocks * sizeof(DCTCoeffs));
  RunLevel* rls = (RunLevel*)malloc(nb_blocks * 63 * sizeof(RunLevel));
  size_t result = 0;
  if (coeffs == NULL || infos == NULL || rls == NULL) goto end;

  image_to_blocks_ex(pix, planes, w, h, stride, p->yuv_mode, p->pix_fmt, 1, coeffs);

  write_app0(e);
  if (adaptive) {                                                /* enc.cc:425-429 */
    int32_t* counts = (int32_t*)malloc(sizeof(int32_t) * 2 * 64 * (HBINS + 1));
    if (counts == NULL) goto end;
    sjo_collect_histograms(coeffs, (int)nb_mcus, p->yuv_mode, counts);
    uint8_t q2[2][64], mq2[2][64];
    for (int i = 0; i < 2; ++i) { memcpy(q2[i], e->quants[i].quant, 64); memcpy(mq2[i], e->quants[i].min_quant, 64); }
    sjo_analyse_histo(counts, e->g.nb_comps, q2, (const uint8_t(*)[64])mq2, p->qdelta_max_luma,
                      p->qdelta_max_chroma);
    free(counts);
    for (int c = (e->g.nb_comps > 1 ? 1 : 0); c >= 0; --c) {     /* histogram.cc:311-312 */
      memcpy(e->quants[c].quant, q2[c], 64);
      finalize(&e->quants[c], p->q_bias);
    }
  }
  write_dqt(e);
  write_sof(e);

  if (trellis || !optimize) init_codes(e, 0);                    /* enc.cc:334 / headers.cc:222 */
  {
    /* pass 1: quantise every block, DC diffs, (stats).  enc.cc:286-305 / :344-372 */
    RunLevel* rl = rls;
    size_t n = 0;
    const int16_t* in = coeffs;
    for (size_t m = 0; m < nb_mcus; ++m) {
      for (int c = 0; c < e->g.nb_comps; ++c) {
        for (int i = 0; i < e->g.nb_blocks[c]; ++i, ++n, in += 64) {
          const Quantizer* Q = &e->quants[e->g.quant_idx[c]];
          const int dc = trellis ? trellis_quantize_block(in, c, Q, &infos[n], rl)
                                 : quantize_block(in, c, Q, &infos[n], rl);
          infos[n].dc_code = dc_diff_code(dc, &e->DCs[c]);
          if (optimize) add_stats(e, &infos[n], rl);
          rl += infos[n].nb_coeffs;
        }
      }
    }
  }
  if (optimize) compile_stats(e);                                /* enc.cc:374 */
  write_dht(e);
  write_sos(e);
  {
    const RunLevel* rl = rls;
    for (size_t n = 0; n < nb_blocks; ++n) {                     /* enc.cc:309-319 */
      code_block(e, &infos[n], rl);
      rl += infos[n].nb_coeffs;
    }
  }
  bw_pad_flush(&e->bw);                                          /* headers.cc:262-268 */
  bw_byte(&e->bw, 0xff); bw_byte(&e->bw, 0xd9);
  if (e->bw.ok) { *out = e->bw.buf; result = e->bw.pos; e->bw.buf = NULL; }
end:
  free(e->bw.buf);
  free(coeffs); free(infos); free(rls);
  free(e);
  return result;
}

size_t sjo_sjpeg_encode(const uint8_t* rgb, int w, int h, int stride, float quality, int method,
                        int yuv_mode, uint8_t** out) {
  if (rgb == NULL || out == NULL) return 0;
  sjo_params p;
  sjo_default_params(&p, quality, method, yuv_mode);
  return sjo_encode(rgb, w, h, stride, &p, out);
}

void sjo_free(uint8_t* p) { free(p); }

/* A horizontal stripe of whole MCU rows coded on its own (multi-GPU row striping):
 * the same per-block pipeline as sjo_encode() for method 0 (default Huffman tables), but the DC
 * predictors start from dc_pred[] (entropy.cc:133-136 carried across the stripe boundary) and the
 * result is the RAW entropy-coded bit string: no 0xFF stuffing, no padding, no markers.
 * *bits_out (malloc) holds ceil(nbits/8) bytes, MSB first, last byte zero-filled. */
size_t sjo_encode_stripe(const uint8_t* pix, int w, int hs, int stride, const sjo_params* p,
                         const int dc_pred[3], int dc_last[3], uint8_t** bits_out, uint64_t* nbits) {
  if (pix == NULL || p == NULL || bits_out == NULL || nbits == NULL) return 0;
  *bits_out = NULL;
  *nbits = 0;
  Enc* e = (Enc*)calloc(1, sizeof(Enc));
  if (e == NULL) return 0;
  if (!init_geom(p->yuv_mode, &e->g)) { free(e); return 0; }
  e->W = w; e->H = hs; e->yuv_mode = p->yuv_mode;
  e->mb_w = (w + e->g.block_w - 1) / e->g.block_w;
  e->mb_h = (hs + e->g.block_h - 1) / e->g.block_h;
  e->bw.ok = 1;
  e->bw.raw = 1;
  for (int i = 0; i < 2; ++i) {
    memcpy(e->quants[i].quant, p->quant[i], 64);
    memcpy(e->quants[i].min_quant, p->min_quant[i], 64);
    finalize(&e->quants[i], p->q_bias);
    std_table(0, i, &e->huff[i]);
    std_table(1, i, &e->huff[2 + i]);
  }
  init_codes(e, 0);
  for (int c = 0; c < 3; ++c) e->DCs[c] = dc_pred ? dc_pred[c] : 0;
  const size_t nb_mcus = (size_t)e->mb_w * e->mb_h;
  int16_t* coeffs = (int16_t*)malloc(nb_mcus * e->g.mcu_blocks * 64 * sizeof(int16_t));
  size_t result = 0;
  if (coeffs == NULL) goto end;
  sjo_image_to_coeffs(pix, w, hs, stride, p->yuv_mode, p->pix_fmt, coeffs);
  {
    RunLevel rl[64];
    DCTCoeffs info;
    const int16_t* in = coeffs;
    for (size_t m = 0; m < nb_mcus; ++m) {
      for (int c = 0; c < e->g.nb_comps; ++c) {
        for (int i = 0; i < e->g.nb_blocks[c]; ++i, in += 64) {
          const int dc = quantize_block(in, c, &e->quants[e->g.quant_idx[c]], &info, rl);
          info.dc_code = dc_diff_code(dc, &e->DCs[c]);
          code_block(e, &info, rl);
        }
      }
    }
  }
  *nbits = e->bw.nbits;
  bw_flush_bits(&e->bw);
  if (e->bw.nb > 0) bw_byte(&e->bw, (int)(e->bw.acc >> 56));   /* zero-filled tail */
  if (dc_last) for (int c = 0; c < 3; ++c) dc_last[c] = e->DCs[c];
  if (e->bw.ok) { *bits_out = e->bw.buf; result = e->bw.pos; e->bw.buf = NULL; }
end:
  free(e->bw.buf);
  free(coeffs);
  free(e);
  return result;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic inputs.  LCG of tests/unit_test.cc:73-79.
 * ---------------------------------------------------------------------------------------- */
void sjo_make_rgb(char gen, int w, int h, uint32_t seed, uint8_t* out) {
  uint32_t s = seed;
#define RND8() (s = 1103515245u * s + 12345u, (uint8_t)(s >> 16))
  for (int y = 0; y < h; ++y) {
    for (int x = 0; x < w; ++x) {
      uint8_t* p = out + 3 * ((size_t)y * w + x);
      if (gen == 'A') {                     /* tests/unit_test.cc:82-94 */
        p[0] = (uint8_t)(x * 5 + (RND8() >> 3));
        p[1] = (uint8_t)(y * 3 + (RND8() >> 4));
        p[2] = (uint8_t)(((x / 8) ^ (y / 8)) * 51);
      } else {                              /* gen B */
        const int tx = (x >> 2) & 255, ty = (y >> 2) & 255;
        const int a = tx < 128 ? tx : 255 - tx, b = ty < 128 ? ty : 255 - ty;
        p[0] = (uint8_t)(64 + a + (RND8() >> 5));
        p[1] = (uint8_t)(64 + b + (RND8() >> 5));
        p[2] = (uint8_t)(64 + ((a + b) >> 1) + (RND8() >> 5));
      }
    }
  }
#undef RND8
}
