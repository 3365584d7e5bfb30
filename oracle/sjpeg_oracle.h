/*
 * sjpeg_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C (C99) CPU restatement of the baseline-JPEG encode path of webmproject/sjpeg
 * (reference tree: /root/reference, commit 6b8cd89).  It exists so that the CUDA path in
 * sjpeg_b200/ can be checked bit-for-bit, stage by stage, on a box that does not have the
 * reference sources.  Nothing in the product (sjpeg_b200/, include/) may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * PARITY PIN: the reference's own tests hold no golden vectors for this path;
 * the restatement is pinned instead against the UNMODIFIED reference compiled here
 * (oracle/_ref/libsjpeg_ref.so, built by oracle/Makefile) -- byte equality of whole JPEG
 * files over the matrix in tests/test_oracle.py -- and against the md5 table the
 * compiled reference produced (tests/golden/ref_md5.json, made by tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef SJPEG_ORACLE_H_
#define SJPEG_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* yuv modes, numbering of sjpeg.h:54-60 */
enum { SJO_YUV_AUTO = 0, SJO_YUV_420 = 1, SJO_YUV_SHARP = 2, SJO_YUV_444 = 3, SJO_YUV_400 = 4 };
/* pixel formats (sjpegi.h PixelFormat) */
enum { SJO_RGB = 0, SJO_BGRA = 1, SJO_RGBA = 2 };

typedef struct {
  int yuv_mode;            /* 420 / 444 / 400, or SHARP (RGB input only) */
  int method;              /* 0..8, clamped like enc.cc:121-129 */
  int pix_fmt;             /* SJO_RGB / SJO_BGRA / SJO_RGBA */
  uint8_t quant[2][64];    /* natural order, luma then chroma */
  uint8_t min_quant[2][64];
  int q_bias;              /* enc.cc:46 default 0x78 */
  int qdelta_max_luma;     /* enc.cc:48 default 12 */
  int qdelta_max_chroma;   /* enc.cc:49 default 1 */
} sjo_params;

/* defaults of Encoder::Encoder (enc.cc:66-91) + SetQuality(quality) + SetCompressionMethod */
void sjo_default_params(sjo_params* p, float quality, int method, int yuv_mode);

/* quantize.cc:77-96 : quality -> the two 8-bit matrices */
void sjo_quality_to_matrices(float quality, uint8_t out[2][64]);

/* quantize.cc:123-148 : iquant / bias / qthresh of one matrix (clamps quant[] to min_quant[]) */
void sjo_finalize_quantizer(uint8_t quant[64], const uint8_t min_quant[64], int q_bias,
                            uint16_t iquant[64], uint16_t bias[64], uint16_t qthresh[64]);

/* Geometry (encoders.cc:32-88, enc.cc:410-411) */
void sjo_geometry(int yuv_mode, int w, int h, int* mcu_w, int* mcu_h, int* mcu_blocks,
                  int* mb_w, int* mb_h);

/* Stage 1+2: for every MCU in raster order, fetch samples (with edge replication and
 * AverageExtraLuma) and run the integer fDCT.  out = int16[mb_w*mb_h*mcu_blocks][64], natural
 * order, x16 scaled.  (encoders.cc:157-253, colors_rgb.cc:785-879,1212-1232, fdct.cc:596-609) */
void sjo_image_to_coeffs(const uint8_t* pix, int w, int h, int stride, int yuv_mode, int pix_fmt,
                         int16_t* out);
/* same, without the fDCT (the converted samples) */
void sjo_image_to_samples(const uint8_t* pix, int w, int h, int stride, int yuv_mode, int pix_fmt,
                          int16_t* out);
/* fdct.cc:596-609 on nb blocks, in place */
void sjo_fdct(int16_t* blocks, int nb);

/* Stage 3: plain quantiser (quantize.cc:288-320) applied to every block of coeffs[]; writes
 * the signed quantised values in ZIG-ZAG order, int16[nblocks][64] (index 0 = DC).  comp_of_block
 * is implied by yuv_mode (block k of an MCU). */
void sjo_quantize_image(const int16_t* coeffs, int nb_mcus, int yuv_mode,
                        const uint8_t quant[2][64], const uint8_t min_quant[2][64], int q_bias,
                        int16_t* out_zigzag);
/* trellis variant (quantize.cc:388-457), rate from the default AC tables (enc.cc:334) */
void sjo_trellis_quantize_image(const int16_t* coeffs, int nb_mcus, int yuv_mode,
                                const uint8_t quant[2][64], const uint8_t min_quant[2][64],
                                int q_bias, int16_t* out_zigzag);

/* histogram.cc:99-108,317-339: counts[2][64][129] += over all blocks */
void sjo_collect_histograms(const int16_t* coeffs, int nb_mcus, int yuv_mode, int32_t* counts);
/* histogram.cc:126-315: rewrites quant[][] in place from the histograms */
void sjo_analyse_histo(const int32_t* counts, int nb_comps, uint8_t quant[2][64],
                       const uint8_t min_quant[2][64], int qdelta_max_luma, int qdelta_max_chroma);

/* entropy.cc:208-227: symbol statistics from zig-zag quantised blocks.
 * freq_ac[2][256], freq_dc[2][12] */
void sjo_symbol_stats(const int16_t* zz, int nb_mcus, int yuv_mode, uint32_t* freq_ac,
                      uint32_t* freq_dc);
/* entropy.cc:254-430: returns nb_syms; bits[16], syms[<=256] */
int sjo_build_optimal_table(const uint32_t* freq, int size, uint8_t bits[16], uint8_t* syms);

/* Whole encode.  Returns size, *out is malloc()ed (free with sjo_free).  0 on refusal. */
size_t sjo_encode(const uint8_t* pix, int w, int h, int stride, const sjo_params* p,
                  uint8_t** out);
/* SjpegEncode() equivalent for RGB input (api.cc:32-49), yuv_mode in {420,444,400} */
size_t sjo_sjpeg_encode(const uint8_t* rgb, int w, int h, int stride, float quality, int method,
                        int yuv_mode, uint8_t** out);
/* Planar (EncodeYUV420 / EncodeYUV444 / EncodeGray) and semi-planar (EncodeNV12 / EncodeNV21)
 * input, encoders.cc:256-507.  p->yuv_mode selects 420 / 444 / 400; uv_step = 2 for NV12/NV21
 * (u, v point at the first U and V byte of the interleaved plane). */
size_t sjo_encode_planar(const uint8_t* y, int y_stride, const uint8_t* u, int u_stride, const uint8_t* v,
                         int v_stride, int uv_step, int w, int h, const sjo_params* p, uint8_t** out);
void sjo_free(uint8_t* p);

/* "Sharp" RGB -> YUV 4:2:0 conversion (yuv_convert.cc:671-695, oracle/sjpeg_oracle_sharp.c): y is
 * W x H, u and v are ((W+1)/2) x ((H+1)/2), tightly packed.  sjo_encode()/sjo_sjpeg_encode() with
 * yuv_mode SJO_YUV_SHARP run it and then the planar 4:2:0 encoder, as EncoderSharp420 does
 * (encoders.cc:512-541).  Returns 0 on allocation failure. */
int sjo_sharp_yuv(const uint8_t* rgb, int W, int H, int stride, uint8_t* y, uint8_t* u, uint8_t* v);
/* SjpegRiskiness (jpeg_tools.cc:177-236).  table = the reference's 343 x 343 score table
 * (score_7.cc; generated data, not restated -- tests pass the compiled reference's copy). */
int sjo_riskiness(const uint8_t* rgb, int width, int height, int stride, const uint8_t* table, float* risk);

/* Row stripe coded on its own (method 0): raw entropy-coded bits, DC predictors in/out.
 * Used as the CPU stand-in stage by the multi-rank (gloo) tests of the stripe exchange. */
size_t sjo_encode_stripe(const uint8_t* pix, int w, int hs, int stride, const sjo_params* p,
                         const int dc_pred[3], int dc_last[3], uint8_t** bits_out, uint64_t* nbits);

/* Deterministic synthetic inputs of SURVEY.md 8(d): gen 'A' = tests/unit_test.cc:73-94 MakeRGB,
 * gen 'B' = photo-like ramp + noise.  Packed RGB, stride 3*w. */
void sjo_make_rgb(char gen, int w, int h, uint32_t seed, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif  /* SJPEG_ORACLE_H_ */
