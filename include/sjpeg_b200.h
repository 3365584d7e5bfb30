/*
 * sjpeg_b200.h -- C ABI of the B200 (sm_100a) baseline-JPEG encode path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  It is what a
 * maintainer of webmproject/sjpeg would bind in place of the per-MCU CPU loops
 *   Encoder::Encode()              /root/reference/src/enc.cc:391-448
 *   SinglePassScan[Optimized]()    /root/reference/src/enc.cc:276-386
 *   CollectHistograms()            /root/reference/src/histogram.cc:317-339
 * (see INTEGRATION.md for the reference-side stub).  The classic entry points SjpegEncode() /
 * SjpegCompress() / sjpeg::Encode() (include/sjpeg.h, mirroring /root/reference/src/sjpeg.h)
 * are thin wrappers over sjb_encode().
 *
 * All functions return SJB_OK (0) or a negative error; none throws.  There is no CPU fallback:
 * without a CUDA device every compute entry point returns SJB_ERR_CUDA.
 */
#ifndef SJPEG_B200_H_
#define SJPEG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  SJB_OK = 0,
  SJB_ERR_ARG = -1,       /* null pointer, bad dimension / stride / mode (api.cc:35-36, enc.cc:406) */
  SJB_ERR_CUDA = -2,      /* CUDA runtime error or no device */
  SJB_ERR_NOMEM = -3,     /* host or device allocation failed */
  SJB_ERR_CAPACITY = -4   /* caller's output buffer too small (out_size still reports the need) */
};

/* values of SjpegYUVMode, /root/reference/src/sjpeg.h:54-60.  SHARP and AUTO are accepted by
 * sjb_encode() for packed RGB input only (as in the reference, api.cc:208,235): SHARP runs the
 * iterative sharp RGB->YUV 4:2:0 conversion on the device and then the planar encoder
 * (EncoderSharp420, encoders.cc:512-541); AUTO runs the riskiness analyser first and needs the
 * score table (sjb_set_score_table). */
enum { SJB_YUV_AUTO = 0, SJB_YUV_420 = 1, SJB_YUV_SHARP = 2, SJB_YUV_444 = 3, SJB_YUV_400 = 4 };
/* PixelFormat of /root/reference/src/sjpegi.h (kRGBInput, kBGRAInput, kRGBAInput) */
enum { SJB_PIX_RGB = 0, SJB_PIX_BGRA = 1, SJB_PIX_RGBA = 2 };

/* Encoder settings after EncoderParam -> Encoder::InitFromParam (api.cc:145-181). */
typedef struct sjb_params {
  int yuv_mode;             /* SJB_YUV_*                                                   */
  int method;               /* 0..8, clamped (enc.cc:121-129)                              */
  int pix_fmt;              /* SJB_PIX_*                                                   */
  uint8_t quant[2][64];     /* luma / chroma matrices, natural order (Encoder::quants_[].quant_) */
  uint8_t min_quant[2][64]; /* lower bounds (Encoder::quants_[].min_quant_), 1 (or 0) = none */
  int q_bias;               /* AC rounding bias, enc.cc:46 default 0x78                    */
  int qdelta_max_luma;      /* enc.cc:48 default 12; at most 12 (histogram.cc:179-183)     */
  int qdelta_max_chroma;    /* enc.cc:49 default 1;  at most 12                            */
} sjb_params;

typedef struct sjb_context sjb_context;   /* one per (thread, device): stream + scratch */

uint32_t sjb_version(void);
/* number of CUDA devices visible, or 0 */
int sjb_device_count(void);

int sjb_context_create(int device, sjb_context** ctx);
void sjb_context_destroy(sjb_context* ctx);
/* last CUDA error text of this context (empty string if none) */
const char* sjb_last_error(const sjb_context* ctx);

/* Defaults of Encoder::Encoder + SetQuality + SetCompressionMethod (enc.cc:66-129): what
 * SjpegEncode(rgb,w,h,stride,&out,quality,method,yuv_mode) uses (api.cc:32-49). */
void sjb_params_default(sjb_params* p, float quality, int method, int yuv_mode);
/* GetQFactor + SetQuantMatrix (quantize.cc:77-96) */
void sjb_quality_to_matrices(float quality, uint8_t out[2][64]);

/* Upper bound of the JPEG size for these dimensions (for sizing 'out'). */
size_t sjb_max_output_size(int width, int height, int yuv_mode);

/*
 * Whole encode of one picture = Encoder::Encode() (enc.cc:391-448) for passes == 1.
 *   pix           first row of the picture; host pointer, or device pointer if pix_on_device
 *   stride        bytes between rows, may be negative, |stride| >= bytes_per_pixel * width
 *   out           host buffer of out_capacity bytes (device buffer if out_on_device)
 *   out_size      receives the JPEG size
 */
int sjb_encode(sjb_context* ctx, const uint8_t* pix, int pix_on_device, int width, int height,
               long long stride, const sjb_params* params, uint8_t* out, int out_on_device,
               size_t out_capacity, size_t* out_size);

/*
 * Planar / semi-planar YUV input: what EncodeYUV420 / EncodeYUV444 / EncodeGray / EncodeNV12 /
 * EncodeNV21 do (/root/reference/src/encoders.cc:256-507, sjpeg.h:313-349).  params->yuv_mode selects the layout:
 *   SJB_YUV_420 : y is width x height, u and v are (width+1)/2 x (height+1)/2; uv_step = 1 for
 *                 separate planes, 2 for one interleaved plane (NV12: u = uv, v = uv + 1; NV21:
 *                 v = vu, u = vu + 1; both strides = the plane's stride)
 *   SJB_YUV_444 : three width x height planes        SJB_YUV_400 : y only (u, v ignored)
 * No colour conversion: samples are pixel - 128; clipped MCUs replicate each plane's last row and
 * column.  Other arguments as sjb_encode (out == NULL returns the size with SJB_ERR_CAPACITY).
 */
int sjb_encode_planar(sjb_context* ctx, const uint8_t* y, long long y_stride, const uint8_t* u,
                      long long u_stride, const uint8_t* v, long long v_stride, int uv_step, int on_device,
                      int width, int height, const sjb_params* params, uint8_t* out, int out_on_device,
                      size_t out_capacity, size_t* out_size);

/*
 * Sharp RGB -> YUV 4:2:0 conversion on its own = sjpeg::ApplySharpYUVConversion
 * (/root/reference/src/yuv_convert.cc:671-695, sjpegi.h:118): packed RGB in, three tightly packed
 * planes out (y: width x height; u, v: (width+1)/2 x (height+1)/2).  Stage-level entry point of
 * the SJB_YUV_SHARP path (parity tests, callers that want the planes).
 */
int sjb_sharp_yuv(sjb_context* ctx, const uint8_t* rgb, int rgb_on_device, int width, int height,
                  long long stride, uint8_t* y, uint8_t* u, uint8_t* v, int out_on_device);

/*
 * Riskiness analyser = SjpegRiskiness (/root/reference/src/sjpeg.h:139-150, jpeg_tools.cc:177-236):
 * recommends 4:2:0 / sharp 4:2:0 / 4:4:4 / 4:0:0 for a packed RGB picture and reports the 0..100
 * risk score.  The analyser looks pixel triples up in the reference's GENERATED 343 x 343 score
 * table (/root/reference/src/score_7.cc, sjpeg::kSharpnessScore); that table is a data asset of
 * the reference and is not reproduced in this repository.  The library takes it from, in order:
 * sjb_set_score_table() (the host binding passes its own array once per process);
 * the file SJPEG_B200_SCORE_TABLE names (117649 bytes); sjpeg::kSharpnessScore of a reference libsjpeg
 * loaded in the same process; sjpeg_score_table.bin next to libsjpeg_b200.so, which csrc/Makefile
 * writes at build time from the reference's score_7.cc where that source tree is present.
 * Without a table sjb_riskiness and SJB_YUV_AUTO return SJB_ERR_ARG (never another mode silently).
 */
int sjb_set_score_table(const uint8_t* table, size_t size /* 343*343 */);   /* NULL clears */
int sjb_has_score_table(void);
int sjb_riskiness(sjb_context* ctx, const uint8_t* rgb, int rgb_on_device, int width, int height,
                  long long stride, int* yuv_mode, float* risk);

/*
 * Multi-pass search for a target size or PSNR = Encoder::LoopScan (/root/reference/src/dichotomy.cc:113-205,
 * used when EncoderParam::passes > 1, api.cc:169-176).  The unquantised coefficients (and, for the
 * adaptive methods, their histogram) stay in device memory; every pass re-quantises them with the
 * matrices the callbacks propose and measures the size (symbol statistics / exact bit count incl.
 * stuffing, dichotomy.cc:210-298) or the PSNR (quantize.cc:547-559, dichotomy.cc:302-323).
 * The callbacks are SearchHook's virtuals (sjpeg.h:355-372).  Arm a search with
 * sjb_context_set_search(); it applies to the NEXT sjb_encode / sjb_encode_planar call on the
 * context and is then cleared.
 */
typedef struct sjb_search {
  int passes;                 /* 2..20 */
  int for_size;               /* 1: target is a size in bytes, 0: a PSNR in dB */
  float target;
  size_t header_extra_bytes;  /* metadata part of Encoder::HeaderSize(), dichotomy.cc:212-229 */
  void* user;
  void (*begin_pass)(void* user, int pass);
  void (*next_matrix)(void* user, int idx, uint8_t dst[64]);   /* SearchHook::NextMatrix */
  int (*update)(void* user, float result);                     /* SearchHook::Update: non-zero = done */
  /* filled on return: index of the pass whose matrices were kept, and its measured value */
  int best_pass;
  float best_result;
} sjb_search;
int sjb_context_set_search(sjb_context* ctx, sjb_search* search /* NULL disarms */);

/* Copies the JPEG produced by the most recent sjb_encode() of this context (it stays resident in
 * device memory until the next encode).  Lets a caller learn the size first -- sjb_encode with
 * out == NULL returns SJB_ERR_CAPACITY and the size -- and then fetch into an exact allocation. */
int sjb_fetch_output(sjb_context* ctx, uint8_t* out, int out_on_device, size_t out_capacity);

/*
 * Batch of independent pictures with identical geometry and settings (config 5 of BASELINE.json:
 * frames are the unit of sharding).  pix[i] / out[i] as above; sizes[i] receives each size.
 * Pictures are processed in groups (one kernel launch covers a whole group) and the groups are
 * overlapped on the context's streams; for methods >= 1 the host phases (matrices from the
 * histograms, optimal Huffman tables) of one group run while younger groups are already queued.
 * params->yuv_mode may also be SJB_YUV_AUTO (packed RGB: the analyser runs on every picture, the
 * batch is split by the mode each one gets) or SJB_YUV_SHARP (conversions of several pictures run
 * concurrently, then the planar pipeline) -- the batched form of SjpegCompress().
 */
int sjb_encode_batch(sjb_context* ctx, int n, const uint8_t* const* pix, int pix_on_device, int width,
                     int height, long long stride, const sjb_params* params, uint8_t* const* out,
                     int out_on_device, size_t out_capacity, size_t* sizes);

/* Batch of planar / semi-planar pictures of one geometry (layouts as sjb_encode_planar; the strides and
 * uv_step are common to the batch): y[i] / u[i] / v[i] are the planes of picture i (u, v may be NULL for
 * SJB_YUV_400).  Same pipeline, group sizes and overlap as sjb_encode_batch -- the hand-off for decoded
 * video frames (NV12) that are already in device memory (on_device = 1). */
int sjb_encode_planar_batch(sjb_context* ctx, int n, const uint8_t* const* y, long long y_stride, const uint8_t* const* u,
                            long long u_stride, const uint8_t* const* v, long long v_stride, int uv_step, int on_device,
                            int width, int height, const sjb_params* params, uint8_t* const* out, int out_on_device,
                            size_t out_capacity, size_t* sizes);

/*
 * Row stripes: a picture split into horizontal stripes of whole MCU rows, one stripe per GPU
 * (BASELINE.json config 5; the reference has no restart markers, headers.cc:242-258, so the
 * stripes of one picture form a single bit string and need two tiny exchanges between ranks).
 * A session holds this rank's stripe of each of n pictures.  Method 0 (default tables) only.
 *   1. transform: colour convert + fDCT + quantise; last_dc[n][3] = quantised DC of the last
 *      Y / U / V block of every stripe                         -> exchange with the next rank
 *   2. code:      entropy-code with dc_pred[n][3] = last_dc of the previous rank's stripe (zeros
 *      on the first rank, entropy.cc:155-159); bits[n] = bit count  -> all-gather, prefix sum
 *   3. finish:    byte-align to the global bit offset, 0xFF-stuff (bit_writer.h:172-196) the bytes
 *      owned entirely by this stripe into out[i]; the byte shared with the previous stripe
 *      (head_byte, present iff offset % 8 != 0) and with the next one (tail_byte, tail_bits of it
 *      ours) are returned for the gatherer to OR together and stuff.  The last rank pads with
 *      1-bits and appends EOI (headers.cc:262-268).  sjb_picture_header gives the bytes before
 *      the scan.
 */
typedef struct sjb_stripes sjb_stripes;
int sjb_stripes_create(sjb_context* ctx, int n, int width, int stripe_height, const sjb_params* params,
                       sjb_stripes** out);
void sjb_stripes_destroy(sjb_stripes* s);
int sjb_stripes_transform(sjb_stripes* s, const uint8_t* const* pix, int pix_on_device, long long stride,
                          int* last_dc);
int sjb_stripes_code(sjb_stripes* s, const int* dc_pred, unsigned long long* bits);
int sjb_stripes_finish(sjb_stripes* s, const unsigned long long* bit_offsets, int is_first, int is_last,
                       uint8_t* const* out, size_t out_capacity, size_t* sizes, unsigned char* head_byte,
                       unsigned char* tail_byte, unsigned char* tail_bits);
/* SOI..SOS bytes of a method-0 picture (headers.cc:48-61,182-258); host only */
int sjb_picture_header(const sjb_params* params, int width, int height, uint8_t* out, size_t out_capacity,
                       size_t* out_size);

/*
 * Row stripes with the exchange inside the library (NCCL over NVLink, bound at run time with
 * dlopen("libnccl.so.2"); nothing of this is needed, or loaded, for single-GPU use).  One process
 * per GPU.  Bootstrap like any NCCL program: rank 0 calls sjb_comm_unique_id(), the 128 bytes reach
 * the other ranks by whatever channel the host program has (torch.distributed, MPI, a file), every
 * rank calls sjb_comm_create() (collective).
 *
 * sjb_stripes_encode() is collective: rank r passes, for each of the n pictures, the pixel rows
 * [y0, y1) that sjb_stripe_rows(height, yuv_mode, world, r, ...) assigns to it -- whole MCU rows,
 * balanced, the last stripe owning a clipped MCU row; ranks beyond the number of MCU rows get an
 * empty range and pass any non-NULL pointers -- and rank 0 receives the n complete JPEGs in out[]
 * (host buffers of out_capacity bytes) and their sizes.  All compression methods 0..8: adapted
 * matrices and optimised Huffman tables are derived from all-reduced histograms / symbol counts
 * (SinglePassScanOptimized enc.cc:323-386, CollectHistograms histogram.cc:317-339), identically on
 * every rank.  Output bytes equal the single-GPU encode of the whole picture.
 */
typedef struct sjb_comm sjb_comm;
int sjb_comm_unique_id(uint8_t id[128]);
int sjb_comm_create(sjb_context* ctx, const uint8_t id[128], int rank, int world, sjb_comm** comm);
void sjb_comm_destroy(sjb_comm* comm);
int sjb_stripe_rows(int height, int yuv_mode, int world, int rank, int* y0, int* y1);
int sjb_stripes_encode(sjb_comm* comm, int n, const uint8_t* const* pix, int pix_on_device, int width, int height,
                       long long stride, const sjb_params* params, uint8_t* const* out, size_t out_capacity,
                       size_t* sizes);
/* Frames sharded across the ranks: collects on rank 0 the JPEGs every rank left in DEVICE memory
 * (sjb_encode_batch with out_on_device = 1) -- rank order, back to back in blob (host memory), their
 * sizes in out_sizes[0 .. *n_total).  Collective; blob / out_sizes / n_total are used on rank 0 only. */
int sjb_gather_frames(sjb_comm* comm, int n_local, const uint8_t* const* dev_jpegs, const size_t* sizes_local,
                      uint8_t* blob, size_t blob_capacity, size_t* out_sizes, int out_sizes_capacity, int* n_total);
/* Host-only merge used by rank 0 (exported for tests): header + the bytes each stripe emitted, in
 * order, OR-merging the byte neighbouring stripes share and stuffing it (bit_writer.h:172-196).
 * flags[r] = head_byte | tail_byte << 8 | tail_bits << 16 | head_open << 24: the stripe's share of
 * the byte it begins in / ends in, the number of bits it owns of the latter (0 = ends on a byte
 * boundary), and whether it ends inside its head byte (a stripe shorter than the gap it starts in). */
int sjb_stripes_assemble(const uint8_t* header, size_t header_len, int stripes, const uint8_t* const* part,
                         const size_t* size, const unsigned* flags, uint8_t* out, size_t out_capacity,
                         size_t* out_size);

/* Pinned host memory helpers (for callers that want async copies at PCIe speed). */
void* sjb_host_alloc(size_t bytes);
/* Same, write-combined (cudaHostAllocWriteCombined): not cached on the host side, so the device reads
 * it over PCIe without snooping the CPU caches -- for input buffers the host only ever WRITES
 * sequentially (reading them back on the host is very slow). */
void* sjb_host_alloc_wc(size_t bytes);
void sjb_host_free(void* p);

/* ---- stage-level entry points (used by the parity tests and the benchmark) ---------------- */

/* F1 only: colour convert + fDCT (+ quantise).  coef (host, int16[nb_blocks*64]):
 *   quantise = 0: unquantised x16 coefficients, natural order
 *   quantise = 1: quantised values in zig-zag order; nzmask (host, uint8[nb_blocks], may be NULL)
 *                 receives the kernels' side output: bit c set <=> positions 8c..8c+7 not all 0
 * Blocks are in scan order: MCU raster, Y.. U V inside an MCU (enc.cc:286-305). */
int sjb_stage_coefficients(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                           const sjb_params* params, int quantise, int16_t* coef, uint8_t* nzmask);
/* histogram.cc:99-108 over the whole picture: counts = int32[2][64][129] (host) */
int sjb_stage_histogram(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                        const sjb_params* params, int32_t* counts);
/* histogram.cc:126-315 on the device (kernels A1) after the histogram above: the matrices the adaptive
 * methods (3..8) put into the DQT segment for this picture, quant[2][64] in natural order, clamped to
 * params->min_quant as FinalizeQuantizer does (quantize.cc:116-148) */
int sjb_stage_adapted_matrices(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                               const sjb_params* params, uint8_t quant[2][64]);
/* entropy.cc:208-227 over the whole picture for the plain quantiser: freq_ac[2][256], freq_dc[2][12] */
int sjb_stage_symbol_stats(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                           const sjb_params* params, uint32_t* freq_ac, uint32_t* freq_dc);

/* Time (ms, CUDA events on the context's stream) of the kernels of the last sjb_encode call:
 * [0] F1 (all launches), [1] entropy stage (memset + E + S), [2] whole device pipeline. */
int sjb_last_timings(const sjb_context* ctx, float ms[3]);

/* Per-kernel-stage times (ms, CUDA events recorded right before and after each stage's launches) of
 * the last TIMED group on the context's first lane -- the group of a single sjb_encode, or the last
 * lane-0 group of sjb_bench_device: [0] F1, [1] H1 histogram, [2] A1 histogram analysis + Q1 re-quantise or T1 trellis
 * (incl. its block sort), [3] S1 symbol statistics, [4] E entropy coder, [5] S byte stuffing;
 * -1 for a stage the method does not run.  *frames = pictures in that group (one launch each). */
int sjb_last_stage_timings(sjb_context* ctx, float ms[6], int* frames);

/* Device-resident benchmark loop: encodes the n device pictures round-robin 'iters' times with
 * inputs and outputs staying in HBM; returns total elapsed ms (CUDA events) and, optionally, the
 * average duration of the fused F1 kernel per picture in f1_ms. */
int sjb_bench_device(sjb_context* ctx, int n, const uint8_t* const* dev_pix, int width, int height,
                     long long stride, const sjb_params* params, int iters, float* total_ms,
                     float* f1_ms, size_t* jpeg_bytes, unsigned long long* launches);

/* Host copy of the JPEG that the LAST round of the most recent sjb_bench_device call left in HBM
 * for picture `index` (so that the bytes of the timed path itself can be checked).  SJB_ERR_ARG
 * when a later group of that call has reused the picture's buffers (more groups per round than
 * lanes): bench fewer pictures per call to verify them all. */
int sjb_bench_output(sjb_context* ctx, int index, uint8_t* out, size_t out_capacity, size_t* out_size);

/* Times ONLY the fused F1 kernel (convert + fDCT + quantise): launches it back to back over the
 * n device pictures in groups of *frames_per_launch pictures (the group size the encoder itself
 * uses for this geometry), 'iters' rounds, on one stream; ms_per_launch = CUDA-event time /
 * number of launches.  n pictures larger than L2 in total keep the input cold. */
int sjb_bench_f1(sjb_context* ctx, int n, const uint8_t* const* dev_pix, int width, int height,
                 long long stride, const sjb_params* params, int iters, float* ms_per_launch,
                 int* frames_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* SJPEG_B200_H_ */
