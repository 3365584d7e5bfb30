// kernels.cuh -- launch wrappers of the sm_100a kernels (kernels.cu).  Host code (engine.cu)
// sees only these plain functions; every pointer is a device pointer unless stated.
//
// All kernels work on a GROUP of up to kMaxGroup pictures of identical geometry in one launch
// (gridDim.y = picture): a single 4K picture is only ~2000 warp-tiles, too few to balance 592
// warp schedulers, and per-picture launches of the small entropy kernels are latency bound.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "block_ops.cuh"

namespace sjb {

// 16 pictures per launch (8 in round 1): +3.5 % on the 16-frame 4K pipeline (458 -> 474 Gpix/s gen B,
// 200 -> 205 gen A, same box), the last partly filled wave of CTAs of every kernel weighing half as much.
#ifndef SJB_MAX_GROUP
#define SJB_MAX_GROUP 16
#endif
enum { kMaxGroup = SJB_MAX_GROUP };
#ifndef SJB_TILE_BLOCKS
#define SJB_TILE_BLOCKS 256
#endif
// 8x8 blocks per tile of the entropy kernel (= worker threads per CTA).  Measured on B200, 4K gen B, entropy +
// stuffing per 4 pictures: 128 -> 71 us (twice the tiles in the look-back chain), 256 -> 56 us, 512 with 256-bit
// slots -> 54 us but 30 % slower on busy pictures (more blocks overflow the slots); 256 is the default.
enum { kTileBlocks = SJB_TILE_BLOCKS };
#ifndef SJB_STUFF_THREADS
#define SJB_STUFF_THREADS 256
#endif
enum { kStuffThreads = SJB_STUFF_THREADS, kStuffTileBytes = 16 * kStuffThreads };   // stream bytes per CTA iteration of the stuffing kernel

struct FrameSet {
  const uint8_t* pix[kMaxGroup];   // row 0 of each picture (packed RGB/RGBA/BGRA, or the Y plane)
  const uint8_t* pix_u[kMaxGroup]; // planar input: row 0 of the U and V planes (for NV12/NV21: the
  const uint8_t* pix_v[kMaxGroup]; //   first U and V byte of the interleaved plane)
  long long stride_u, stride_v;
  int planar;                      // 0 = packed pixels, 1 = planar / semi-planar YUV (encoders.cc:256-507)
  int uv_step;                     // bytes between chroma samples: 1 planar, 2 interleaved
  int frames;
  long long stride;                // bytes between rows, may be negative
  int width, height;
  int yuv_mode;                    // kYuv420 / kYuv444 / kYuv400
  int pix_fmt;                     // kFmtRGB / kFmtBGRA / kFmtRGBA
  int mcus_x, mcus_y;
  int mcu_blocks, luma_blocks;
  unsigned blocks_per_frame;
};

// Device-side scalars of one picture (one 32-byte record).
struct StreamInfo {
  unsigned long long total_bits;    // entropy-coded bits before padding
  unsigned long long stuffed_bytes; // number of 0xFF bytes that got a 0x00 appended
  unsigned long long out_size;      // header + scan (+ EOI) bytes written
  unsigned char head_byte;          // stripe mode: first, shared byte (low 8-shift bits are ours)
  unsigned char tail_byte;          // stripe mode: last, shared byte (top tail_bits bits are ours)
  unsigned char tail_bits;          // 0 = no shared tail byte
  unsigned char head_open;          // stripe mode: the stripe ENDS inside its head byte (owns no byte boundary)
  unsigned char pad[4];
};

// Per-lane device arrays, picture-major with fixed pitches (in elements of the pointed type).
struct GroupBuffers {
  int16_t* coef;            size_t coef_pitch;    // [frames][padded blocks*64], sector-interleaved (block_ops.cuh)
  uint8_t* nzmask;          size_t mask_pitch;    // [frames][blocks] chunk bitmaps
  uint32_t* words;          size_t words_pitch;   // [frames][worst-case stream words], zero between encodes
  uint8_t* out;             size_t out_pitch;     // [frames][worst-case file bytes]
  unsigned long long* bit_state; size_t bit_state_pitch;   // look-back descriptors of the entropy kernel
  unsigned long long* ff_state;  size_t ff_state_pitch;    // look-back descriptors of the stuffing kernel
  StreamInfo* info;         // [frames]
  CodeTabs* tabs;           // [frames]
  QuantTabs* qtabs;         // [frames]  (adaptive methods: per-picture matrices)
  int32_t* hist;            // [frames][2][64][129]
  uint32_t* freq;           // [frames][2][272]
  uint8_t* quant;           // [frames][2][64]
  const int* dc_init;       // [frames][3] DC predictors at the first MCU (Y,U,V); null = zeros
  const unsigned long long* bit_offsets;   // stripes: [frames] global bit offset of each stripe (device); the
                                           // stuffing kernel then takes shift = offset % 8 from here
};

// Per-picture arguments of the stuffing kernel.  A whole picture uses shift 0 and
// kStuffFirst|kStuffLast.  A row stripe of a picture split across GPUs starts at
// global bit offset O: shift = O % 8 re-aligns its bits to the global byte grid; the byte it
// shares with the previous stripe (head) and the one it shares with the next (tail) are not
// emitted but reported in StreamInfo for the gatherer to merge.
enum { kStuffFirst = 1, kStuffLast = 2, kStuffKeepWords = 4 };
struct StuffArgs {
  unsigned header_len[kMaxGroup];
  unsigned shift[kMaxGroup];
  unsigned flags[kMaxGroup];
};

// F1: colour convert + fDCT (+ quantise) for the MCU rectangle [mx0,mx1) x [my0,my1) of every
// picture of the set.
//   raw = true : coef receives the unquantised x16 coefficients, natural order
//   raw = false: coef receives quantised values in zig-zag order, nzmask the non-zero CHUNK bitmaps
//                (bit c set <=> zig-zag positions 8c..8c+7 are not all zero; bit 0 includes the DC)
// generic path: any stride / alignment / pixel format, edge replication (encoders.cc:157-253)
void LaunchF1Generic(const FrameSet& fs, int mx0, int my0, int mx1, int my1, bool raw,
                     const QuantTabs& qt, const GroupBuffers& gb, cudaStream_t s);
// fast path: RGB24, full MCUs only, rows 16-byte aligned (stride % 16 == 0, every base % 16 == 0);
// MCU rows [my0,my1) x columns [0, mx_full).  Bulk-async (TMA engine) staged strips.
bool F1FastEligible(const FrameSet& fs);
void LaunchF1Fast(const FrameSet& fs, int mx_full, int my0, int my1, bool raw, const QuantTabs& qt,
                  const GroupBuffers& gb, cudaStream_t s);

// Q1: quantise stored raw coefficients in place (natural -> zig-zag) + bitmap; tables gb.qtabs[frame]
// raw_src = null: in place (gb.coef holds the raw values); else raw values are read from raw_src
// (same pitch as gb.coef) and the result goes to gb.coef, leaving the raw copy intact (search).
void LaunchRequantize(const FrameSet& fs, const GroupBuffers& gb, const int16_t* raw_src, cudaStream_t s);
// PSNR search: err[frame] += sum of squared quantisation errors of the raw coefficients
void LaunchQuantError(const FrameSet& fs, const GroupBuffers& gb, const int16_t* raw, unsigned long long* err,
                      cudaStream_t s);
// H1: histogram of |coef| >> 2 per matrix and position into gb.hist[frame] (pre-zeroed)
void LaunchHistogram(const FrameSet& fs, const GroupBuffers& gb, cudaStream_t s);
// A1: adaptive quantisation, the analysis of the histograms on the device (histogram.cc:126-315 +
// quantize.cc:116-148): gb.hist[frame] -> gb.quant[frame] (the matrices that go into the DQT, natural
// order) and gb.qtabs[frame] (the quantiser constants Q1 / T1 read).  The settings are the same for
// every picture of the group.  fit: scratch, AqFit[frames][2][64]; fail[frame] is set to 1 when a
// derived matrix entry cannot be expressed by the fused quantiser (the host then refuses the encode,
// as FinalizeQuantizer does).  Two launches.
struct AqParams {
  uint8_t quant0[2][64];      // starting matrices, natural order
  uint8_t min_quant[2][64];
  int qdelta_max[2];          // luma, chroma (enc.cc:48-49)
  int q_bias;
  int nb_comps;               // 1: luma only (the chroma table is filled with the luma one's constants)
};
enum { kAnalyseLaunches = 2 };
void LaunchAnalyseHistograms(int frames, const GroupBuffers& gb, const AqParams& ap, AqFit* fit, int* fail, cudaStream_t s);
// T1: trellis quantisation of raw coefficients in place (quantize.cc:388-457) + bitmap; tables
// gb.qtabs[frame], matrices gb.quant[frame], rate from the AC code lengths in gb.tabs[frame]
// sort_state: uint32[frames][128] scratch; perm: uint32[frames][perm_pitch >= blocks] scratch (the order
// in which the blocks are handed to the threads: sorted by work).  Three launches + one memset.
enum { kTrellisLaunches = 3 };
void LaunchTrellis(const FrameSet& fs, const GroupBuffers& gb, const int16_t* raw_src, uint32_t* sort_state, uint32_t* perm,
                   size_t perm_pitch, cudaStream_t s);
// S1: symbol statistics into gb.freq[frame] (slot < 256 AC symbol, 256+n DC size), pre-zeroed
void LaunchSymbolStats(const FrameSet& fs, const GroupBuffers& gb, cudaStream_t s);

// E: bits per block, decoupled look-back prefix over 256-block tiles, bit packing into gb.words;
// the last tile of each picture writes info.total_bits.  gb.bit_state must be zero on entry.
void LaunchEntropyPack(const FrameSet& fs, const GroupBuffers& gb, cudaStream_t s);
// S: 0xFF stuffing with a decoupled look-back over 4 KB stream tiles: scatter to out + header_len,
// padding, EOI, info.out_size; zeroes the consumed stream words.  gb.ff_state zero on entry.
void LaunchStuff(const FrameSet& fs, const GroupBuffers& gb, const StuffArgs& args, cudaStream_t s);
// quantised DC of the last Y / U / V block of every picture -> out[frames][3] (stripe hand-over)
void LaunchLastDc(const FrameSet& fs, const GroupBuffers& gb, int* out, cudaStream_t s);

// ---- row stripes across GPUs: the small device-side steps between the collectives (engine) ----
// bits[f] = info[f].total_bits
void LaunchStripeBits(const GroupBuffers& gb, int frames, unsigned long long* bits, cudaStream_t s);
// offsets[i] = sum of all_bits[r][i] over the ranks r < rank (all_bits = [world][n], ranks without
// rows contribute 0)
void LaunchStripeOffsets(const unsigned long long* all_bits, int n, int rank, unsigned long long* offsets, cudaStream_t s);
// meta[f] = {out_size, head_byte | tail_byte << 8 | tail_bits << 16}
void LaunchStripeMeta(const GroupBuffers& gb, int frames, unsigned long long* meta, cudaStream_t s);
// small copy between pinned host and device memory (either direction) done by a kernel, so that it
// does not queue behind bulk transfers on the copy engine
void LaunchCopySmall(void* dst, const void* src, size_t bytes, cudaStream_t s);
// packs the emitted bytes of the n stripes (slot f of its group at src[f], sizes in meta[f*2]) back to
// back into dst; one launch per group: dst_offset_base = where the group's first stripe goes is
// computed on the device from the sizes of all stripes before it (meta of the whole batch)
void LaunchStripeCompact(const uint8_t* group_out, size_t out_pitch, int first, int frames, const unsigned long long* meta_all_local,
                         uint8_t* dst, cudaStream_t s);

}  // namespace sjb
