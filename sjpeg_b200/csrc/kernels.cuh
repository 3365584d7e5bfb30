// kernels.cuh -- launch wrappers of the sm_100a kernels (kernels.cu).  Host code (engine.cu)
// sees only these plain functions; every pointer is a device pointer unless stated.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "block_ops.cuh"

namespace sjb {

enum { kTileBlocks = 256 };        // 8x8 blocks per CTA in the entropy kernels (scan tile)
enum { kStuffTileBytes = 4096 };   // stream bytes per CTA tile in the stuffing kernels

struct ImageDesc {
  const uint8_t* pix;   // row 0 of the image (device)
  long long stride;     // bytes between rows, may be negative
  int width, height;
  int yuv_mode;         // kYuv420 / kYuv444 / kYuv400
  int pix_fmt;          // kFmtRGB / kFmtBGRA / kFmtRGBA
  int mcus_x, mcus_y;
};

// Device-side scalars of one encode (one cache line).
struct StreamInfo {
  unsigned long long total_bits;    // entropy-coded bits before padding
  unsigned long long stuffed_bytes; // number of 0xFF bytes that got a 0x00 appended
  unsigned long long out_size;      // header + scan + EOI
  unsigned long long pad;
};

// F1: colour convert + fDCT (+ quantise) for the MCU rectangle [mx0,mx1) x [my0,my1).
//   raw = true : coef receives the unquantised x16 coefficients, natural order
//   raw = false: coef receives quantised values in zig-zag order, nzmask the non-zero PAIR bitmaps
//                (bit p set <=> zig-zag positions 2p, 2p+1 are not both zero; bit 0 includes the DC)
// generic path: any stride / alignment / pixel format, edge replication (encoders.cc:157-253)
void LaunchF1Generic(const ImageDesc& img, int mx0, int my0, int mx1, int my1, bool raw,
                     const QuantTabs& qt, int16_t* coef, uint32_t* nzmask, cudaStream_t s);
// fast path: 4:2:0 / 4:4:4 / 4:0:0, RGB24, full MCUs only, rows 16-byte aligned (stride % 16 == 0,
// base % 16 == 0); rows [my0,my1) x all columns [0, mx_full).  Bulk-async (TMA) staged tiles.
bool F1FastEligible(const ImageDesc& img);
void LaunchF1Fast(const ImageDesc& img, int mx_full, int my0, int my1, bool raw, const QuantTabs& qt,
                  int16_t* coef, uint32_t* nzmask, cudaStream_t s);

// Q1: quantise stored raw coefficients in place (natural int16 -> zig-zag int16) + nzmask
void LaunchRequantize(int16_t* coef, uint32_t* nzmask, size_t nb_blocks, int mcu_blocks,
                      int luma_blocks, const QuantTabs& qt, cudaStream_t s);
// H1: histogram of |coef| >> 2 per matrix and position: counts[2][64][129] (int32, pre-zeroed)
void LaunchHistogram(const int16_t* raw_coef, size_t nb_blocks, int mcu_blocks, int luma_blocks,
                     int32_t* counts, cudaStream_t s);
// T1: trellis quantisation of raw coefficients in place (quantize.cc:388-457) + nzmask.
// quant = the two 8-bit matrices (device, natural order); rate from ac code lengths in tabs.
void LaunchTrellis(int16_t* coef, uint32_t* nzmask, size_t nb_blocks, int mcu_blocks,
                   int luma_blocks, const QuantTabs& qt, const uint8_t* quant, const CodeTabs* tabs,
                   cudaStream_t s);

// S1: symbol statistics: freq[2][272] (slot < 256 AC symbol, 256+n DC size), pre-zeroed
void LaunchSymbolStats(const int16_t* zz, const uint32_t* nzmask, size_t nb_blocks, int mcu_blocks,
                       int luma_blocks, uint32_t* freq, cudaStream_t s);
// E1: bits per block + per-tile sums
void LaunchBlockBits(const int16_t* zz, const uint32_t* nzmask, size_t nb_blocks, int mcu_blocks,
                     int luma_blocks, const CodeTabs* tabs, uint32_t* block_bits,
                     uint32_t* tile_sums, cudaStream_t s);
// E2: exclusive scan of the tile sums (u32 -> u64) and the grand total into info->total_bits
void LaunchScanTiles(const uint32_t* tile_sums, size_t nb_tiles, unsigned long long* tile_offsets,
                     StreamInfo* info, cudaStream_t s);
// E3: pack the code words at their bit offsets into the (zeroed) word stream
void LaunchPack(const int16_t* zz, const uint32_t* nzmask, size_t nb_blocks, int mcu_blocks,
                int luma_blocks, const CodeTabs* tabs, const uint32_t* block_bits,
                const unsigned long long* tile_offsets, uint32_t* stream, cudaStream_t s);
// E4: 0xFF stuffing.  Count per tile, scan, then scatter to out + header_len, append padding and
// EOI, zero the consumed stream words, write info->out_size.
void LaunchStuff(uint32_t* stream, size_t max_stream_words, uint32_t* ff_tile_sums,
                 unsigned long long* ff_tile_offsets, StreamInfo* info, uint8_t* out,
                 size_t header_len, cudaStream_t s);

}  // namespace sjb
