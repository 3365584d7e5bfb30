// host_codec.h -- the small, order-sensitive host-side pieces of the encode path: quantiser
// tables, Huffman table construction, adaptive-quantisation analysis and the JPEG header bytes.
// They are microseconds of scalar (partly floating-point) work per image and stay on the host by
// design; the per-pixel / per-block work is in kernels.cu.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "block_ops.cuh"

namespace sjb {

enum { kHistoBins = 128, kHistoStride = 129 };   // sjpegi.h:268-283 (one spare bin per row)

// /root/reference/src/quantize.cc:77-96 : JPEG quality -> luma/chroma 8-bit matrices
void QualityToMatrices(float quality, uint8_t out[2][64]);
// quantize.cc:88-96 with an explicit percentage
void ScaleMatrix(const uint8_t in[64], float percent, uint8_t out[64]);
// quantize.cc:98-104
void MinMatrixWithTolerance(const uint8_t in[64], int tolerance, uint8_t out[64]);
// K.1 tables (quantize.cc:57-75)
const uint8_t* DefaultMatrix(int idx);

// quantize.cc:123-148 : clamps quant[] to min_quant[] and derives the device constants.
// Returns false if an entry cannot be expressed in the fused (threshold-free) form.
bool FinalizeQuantizer(uint8_t quant[64], const uint8_t min_quant[64], int q_bias, QuantTab* out);

struct HuffSpec {        // a DHT payload: counts per length + symbols by increasing length
  uint8_t bits[16];
  uint8_t syms[256];
  int nb_syms;
};
void DefaultHuffSpec(bool ac, int chroma, HuffSpec* out);        // T.81 K.3 (entropy.cc:31-82)
void CodesFromSpec(const HuffSpec& spec, uint32_t* table);        // entropy.cc:98-112
void OptimalHuffSpec(const uint32_t* freq, int size, HuffSpec* out);   // entropy.cc:254-430

// histogram.cc:126-315 : counts = int32[2][64][129]; rewrites quant[][] in place
void AnalyseHistograms(const int32_t* counts, int nb_comps, uint8_t quant[2][64],
                       const uint8_t min_quant[2][64], int qdelta_max_luma, int qdelta_max_chroma);

struct FrameGeometry {
  int yuv_mode, width, height;
  int nb_comps, mcu_blocks, luma_blocks, mcu_size;   // mcu_size = 16 (420) or 8
  int mcus_x, mcus_y;
  size_t nb_mcus() const { return (size_t)mcus_x * mcus_y; }
  size_t nb_blocks() const { return nb_mcus() * mcu_blocks; }
};
bool MakeGeometry(int yuv_mode, int width, int height, FrameGeometry* g);   // encoders.cc:32-88

// headers.cc:48-61,182-258 : SOI+APP0, DQT, SOF0, DHT x(2|4), SOS
void AppendHeaders(const FrameGeometry& g, const uint8_t quant[2][64], const HuffSpec spec[4],
                   std::vector<uint8_t>* out);

// Gamma tables of the sharp RGB->YUV conversion (yuv_convert.cc:102-151): g2l[1024] gamma ->
// linear with 14 fractional bits, l2g[34] interpolation nodes of the inverse.  Double-precision
// pow() on the host, exactly as the reference builds them at first use.
void MakeSharpGammaTables(uint32_t g2l[1024], uint32_t l2g[34]);

// Final decision of the riskiness analyser (jpeg_tools.cc:212-235) from the three sums the device
// pass produces.  Returns the SjpegYUVMode value (420 = 1, SHARP = 2, 444 = 3, 400 = 4).
int RiskinessDecision(unsigned long long score_sum, unsigned long long score_num, unsigned long long gray_num,
                      int width, int height, float* risk);

}  // namespace sjb
