// TEST INFRASTRUCTURE ONLY.  Thin extern "C" door onto the UNMODIFIED reference's C++ API
// (sjpeg::Encode with an EncoderParam, /root/reference/src/sjpeg.h:187-292) so that Python tests
// can drive parameters the plain-C SjpegEncode() does not expose (custom matrices, bias, deltas,
// trellis flag).  Compiled only where /root/reference exists; lives in oracle/_ref/.
#include <cstring>
#include "sjpeg.h"

extern "C" size_t ref_encode_param(const uint8_t* rgb, int w, int h, int stride, int yuv_mode,
                                   const uint8_t* quant /* [2][64] or null */, float quality,
                                   int huffman, int adaptive, int trellis, int bias, int qd_luma,
                                   int qd_chroma, uint8_t** out) {
  sjpeg::EncoderParam p(quality);
  if (quant != nullptr) {
    uint8_t m[2][64];
    memcpy(m, quant, sizeof(m));
    p.SetQuantization(m, 100.f);
  }
  p.yuv_mode = static_cast<SjpegYUVMode>(yuv_mode);
  p.Huffman_compress = huffman != 0;
  p.adaptive_quantization = adaptive != 0;
  p.use_trellis = trellis != 0;
  if (bias >= 0) p.quantization_bias = bias;
  if (qd_luma >= 0) p.qdelta_max_luma = qd_luma;
  if (qd_chroma >= 0) p.qdelta_max_chroma = qd_chroma;
  return sjpeg::Encode(rgb, w, h, stride, p, out);
}
