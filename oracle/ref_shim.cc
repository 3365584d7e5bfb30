// TEST INFRASTRUCTURE ONLY.  Thin extern "C" door onto the UNMODIFIED reference's C++ API
// (sjpeg::Encode with an EncoderParam, /root/reference/src/sjpeg.h:187-292) so that Python tests
// can drive parameters the plain-C SjpegEncode() does not expose (custom matrices, bias, deltas,
// trellis flag).  Compiled only where /root/reference exists; lives in oracle/_ref/.
#include <cstring>
#include <memory>
#include <string>
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

// planar / semi-planar entry points of the reference (sjpeg.h:313-349) behind one C door.
// kind: 0 = EncodeYUV420, 1 = EncodeYUV444, 2 = EncodeNV12, 3 = EncodeNV21, 4 = EncodeGray
extern "C" size_t ref_encode_planar(int kind, const uint8_t* y, int y_stride, const uint8_t* u, int u_stride,
                                    const uint8_t* v, int v_stride, int w, int h, float quality, int huffman,
                                    int adaptive, int trellis, uint8_t** out) {
  sjpeg::EncoderParam p(quality);
  p.Huffman_compress = huffman != 0;
  p.adaptive_quantization = adaptive != 0;
  p.use_trellis = trellis != 0;
  std::string s;
  std::shared_ptr<sjpeg::ByteSink> sink = sjpeg::MakeByteSink(&s);
  bool ok = false;
  switch (kind) {
    case 0: ok = sjpeg::EncodeYUV420(y, y_stride, u, u_stride, v, v_stride, w, h, p, sink.get()); break;
    case 1: ok = sjpeg::EncodeYUV444(y, y_stride, u, u_stride, v, v_stride, w, h, p, sink.get()); break;
    case 2: ok = sjpeg::EncodeNV12(y, y_stride, u, u_stride, w, h, p, sink.get()); break;
    case 3: ok = sjpeg::EncodeNV21(y, y_stride, u, u_stride, w, h, p, sink.get()); break;
    case 4: ok = sjpeg::EncodeGray(y, w, h, y_stride, p, sink.get()); break;
  }
  if (!ok || s.empty()) return 0;
  *out = new uint8_t[s.size()];
  memcpy(*out, s.data(), s.size());
  return s.size();
}

// metadata (sjpeg.h:238-259): the four payloads as (pointer, length) pairs
extern "C" size_t ref_encode_meta(const uint8_t* rgb, int w, int h, int stride, int yuv_mode, float quality,
                                  const char* exif, size_t exif_len, const char* iccp, size_t iccp_len,
                                  const char* xmp, size_t xmp_len, const char* app, size_t app_len, int xmp_split,
                                  uint8_t** out) {
  sjpeg::EncoderParam p(quality);
  p.yuv_mode = static_cast<SjpegYUVMode>(yuv_mode);
  if (exif_len) p.exif.assign(exif, exif_len);
  if (iccp_len) p.iccp.assign(iccp, iccp_len);
  if (xmp_len) p.xmp.assign(xmp, xmp_len);
  if (app_len) p.app_markers.assign(app, app_len);
  p.xmp_split_point = static_cast<uint16_t>(xmp_split);
  return sjpeg::Encode(rgb, w, h, stride, p, out);
}

// target size / PSNR search (sjpeg.h:214-226, dichotomy.cc)
extern "C" size_t ref_encode_search(const uint8_t* rgb, int w, int h, int stride, int yuv_mode, float quality,
                                    int target_mode, float target_value, int passes, float tolerance, int huffman,
                                    int adaptive, int trellis, float* final_q, float* final_value, uint8_t** out) {
  sjpeg::EncoderParam p(quality);
  p.yuv_mode = static_cast<SjpegYUVMode>(yuv_mode);
  p.Huffman_compress = huffman != 0;
  p.adaptive_quantization = adaptive != 0;
  p.use_trellis = trellis != 0;
  p.target_mode = static_cast<sjpeg::EncoderParam::TargetMode>(target_mode);
  p.target_value = target_value;
  p.passes = passes;
  p.tolerance = tolerance;
  sjpeg::SearchHook hook;
  p.search_hook = &hook;
  const size_t n = sjpeg::Encode(rgb, w, h, stride, p, out);
  if (final_q) *final_q = hook.q;
  if (final_value) *final_value = hook.value;
  return n;
}

#ifdef SJPEG_REF_INTERNALS   // only when linked with the reference's own objects (oracle/Makefile)
// sharp RGB->YUV420 pre-pass and the riskiness score table: internal symbols of the reference
// (declared in /root/reference/src/sjpegi.h:89,118), reached here so that tests can compare planes
// and feed the reference's own generated table (score_7.cc) to the code under test.
namespace sjpeg {
extern const uint8_t kSharpnessScore[];
void ApplySharpYUVConversion(const uint8_t* const rgb, int W, int H, int stride, uint8_t* y_plane,
                             uint8_t* u_plane, uint8_t* v_plane);
}
extern "C" void ref_sharp_yuv(const uint8_t* rgb, int w, int h, int stride, uint8_t* y, uint8_t* u, uint8_t* v) {
  sjpeg::ApplySharpYUVConversion(rgb, w, h, stride, y, u, v);
}
extern "C" const uint8_t* ref_score_table(size_t* size) {
  if (size) *size = 343u * 343u;
  return sjpeg::kSharpnessScore;
}
#endif  // SJPEG_REF_INTERNALS
