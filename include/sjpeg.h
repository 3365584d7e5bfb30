/*
 * sjpeg.h -- source-compatible mirror of the public interface of webmproject/sjpeg
 * (/root/reference/src/sjpeg.h) for the B200 build.  Same names, argument meaning, struct
 * layout and error behaviour, so that code written against the reference header compiles and
 * links against libsjpeg_b200.so unchanged.  The encode path runs on the GPU through the C ABI
 * of sjpeg_b200.h; see INTEGRATION.md for what is and is not covered:
 *
 *   on the accelerated path : SjpegEncode, SjpegCompress, sjpeg::Encode (3 overloads),
 *                             sjpeg::EncodeRGBA / EncodeBGRA, EncodeGray / EncodeYUV420 / EncodeYUV444 /
 *                             EncodeNV12 / EncodeNV21 (planar and semi-planar input), yuv modes
 *                             420 / 444 / 400 / SHARP / AUTO, compression methods 0..8, custom matrices /
 *                             bias / deltas / trellis, the multi-pass target-size / PSNR search
 *                             (EncoderParam::passes > 1, SearchHook), SjpegRiskiness
 *   host utilities          : SjpegVersion, SjpegFreeBuffer, SjpegQuantMatrix, SjpegEstimateQuality,
 *                             SjpegDimensions, SjpegFindQuantizer, MakeByteSink, metadata segments
 *                             (EXIF / ICC / XMP / extended XMP / app markers)
 *   one precondition        : SJPEG_YUV_AUTO, SjpegCompress and SjpegRiskiness look pixels up in the
 *                             reference's generated score table (src/score_7.cc); sjpeg_b200.h says
 *                             where the library finds it.  Without it those calls FAIL (0 / false /
 *                             SJPEG_YUV_AUTO) rather than encode in a mode the reference would not pick.
 */
#ifndef SJPEG_JPEG_H_
#define SJPEG_JPEG_H_

#include <inttypes.h>
#include <stddef.h>

#include <memory>
#include <string>
#include <vector>

#define SJPEG_VERSION 0x000101 /* 0.1.1, reference sjpeg.h:28 */

#if defined(__cplusplus) || defined(c_plusplus)
extern "C" {
#endif

/* reference sjpeg.h:34 */
uint32_t SjpegVersion();

/* reference sjpeg.h:45 / api.cc:53-57 : method 4, automatic colour mode.
 * *out_data is allocated with new[]; release with delete[] or SjpegFreeBuffer().  0 on error. */
size_t SjpegCompress(const uint8_t* rgb, int width, int height, float quality, uint8_t** out_data);

/* reference sjpeg.h:54-60 */
typedef enum {
  SJPEG_YUV_AUTO = 0,
  SJPEG_YUV_420,
  SJPEG_YUV_SHARP,
  SJPEG_YUV_444,
  SJPEG_YUV_400
} SjpegYUVMode;

/* reference sjpeg.h:104 / api.cc:32-49.  stride in bytes, |stride| >= 3*width, may be negative.
 * compression_method is clamped to 0..8 (table in the reference header, sjpeg.h:75-101). */
size_t SjpegEncode(const uint8_t* rgb, int width, int height, int stride, uint8_t** out_data,
                   float quality, int compression_method, SjpegYUVMode yuv_mode);

/* reference sjpeg.h:113 */
void SjpegFreeBuffer(const uint8_t* buffer);

/* reference sjpeg.h:122-149 : bitstream / analysis helpers */
bool SjpegDimensions(const uint8_t* data, size_t size, int* width, int* height, int* is_yuv420);
int SjpegFindQuantizer(const uint8_t* data, size_t size, uint8_t quant[2][64]);
float SjpegEstimateQuality(const uint8_t matrix[64], bool for_chroma);
void SjpegQuantMatrix(float quality, bool for_chroma, uint8_t matrix[64]);
SjpegYUVMode SjpegRiskiness(const uint8_t* rgb, int width, int height, int stride, float* risk);

#if defined(__cplusplus) || defined(c_plusplus)
} /* extern "C" */
#endif

/* std::string flavours, reference sjpeg.h:158-164 */
bool SjpegCompress(const uint8_t* rgb, int width, int height, float quality, std::string* output);
bool SjpegDimensions(const std::string& jpeg_data, int* width, int* height, int* is_yuv420);
int SjpegFindQuantizer(const std::string& jpeg_data, uint8_t quant[2][64]);

namespace sjpeg {

struct Encoder;
struct SearchHook;
struct ByteSink;
struct MemoryManager;

/* reference sjpeg.h:187-275.  Field order and types are part of the ABI. */
struct EncoderParam {
  EncoderParam();
  explicit EncoderParam(float quality_factor);

  void SetQuality(float quality_factor);
  void SetQuantization(const uint8_t m[2][64], float reduction = 100.f);
  const uint8_t* GetQuantMatrix(int idx) const { return quant_[idx]; }
  void SetLimitQuantization(bool limit_quantization = true, int tolerance = 0);
  void SetMinQuantization(const uint8_t m[2][64], int min_quant_tolerance = 0);

  SjpegYUVMode yuv_mode;
  bool Huffman_compress;
  bool adaptive_quantization;
  bool adaptive_bias;
  bool use_trellis;

  typedef enum { TARGET_NONE = 0, TARGET_SIZE = 1, TARGET_PSNR = 2 } TargetMode;
  TargetMode target_mode;
  float target_value;
  int passes;
  float tolerance;
  float qmin, qmax;

  int quantization_bias;
  int qdelta_max_luma;
  int qdelta_max_chroma;

  sjpeg::SearchHook* search_hook;

  std::string exif;
  std::string iccp;
  std::string app_markers;
  std::string xmp;
  uint16_t xmp_split_point = 0u;
  void ResetMetadata();

  sjpeg::MemoryManager* memory;

 protected:
  uint8_t quant_[2][64];
  uint8_t min_quant_[2][64];
  bool use_min_quant_;
  int min_quant_tolerance_;

 protected:
  void Init(float quality_factor);
  friend struct sjpeg::Encoder;
};

/* reference sjpeg.h:280-292 */
bool Encode(const uint8_t* rgb, int width, int height, int stride, const EncoderParam& param,
            std::string* output);
size_t Encode(const uint8_t* rgb, int width, int height, int stride, const EncoderParam& param,
              uint8_t** out_data);
bool Encode(const uint8_t* rgb, int width, int height, int stride, const EncoderParam& param,
            sjpeg::ByteSink* sink);

/* reference sjpeg.h:300-311 : 4 bytes per pixel, alpha ignored */
bool EncodeBGRA(const uint8_t* bgra, int width, int height, int stride, const EncoderParam& param,
                sjpeg::ByteSink* sink);
bool EncodeBGRA(const uint8_t* bgra, int width, int height, int stride, const EncoderParam& param,
                std::string* output);
bool EncodeRGBA(const uint8_t* rgba, int width, int height, int stride, const EncoderParam& param,
                sjpeg::ByteSink* sink);
bool EncodeRGBA(const uint8_t* rgba, int width, int height, int stride, const EncoderParam& param,
                std::string* output);

/* reference sjpeg.h:316-349 : planar / semi-planar input (encoders.cc:256-507), on the device through sjb_encode_planar */
bool EncodeGray(const uint8_t* gray, int width, int height, int stride, const EncoderParam& param,
                sjpeg::ByteSink* sink);
bool EncodeGray(const uint8_t* gray, int width, int height, int stride, const EncoderParam& param,
                std::string* output);
bool EncodeNV21(const uint8_t* y, int y_stride, const uint8_t* vu, int vu_stride, int width,
                int height, const EncoderParam& param, sjpeg::ByteSink* output);
bool EncodeNV12(const uint8_t* y, int y_stride, const uint8_t* uv, int uv_stride, int width,
                int height, const EncoderParam& param, sjpeg::ByteSink* output);
bool EncodeYUV444(const uint8_t* Y, int Y_stride, const uint8_t* U, int U_stride, const uint8_t* V,
                  int V_stride, int width, int height, const EncoderParam& param,
                  sjpeg::ByteSink* output);
bool EncodeYUV420(const uint8_t* Y, int Y_stride, const uint8_t* U, int U_stride, const uint8_t* V,
                  int V_stride, int width, int height, const EncoderParam& param,
                  sjpeg::ByteSink* output);

/* reference sjpeg.h:355-373 */
struct SearchHook {
  float q;
  float qmin, qmax;
  float target;
  float tolerance;
  bool for_size;
  float value;
  int pass;
  virtual bool Setup(const EncoderParam& param);
  virtual void NextMatrix(int idx, uint8_t dst[64]);
  virtual bool Update(float result);
  virtual ~SearchHook() {}
};

/* reference sjpeg.h:392-398.  Protocol: Reset() first; Commit(used, extra, &buf) hands back the
 * bytes written since the previous Commit and asks for 'extra' more; the last call is
 * Commit(used, 0, &buf) followed by Finalize(); Reset() again on failure. */
struct ByteSink {
 public:
  virtual ~ByteSink() {}
  virtual bool Commit(size_t used_size, size_t extra_size, uint8_t** data) = 0;
  virtual bool Finalize() = 0;
  virtual void Reset() = 0;
};

/* reference sjpeg.h:401-405 */
std::shared_ptr<ByteSink> MakeByteSink(std::string* output);
template <typename T> std::shared_ptr<ByteSink> MakeByteSink(std::vector<T>* output);
template <> std::shared_ptr<ByteSink> MakeByteSink(std::vector<uint8_t>* output);

/* reference sjpeg.h:410-415 : host allocations of the codec go through this */
struct MemoryManager {
 public:
  virtual ~MemoryManager() {}
  virtual void* Alloc(size_t size) = 0;
  virtual void Free(void* const ptr) = 0;
};

}  // namespace sjpeg

#endif /* SJPEG_JPEG_H_ */
