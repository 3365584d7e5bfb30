// sjpeg_api.cc -- the reference's public entry points (include/sjpeg.h, mirroring
// /root/reference/src/sjpeg.h and api.cc) implemented on top of the C ABI of the GPU path.
// Argument checks, ownership (new[] buffers), sink protocol and error returns follow
// /root/reference/src/api.cc:32-67,145-201; the per-MCU work is in kernels.cu.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>

#include <algorithm>
#include <new>
#include <string>

#include "../../include/sjpeg.h"
#include "../../include/sjpeg_b200.h"
#include "host_codec.h"

namespace {

// One GPU context per host thread: the reference is re-entrant (tests/unit_test.cc:114-131 runs
// 8 encodes concurrently), so no state is shared between calling threads.
struct ThreadContext {
  sjb_context* ctx = nullptr;
  ~ThreadContext() { sjb_context_destroy(ctx); }
  sjb_context* get() {
    if (ctx == nullptr) {
      const char* dev = getenv("SJPEG_B200_DEVICE");
      if (sjb_context_create(dev ? atoi(dev) : 0, &ctx) != SJB_OK) ctx = nullptr;
    }
    return ctx;
  }
};
thread_local ThreadContext tls_context;

struct DefaultMemory : public sjpeg::MemoryManager {
  void* Alloc(size_t size) override { return malloc(size); }
  void Free(void* const ptr) override { free(ptr); }
} default_memory;

int BytesPerPixel(int fmt) { return fmt == SJB_PIX_RGB ? 3 : 4; }

// api.cc:183-192 + encoders.cc:546-568 + api.cc:145-181, ending in a sink
bool CommitWithMetadata(const uint8_t* jpeg, size_t size, const std::string& meta, sjpeg::ByteSink* sink);

struct SearchState;
bool ArmSearch(sjb_context* ctx, const sjpeg::EncoderParam& param, sjpeg::SearchHook* fallback, SearchState* st,
               sjb_search* search);
void FinishSearch(const sjpeg::EncoderParam& param, SearchState* st, const sjb_search& search, bool ok);

bool EncodeToSink(const uint8_t* pix, int width, int height, int stride, int fmt, const sjb_params& params,
                  sjpeg::ByteSink* sink, sjpeg::MemoryManager* memory, const std::string& meta = std::string(),
                  const sjpeg::EncoderParam* full = nullptr);
// Encoder::InitFromParam, api.cc:145-181
// true if the riskiness table is available; otherwise says so on stderr (once per process)
bool HaveScoreTableOrComplain() {
  if (sjb_has_score_table()) return true;
  static std::atomic<bool> said(false);
  if (!said.exchange(true)) {
    fprintf(stderr, "sjpeg_b200: SJPEG_YUV_AUTO / SjpegCompress / SjpegRiskiness need the reference's riskiness score "
                    "table and none was found (sjb_set_score_table, SJPEG_B200_SCORE_TABLE, a co-loaded libsjpeg, or "
                    "sjpeg_score_table.bin next to the library): the call fails instead of picking another mode\n");
  }
  return false;
}
bool ParamsFromEncoderParam(const sjpeg::EncoderParam& param, const uint8_t quant[2][64],
                            const uint8_t min_quant[2][64], bool use_min_quant, int tolerance, int fmt,
                            sjb_params* out) {
  memset(out, 0, sizeof(*out));
  SjpegYUVMode mode = param.yuv_mode;
  // The riskiness analyser (jpeg_tools.cc:177-236) runs on the device but needs the reference's
  // generated score table (include/sjpeg_b200.h: sjb_set_score_table and the places it is looked
  // for).  Without it AUTO cannot be honoured and the call FAILS -- encoding in some other mode
  // would be a silently different file.
  if (mode == SJPEG_YUV_AUTO && !HaveScoreTableOrComplain()) return false;
  if (mode != SJPEG_YUV_AUTO && mode != SJPEG_YUV_420 && mode != SJPEG_YUV_SHARP && mode != SJPEG_YUV_444 &&
      mode != SJPEG_YUV_400) {
    return false;
  }
  out->yuv_mode = mode;
  out->pix_fmt = fmt;
  for (int i = 0; i < 2; ++i) {
    sjb::ScaleMatrix(quant[i], 100.f, out->quant[i]);   // enc.cc:106-109
    if (use_min_quant) sjb::MinMatrixWithTolerance(min_quant[i], tolerance, out->min_quant[i]);
    else memset(out->min_quant[i], 1, 64);
  }
  int method = param.Huffman_compress ? 1 : 0;
  if (param.adaptive_quantization) method += 3;
  if (param.use_trellis) method = (method == 4) ? 7 : (method == 6) ? 8 : method;
  out->method = method;
  out->q_bias = param.quantization_bias;
  out->qdelta_max_luma = param.qdelta_max_luma;
  out->qdelta_max_chroma = param.qdelta_max_chroma;
  if (out->q_bias < 0 || out->q_bias > 255) return false;
  return true;
}

// ---------------------------------------------------------------------------------------------
// Metadata segments (headers.cc:63-180): host-side bytes inserted between APP0 and DQT.
// ---------------------------------------------------------------------------------------------
// RFC 1321 MD5, upper-case hex digest (the GUID of extended XMP, headers.cc:127-129)
std::string Md5HexUpper(const std::string& data) {
  static const int kShift[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9,
                                 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
  uint32_t K[64];
  for (int i = 0; i < 64; ++i) K[i] = static_cast<uint32_t>(floor(fabs(sin(i + 1.0)) * 4294967296.0));
  uint32_t h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
  std::string msg = data;
  const uint64_t bit_len = static_cast<uint64_t>(data.size()) * 8;
  msg.push_back(static_cast<char>(0x80));
  while (msg.size() % 64 != 56) msg.push_back(0);
  for (int i = 0; i < 8; ++i) msg.push_back(static_cast<char>((bit_len >> (8 * i)) & 0xff));
  for (size_t off = 0; off < msg.size(); off += 64) {
    uint32_t w[16];
    for (int i = 0; i < 16; ++i) {
      const uint8_t* b = reinterpret_cast<const uint8_t*>(&msg[off + 4 * i]);
      w[i] = b[0] | (b[1] << 8) | (b[2] << 16) | (static_cast<uint32_t>(b[3]) << 24);
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
    for (int i = 0; i < 64; ++i) {
      uint32_t f;
      int g;
      if (i < 16) { f = (b & c) | (~b & d); g = i; }
      else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
      else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
      else { f = c ^ (b | ~d); g = (7 * i) & 15; }
      const uint32_t t = a + f + K[i] + w[g];
      a = d; d = c; c = b;
      b += (t << kShift[i]) | (t >> (32 - kShift[i]));
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
  }
  static const char kHex[] = "0123456789ABCDEF";
  std::string out;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) {
      const uint32_t byte = (h[i] >> (8 * j)) & 0xff;
      out.push_back(kHex[byte >> 4]);
      out.push_back(kHex[byte & 15]);
    }
  }
  return out;
}

void Put16(std::string* o, size_t v) {
  o->push_back(static_cast<char>((v >> 8) & 0xff));
  o->push_back(static_cast<char>(v & 0xff));
}
void Put32(std::string* o, size_t v) {
  Put16(o, (v >> 16) & 0xffff);
  Put16(o, v & 0xffff);
}

bool AppendXmp(const std::string& data, std::string* o);

// headers.cc:115-160
bool AppendXmpExtended(const std::string& data, size_t split_point, std::string* o) {
  const size_t kMainSize = 65503;
  if (data.size() < kMainSize) return true;
  if (data.size() > (1u << 31)) return false;
  size_t split = (split_point == 0) ? kMainSize : split_point;
  split = std::min(split, data.size());
  const size_t note = data.find("xmpNote:HasExtendedXMP=\"");
  if (note == std::string::npos) return false;
  if (note + 24 + 32 + 1 > split) return false;
  if (data[note + 24 + 32] != '"') return false;
  std::string main_data(data, 0, split), ext(data, split);
  main_data.replace(note + 24, 32, Md5HexUpper(ext));
  const std::string guid = main_data.substr(note + 24, 32);
  if (!AppendXmp(main_data, o)) return false;
  static const char kExt[] = "http://ns.adobe.com/xmp/extension/";
  const size_t kBuf = 65458, kHeader = sizeof(kExt) + 40;
  const size_t chunks = ext.size() / kBuf + 1;
  size_t pos = 0;
  for (size_t c = 0; c < chunks; ++c) {
    const size_t n = std::min(kBuf, ext.size() - pos);
    Put16(o, 0xffe1);
    Put16(o, 2 + kHeader + n);
    o->append(kExt, sizeof(kExt));
    o->append(guid);
    Put32(o, ext.size());
    Put32(o, pos);
    o->append(ext, pos, n);
    pos += n;
  }
  return true;
}

thread_local size_t tls_xmp_split = 0;

// headers.cc:162-180
bool AppendXmp(const std::string& data, std::string* o) {
  if (data.empty()) return true;
  static const char kXmp[] = "http://ns.adobe.com/xap/1.0/";
  const size_t size = 2 + data.size() + sizeof(kXmp);
  if (size <= 0xffff) {
    Put16(o, 0xffe1);
    Put16(o, size);
    o->append(kXmp, sizeof(kXmp));
    o->append(data);
    return true;
  }
  return AppendXmpExtended(data, tls_xmp_split, o);
}

// enc.cc:415-421: app markers as is, then EXIF, ICC, XMP
bool BuildMetadata(const sjpeg::EncoderParam& p, std::string* o) {
  o->clear();
  o->append(p.app_markers);
  if (!p.exif.empty()) {                                   // headers.cc:72-85
    const size_t size = p.exif.size() + 6 + 2;
    if (size > 0xffff) return false;
    Put16(o, 0xffe1);
    Put16(o, size);
    o->append("Exif\0\0", 6);
    o->append(p.exif);
  }
  if (!p.iccp.empty()) {                                   // headers.cc:87-113
    const size_t kMaxChunk = 0xffff - 12 - 4;
    const size_t chunks = (p.iccp.size() + kMaxChunk - 1) / kMaxChunk;
    if (chunks >= 256) return false;
    size_t pos = 0;
    for (size_t seq = 1; pos < p.iccp.size(); ++seq) {
      const size_t n = std::min(kMaxChunk, p.iccp.size() - pos);
      Put16(o, 0xffe2);
      Put16(o, n + 12 + 4);
      o->append("ICC_PROFILE\0", 12);
      o->push_back(static_cast<char>(seq & 0xff));
      o->push_back(static_cast<char>(chunks & 0xff));
      o->append(p.iccp, pos, n);
      pos += n;
    }
  }
  tls_xmp_split = p.xmp_split_point;
  return AppendXmp(p.xmp, o);
}

// Hands a finished file (SOI+APP0 first, 20 bytes: headers.cc:48-61) to the sink, with the metadata
// segments inserted right after APP0.
bool CommitWithMetadata(const uint8_t* jpeg, size_t size, const std::string& meta, sjpeg::ByteSink* sink) {
  const size_t kApp0 = 20;
  if (size < kApp0) return false;
  uint8_t* dst = nullptr;
  const size_t total = size + meta.size();
  if (!sink->Commit(0, total, &dst) || dst == nullptr) return false;
  memcpy(dst, jpeg, kApp0);
  if (!meta.empty()) memcpy(dst + kApp0, meta.data(), meta.size());
  memcpy(dst + kApp0 + meta.size(), jpeg + kApp0, size - kApp0);
  return sink->Commit(total, 0, &dst) && sink->Finalize();
}

// ---------------------------------------------------------------------------------------------
// Multi-pass search glue (api.cc:169-176, dichotomy.cc:113-205): SearchHook <-> sjb_search
// ---------------------------------------------------------------------------------------------
struct SearchState {
  sjpeg::SearchHook* hook;
  float q_of_pass[32];
};
void SearchBeginPass(void* user, int pass) {
  SearchState* st = static_cast<SearchState*>(user);
  st->hook->pass = pass;
  st->q_of_pass[pass & 31] = st->hook->q;
}
void SearchNextMatrix(void* user, int idx, uint8_t dst[64]) { static_cast<SearchState*>(user)->hook->NextMatrix(idx, dst); }
int SearchUpdate(void* user, float result) { return static_cast<SearchState*>(user)->hook->Update(result) ? 1 : 0; }

// metadata part of Encoder::HeaderSize() (dichotomy.cc:212-229): an estimate, reproduced as is
size_t MetadataSizeEstimate(const sjpeg::EncoderParam& p) {
  size_t size = p.app_markers.size();
  if (!p.exif.empty()) size += 8 + p.exif.size();
  if (!p.iccp.empty()) {
    const size_t chunk_max = 0xffff - 12 - 4;
    size += ((p.iccp.size() - 1) / chunk_max + 1) * (12 + 4 + 2) + p.iccp.size();
  }
  if (!p.xmp.empty()) {
    size += 2 + 2 + 29 + p.xmp.size();
    if (p.xmp.size() > 65533) size += (p.xmp.size() / 65458 + 1) * 40;
  }
  return size;
}

// Arms the context's search when param.passes > 1.  Returns false if the hook refuses.
bool ArmSearch(sjb_context* ctx, const sjpeg::EncoderParam& param, sjpeg::SearchHook* fallback, SearchState* st,
               sjb_search* search) {
  const int passes = param.passes < 1 ? 1 : param.passes > 20 ? 20 : param.passes;
  if (passes <= 1) return true;
  st->hook = param.search_hook ? param.search_hook : fallback;
  if (!st->hook->Setup(param)) return false;
  memset(search, 0, sizeof(*search));
  search->passes = passes;
  search->for_size = st->hook->for_size ? 1 : 0;
  search->target = st->hook->target;
  search->header_extra_bytes = MetadataSizeEstimate(param);
  search->user = st;
  search->begin_pass = SearchBeginPass;
  search->next_matrix = SearchNextMatrix;
  search->update = SearchUpdate;
  return sjb_context_set_search(ctx, search) == SJB_OK;
}
void FinishSearch(const sjpeg::EncoderParam& param, SearchState* st, const sjb_search& search, bool ok) {
  const int passes = param.passes < 1 ? 1 : param.passes > 20 ? 20 : param.passes;
  if (passes <= 1 || !ok) return;
  st->hook->q = st->q_of_pass[search.best_pass & 31];      // dichotomy.cc:183-185
  st->hook->value = search.best_result;
}

bool EncodeToSink(const uint8_t* pix, int width, int height, int stride, int fmt, const sjb_params& params,
                  sjpeg::ByteSink* sink, sjpeg::MemoryManager* memory, const std::string& meta,
                  const sjpeg::EncoderParam* full) {
  sink->Reset();                                        // enc.cc:90
  sjb_context* ctx = tls_context.get();
  if (ctx == nullptr) return false;
  if (memory == nullptr) memory = &default_memory;
  SearchState st;
  sjb_search search;
  sjpeg::SearchHook default_hook;
  if (full != nullptr && !ArmSearch(ctx, *full, &default_hook, &st, &search)) return false;
  size_t size = 0;
  const int rc = sjb_encode(ctx, pix, 0, width, height, stride, &params, nullptr, 0, 0, &size);
  sjb_context_set_search(ctx, nullptr);
  if (rc != SJB_ERR_CAPACITY || size == 0) return false;
  if (full != nullptr) FinishSearch(*full, &st, search, true);
  bool ok;
  if (meta.empty() && memory == &default_memory) {
    // no segment to splice in and no caller-supplied allocator to honour: the JPEG goes straight
    // from the device (or the context's pinned head copy) into the sink's own memory
    uint8_t* dst = nullptr;
    ok = sink->Commit(0, size, &dst) && dst != nullptr;
    ok = ok && sjb_fetch_output(ctx, dst, 0, size) == SJB_OK;
    ok = ok && sink->Commit(size, 0, &dst) && sink->Finalize();
  } else {
    // host staging goes through the caller's memory manager (sjpeg.h:410-415)
    uint8_t* staging = static_cast<uint8_t*>(memory->Alloc(size));
    if (staging == nullptr) return false;
    ok = sjb_fetch_output(ctx, staging, 0, size) == SJB_OK;
    ok = ok && CommitWithMetadata(staging, size, meta, sink);
    memory->Free(staging);
  }
  if (!ok) sink->Reset();                               // bit_writer.cc:99-105
  (void)fmt;
  return ok;
}

// sinks (bit_writer.h:51-93)
template <class T>
class ContainerSink : public sjpeg::ByteSink {
 public:
  explicit ContainerSink(T* out) : out_(out), pos_(0) {}
  bool Commit(size_t used, size_t extra, uint8_t** data) override {
    pos_ += used;
    out_->resize(pos_ + extra);
    if (out_->size() != pos_ + extra) return false;
    *data = extra ? reinterpret_cast<uint8_t*>(&(*out_)[pos_]) : nullptr;
    return true;
  }
  bool Finalize() override {
    out_->resize(pos_);
    return true;
  }
  void Reset() override {
    out_->clear();
    pos_ = 0;
  }

 private:
  T* const out_;
  size_t pos_;
};

class NewArraySink : public sjpeg::ByteSink {   // hands over a new[] buffer (api.cc:39,46)
 public:
  NewArraySink() : buf_(nullptr), size_(0), cap_(0) {}
  ~NewArraySink() override { delete[] buf_; }
  bool Commit(size_t used, size_t extra, uint8_t** data) override {
    size_ += used;
    if (size_ + extra > cap_) {
      uint8_t* nb = new (std::nothrow) uint8_t[size_ + extra];
      if (nb == nullptr) return false;
      if (size_) memcpy(nb, buf_, size_);
      delete[] buf_;
      buf_ = nb;
      cap_ = size_ + extra;
    }
    *data = buf_ + size_;
    return true;
  }
  bool Finalize() override { return true; }
  void Reset() override {
    delete[] buf_;
    buf_ = nullptr;
    size_ = cap_ = 0;
  }
  size_t Release(uint8_t** out) {
    *out = buf_;
    const size_t s = size_;
    buf_ = nullptr;
    size_ = cap_ = 0;
    return s;
  }

 private:
  uint8_t* buf_;
  size_t size_, cap_;
};

}  // namespace

// friend of EncoderParam: reads its protected matrices (sjpeg.h:262-274)
namespace sjpeg {
struct Encoder {
  static bool Convert(const EncoderParam& p, int fmt, sjb_params* out) {
    return ParamsFromEncoderParam(p, p.quant_, p.min_quant_, p.use_min_quant_, p.min_quant_tolerance_, fmt, out);
  }
};
}  // namespace sjpeg

// ---------------------------------------------------------------------------------------------
// plain-C entry points (api.cc:32-67)
// ---------------------------------------------------------------------------------------------
extern "C" {

uint32_t SjpegVersion() { return SJPEG_VERSION; }

size_t SjpegEncode(const uint8_t* rgb, int width, int height, int stride, uint8_t** out_data, float quality,
                   int method, SjpegYUVMode yuv_mode) {
  if (rgb == nullptr || out_data == nullptr) return 0;
  if (width <= 0 || height <= 0 || abs(stride) < 3 * width) return 0;
  *out_data = nullptr;
  if (yuv_mode == SJPEG_YUV_AUTO && !HaveScoreTableOrComplain()) return 0;   // see ParamsFromEncoderParam
  if (yuv_mode != SJPEG_YUV_AUTO && yuv_mode != SJPEG_YUV_420 && yuv_mode != SJPEG_YUV_SHARP &&
      yuv_mode != SJPEG_YUV_444 && yuv_mode != SJPEG_YUV_400) {
    return 0;
  }
  sjb_params p;
  sjb_params_default(&p, quality, method, yuv_mode);
  NewArraySink sink;
  if (!EncodeToSink(rgb, width, height, stride, SJB_PIX_RGB, p, &sink, nullptr)) return 0;
  return sink.Release(out_data);
}

size_t SjpegCompress(const uint8_t* rgb, int width, int height, float quality, uint8_t** out_data) {
  return SjpegEncode(rgb, width, height, 3 * width, out_data, quality, 4, SJPEG_YUV_AUTO);
}

void SjpegFreeBuffer(const uint8_t* buffer) { delete[] buffer; }

void SjpegQuantMatrix(float quality, bool for_chroma, uint8_t matrix[64]) {   // jpeg_tools.cc:134-141
  uint8_t m[2][64];
  sjb::QualityToMatrices(quality, m);
  memcpy(matrix, m[for_chroma ? 1 : 0], 64);
}

float SjpegEstimateQuality(const uint8_t matrix[64], bool for_chroma) {       // jpeg_tools.cc:143-168
  int best_q = 0;
  float best = 256.f * 256 * 64 + 1;
  for (int q = 0; q <= 100; ++q) {
    uint8_t m[64];
    SjpegQuantMatrix(static_cast<float>(q), for_chroma, m);
    float score = 0;
    for (int i = 0; i < 64 && score <= best; ++i) {
      const float d = static_cast<float>(m[i]) - matrix[i];
      score += d * d;
    }
    if (score < best) {
      best = score;
      best_q = q;
    }
  }
  return static_cast<float>(best_q);
}

// Walks the marker segments after SOI; returns the offset of the first marker in [lo,hi] or 0.
static size_t FindMarker(const uint8_t* d, size_t size, int lo, int hi, bool stop_at_sos) {
  if (d == nullptr || size < 10 || d[0] != 0xff || d[1] != 0xd8) return 0;
  size_t pos = 2;
  const size_t end = size - 8;
  while (pos < end && d[pos] != 0xff) ++pos;
  while (pos < end) {
    const int marker = d[pos + 1];
    if (marker >= lo && marker <= hi) return pos;
    if (stop_at_sos && marker == 0xda) return 0;
    pos += 2 + ((d[pos + 2] << 8) | d[pos + 3]);
  }
  return 0;
}

bool SjpegDimensions(const uint8_t* data, size_t size, int* width, int* height, int* is_yuv420) {   // jpeg_tools.cc:52-70
  const size_t pos = FindMarker(data, size, 0xc0, 0xc1, false);
  if (pos == 0) return false;
  const uint8_t* s = data + pos;
  const size_t left = size - pos;
  if (left < 11) return false;
  if (height) *height = (s[5] << 8) | s[6];
  if (width) *width = (s[7] << 8) | s[8];
  if (is_yuv420) {
    const size_t nc = s[9];
    if (left < 11 + 3 * nc) { *is_yuv420 = (nc == 3); return false; }
    *is_yuv420 = (nc == 3) && s[11] == 0x22 && s[14] == 0x11 && s[17] == 0x11;
  }
  return true;
}

int SjpegFindQuantizer(const uint8_t* d, size_t size, uint8_t quant[2][64]) {   // jpeg_tools.cc:75-130
  static const uint8_t zz[64] = SJB_ZIGZAG_INIT;
  memset(quant, 0, 128);
  if (d == nullptr || size < 69 || d[0] != 0xff || d[1] != 0xd8) return 0;
  const size_t end = size - 8;
  size_t pos = 2;
  while (pos < end && d[pos] != 0xff) ++pos;
  int seen = 0;
  while (pos < end) {
    const int marker = d[pos + 1];
    const size_t chunk = 2 + ((d[pos + 2] << 8) | d[pos + 3]);
    if (pos + chunk > end || marker == 0xda) break;
    if (marker == 0xdb) {
      size_t i = 4;
      while (i + 1 < chunk) {
        const int pq = d[pos + i] >> 4, tq = d[pos + i] & 15;
        if (pq > 1 || tq > 3) return 0;
        const size_t msize = 64 * pq + 65;
        if (i + msize > chunk) return 0;
        if (tq < 2) {
          for (int j = 0; j < 64; ++j) {
            int v = pq ? ((d[pos + i + 1 + 2 * j] << 8) | d[pos + i + 2 + 2 * j]) : d[pos + i + 1 + j];
            v = v > 255 ? 255 : v < 1 ? 1 : v;
            quant[tq][zz[j]] = static_cast<uint8_t>(v);
          }
        }
        seen |= 1 << tq;
        i += msize;
      }
    }
    pos += chunk;
  }
  return __builtin_popcount(seen & 15);
}

SjpegYUVMode SjpegRiskiness(const uint8_t* rgb, int width, int height, int stride, float* risk) {
  // jpeg_tools.cc:177-236 on the device.  Without the generated score table there is no answer:
  // SJPEG_YUV_AUTO ("undecided", never a value the reference returns) and a message on stderr.
  if (risk) *risk = 0.f;
  if (rgb == nullptr || width <= 0 || height <= 0) return SJPEG_YUV_420;
  if (!HaveScoreTableOrComplain()) return SJPEG_YUV_AUTO;
  sjb_context* ctx = tls_context.get();
  int mode = SJPEG_YUV_420;
  float r = 0.f;
  if (ctx == nullptr || sjb_riskiness(ctx, rgb, 0, width, height, stride, &mode, &r) != SJB_OK) return SJPEG_YUV_420;
  if (risk) *risk = r;
  return static_cast<SjpegYUVMode>(mode);
}

}  // extern "C"

bool SjpegCompress(const uint8_t* rgb, int width, int height, float quality, std::string* output) {
  if (output == nullptr) return false;
  sjpeg::EncoderParam param(quality);     // defaults == method 4 + AUTO (api.cc:83-101)
  return sjpeg::Encode(rgb, width, height, 3 * width, param, output);
}
bool SjpegDimensions(const std::string& d, int* w, int* h, int* is420) {
  return SjpegDimensions(reinterpret_cast<const uint8_t*>(d.data()), d.size(), w, h, is420);
}
int SjpegFindQuantizer(const std::string& d, uint8_t quant[2][64]) {
  return SjpegFindQuantizer(reinterpret_cast<const uint8_t*>(d.data()), d.size(), quant);
}

// ---------------------------------------------------------------------------------------------
// C++ API (api.cc:74-307)
// ---------------------------------------------------------------------------------------------
namespace sjpeg {

EncoderParam::EncoderParam() : search_hook(nullptr), memory(nullptr) { Init(75.f); }
EncoderParam::EncoderParam(float quality_factor) : search_hook(nullptr), memory(nullptr) { Init(quality_factor); }

void EncoderParam::Init(float quality_factor) {          // api.cc:83-101
  Huffman_compress = true;
  adaptive_quantization = true;
  use_trellis = false;
  yuv_mode = SJPEG_YUV_AUTO;
  quantization_bias = 0x78;
  qdelta_max_luma = 12;
  qdelta_max_chroma = 1;
  adaptive_bias = false;
  SetLimitQuantization(false);
  min_quant_tolerance_ = 0;
  SetQuality(quality_factor);
  target_mode = TARGET_NONE;
  target_value = 0;
  passes = 1;
  tolerance = 1.;
  qmin = 0.;
  qmax = 100.;
}

void EncoderParam::SetQuality(float quality_factor) { sjb::QualityToMatrices(quality_factor, quant_); }

void EncoderParam::SetQuantization(const uint8_t m[2][64], float reduction) {   // api.cc:109-119
  if (reduction <= 1.f) reduction = 1.f;
  if (m == nullptr) return;
  for (int c = 0; c < 2; ++c) {
    for (int i = 0; i < 64; ++i) {
      const int v = static_cast<int>(m[c][i] * 100. / reduction + .5);
      quant_[c][i] = static_cast<uint8_t>(v > 255 ? 255 : v < 1 ? 1 : v);
    }
  }
}

void EncoderParam::SetLimitQuantization(bool limit_quantization, int min_quant_tolerance) {
  use_min_quant_ = limit_quantization;
  if (limit_quantization) SetMinQuantization(quant_, min_quant_tolerance);
}

void EncoderParam::SetMinQuantization(const uint8_t m[2][64], int min_quant_tolerance) {
  use_min_quant_ = true;
  memcpy(min_quant_, m, sizeof(min_quant_));
  min_quant_tolerance_ = min_quant_tolerance < 0 ? 0 : min_quant_tolerance > 100 ? 100 : min_quant_tolerance;
}

void EncoderParam::ResetMetadata() {
  iccp.clear();
  exif.clear();
  app_markers.clear();
  xmp.clear();
  xmp_split_point = 0u;
}

static bool EncodePacked(const uint8_t* pix, int width, int height, int stride, int fmt,
                         const EncoderParam& param, ByteSink* sink) {
  if (pix == nullptr || sink == nullptr) return false;
  if (width <= 0 || height <= 0 || abs(stride) < BytesPerPixel(fmt) * width) return false;
  sjb_params p;
  std::string meta;
  if (!Encoder::Convert(param, fmt, &p) || !BuildMetadata(param, &meta)) {
    sink->Reset();
    return false;
  }
  if (fmt != SJB_PIX_RGB && (p.yuv_mode == SJPEG_YUV_AUTO || p.yuv_mode == SJPEG_YUV_SHARP)) {
    // api.cc:208-224,235-251: these two modes work on a scratch RGB copy of 4-byte pixels
    const size_t rgb_stride = 3 * static_cast<size_t>(width);
    std::unique_ptr<uint8_t[]> rgb(new (std::nothrow) uint8_t[rgb_stride * height]);
    if (rgb == nullptr) return false;
    const int ro = (fmt == SJB_PIX_BGRA) ? 2 : 0, bo = 2 - ro;
    for (int y = 0; y < height; ++y) {
      const uint8_t* s = pix + static_cast<ptrdiff_t>(y) * stride;
      uint8_t* d = rgb.get() + y * rgb_stride;
      for (int x = 0; x < width; ++x, s += 4, d += 3) {
        d[0] = s[ro];
        d[1] = s[1];
        d[2] = s[bo];
      }
    }
    p.pix_fmt = SJB_PIX_RGB;
    return EncodeToSink(rgb.get(), width, height, static_cast<int>(rgb_stride), SJB_PIX_RGB, p, sink, param.memory,
                        meta, &param);
  }
  return EncodeToSink(pix, width, height, stride, fmt, p, sink, param.memory, meta, &param);
}

bool Encode(const uint8_t* rgb, int width, int height, int stride, const EncoderParam& param, ByteSink* sink) {
  return EncodePacked(rgb, width, height, stride, SJB_PIX_RGB, param, sink);
}

size_t Encode(const uint8_t* rgb, int width, int height, int stride, const EncoderParam& param,
              uint8_t** out_data) {
  if (out_data == nullptr) return 0;
  NewArraySink sink;
  if (!Encode(rgb, width, height, stride, param, &sink)) return 0;
  return sink.Release(out_data);
}

bool Encode(const uint8_t* rgb, int width, int height, int stride, const EncoderParam& param,
            std::string* output) {
  if (output == nullptr) return false;
  output->clear();
  ContainerSink<std::string> sink(output);
  return Encode(rgb, width, height, stride, param, &sink);
}

bool EncodeBGRA(const uint8_t* bgra, int width, int height, int stride, const EncoderParam& param, ByteSink* sink) {
  return EncodePacked(bgra, width, height, stride, SJB_PIX_BGRA, param, sink);
}
bool EncodeBGRA(const uint8_t* bgra, int width, int height, int stride, const EncoderParam& param,
                std::string* output) {
  if (output == nullptr) return false;
  output->clear();
  ContainerSink<std::string> sink(output);
  return EncodeBGRA(bgra, width, height, stride, param, &sink);
}
bool EncodeRGBA(const uint8_t* rgba, int width, int height, int stride, const EncoderParam& param, ByteSink* sink) {
  return EncodePacked(rgba, width, height, stride, SJB_PIX_RGBA, param, sink);
}
bool EncodeRGBA(const uint8_t* rgba, int width, int height, int stride, const EncoderParam& param,
                std::string* output) {
  if (output == nullptr) return false;
  output->clear();
  ContainerSink<std::string> sink(output);
  return EncodeRGBA(rgba, width, height, stride, param, &sink);
}

// Planar / semi-planar inputs (encoders.cc:256-507): the colour space is implied by the entry
// point, param.yuv_mode is not consulted.
static bool EncodePlanar(const uint8_t* y, int ys, const uint8_t* u, int us, const uint8_t* v, int vs, int uv_step,
                         int mode, int width, int height, const EncoderParam& param, ByteSink* sink) {
  if (sink == nullptr) return false;
  EncoderParam fixed = param;
  fixed.yuv_mode = static_cast<SjpegYUVMode>(mode);
  sjb_params p;
  std::string meta;
  if (!Encoder::Convert(fixed, SJB_PIX_RGB, &p) || !BuildMetadata(param, &meta)) {
    sink->Reset();
    return false;
  }
  sink->Reset();
  sjb_context* ctx = tls_context.get();
  if (ctx == nullptr) return false;
  sjpeg::MemoryManager* memory = param.memory ? param.memory : &default_memory;
  SearchState st;
  sjb_search search;
  sjpeg::SearchHook default_hook;
  if (!ArmSearch(ctx, param, &default_hook, &st, &search)) return false;
  size_t size = 0;
  const int rc = sjb_encode_planar(ctx, y, ys, u, us, v, vs, uv_step, 0, width, height, &p, nullptr, 0, 0, &size);
  sjb_context_set_search(ctx, nullptr);
  if (rc != SJB_ERR_CAPACITY || size == 0) return false;
  FinishSearch(param, &st, search, true);
  uint8_t* staging = static_cast<uint8_t*>(memory->Alloc(size));
  if (staging == nullptr) return false;
  bool ok = sjb_fetch_output(ctx, staging, 0, size) == SJB_OK;
  ok = ok && CommitWithMetadata(staging, size, meta, sink);
  memory->Free(staging);
  if (!ok) sink->Reset();
  return ok;
}

bool EncodeGray(const uint8_t* gray, int width, int height, int stride, const EncoderParam& param, ByteSink* sink) {
  if (gray == nullptr || sink == nullptr) return false;                    // api.cc:283-292
  if (width <= 0 || height <= 0 || abs(stride) < width) return false;
  return EncodePlanar(gray, stride, nullptr, 0, nullptr, 0, 1, SJPEG_YUV_400, width, height, param, sink);
}
bool EncodeGray(const uint8_t* gray, int width, int height, int stride, const EncoderParam& param,
                std::string* output) {
  if (output == nullptr) return false;
  output->clear();
  ContainerSink<std::string> sink(output);
  return EncodeGray(gray, width, height, stride, param, &sink);
}
bool EncodeNV21(const uint8_t* y, int y_stride, const uint8_t* vu, int vu_stride, int width, int height,
                const EncoderParam& param, ByteSink* output) {                 // encoders.cc:346-380
  if (y == nullptr || vu == nullptr || output == nullptr) return false;
  if (width <= 0 || height <= 0 || abs(y_stride) < width || abs(vu_stride) < 2 * ((width + 1) / 2)) return false;
  return EncodePlanar(y, y_stride, vu + 1, vu_stride, vu, vu_stride, 2, SJPEG_YUV_420, width, height, param, output);
}
bool EncodeNV12(const uint8_t* y, int y_stride, const uint8_t* uv, int uv_stride, int width, int height,
                const EncoderParam& param, ByteSink* output) {
  if (y == nullptr || uv == nullptr || output == nullptr) return false;
  if (width <= 0 || height <= 0 || abs(y_stride) < width || abs(uv_stride) < 2 * ((width + 1) / 2)) return false;
  return EncodePlanar(y, y_stride, uv, uv_stride, uv + 1, uv_stride, 2, SJPEG_YUV_420, width, height, param, output);
}
bool EncodeYUV444(const uint8_t* Y, int Y_stride, const uint8_t* U, int U_stride, const uint8_t* V, int V_stride,
                  int width, int height, const EncoderParam& param, ByteSink* output) {   // encoders.cc:421-440
  if (Y == nullptr || U == nullptr || V == nullptr || output == nullptr) return false;
  if (width <= 0 || height <= 0) return false;
  if (abs(Y_stride) < width || abs(U_stride) < width || abs(V_stride) < width) return false;
  return EncodePlanar(Y, Y_stride, U, U_stride, V, V_stride, 1, SJPEG_YUV_444, width, height, param, output);
}
bool EncodeYUV420(const uint8_t* Y, int Y_stride, const uint8_t* U, int U_stride, const uint8_t* V, int V_stride,
                  int width, int height, const EncoderParam& param, ByteSink* output) {   // encoders.cc:492-507
  if (Y == nullptr || U == nullptr || V == nullptr || output == nullptr) return false;
  if (width <= 0 || height <= 0) return false;
  if (abs(Y_stride) < width || abs(U_stride) < (width + 1) / 2 || abs(V_stride) < (width + 1) / 2) return false;
  return EncodePlanar(Y, Y_stride, U, U_stride, V, V_stride, 1, SJPEG_YUV_420, width, height, param, output);
}

// Search hook (dichotomy.cc:41-75): the default dichotomy the multi-pass search (engine.cu::EncodeSearch) calls back into.
bool SearchHook::Setup(const EncoderParam& param) {
  for_size = (param.target_mode == EncoderParam::TARGET_SIZE);
  target = param.target_value;
  tolerance = static_cast<float>(param.tolerance / 100.);
  qmin = param.qmin < 0 ? 0 : param.qmin;
  qmax = param.qmax > 100 ? 100 : (param.qmax < param.qmin ? param.qmin : param.qmax);
  q = SjpegEstimateQuality(param.GetQuantMatrix(0), false);
  q = q < qmin ? qmin : q > qmax ? qmax : q;
  value = 0;
  pass = 0;
  return true;
}
void SearchHook::NextMatrix(int idx, uint8_t dst[64]) { SjpegQuantMatrix(q, idx != 0, dst); }
bool SearchHook::Update(float result) {
  value = result;
  if (fabsf(value - target) < tolerance * target) return true;
  if (value > target) qmax = q; else qmin = q;
  const float last = q;
  q = static_cast<float>((qmin + qmax) / 2.);
  return fabsf(q - last) < 0.15f;
}

std::shared_ptr<ByteSink> MakeByteSink(std::string* output) {
  return std::shared_ptr<ByteSink>(new (std::nothrow) ContainerSink<std::string>(output));
}
template <>
std::shared_ptr<ByteSink> MakeByteSink(std::vector<uint8_t>* output) {
  return std::shared_ptr<ByteSink>(new (std::nothrow) ContainerSink<std::vector<uint8_t> >(output));
}

}  // namespace sjpeg
