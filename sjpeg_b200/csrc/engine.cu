// engine.cu -- host orchestration behind the C ABI (include/sjpeg_b200.h): contexts, device
// scratch, and the kernel-launch sequence that replaces the reference's per-MCU driver loops
// (Encoder::Encode enc.cc:391-448, SinglePassScan :276-307, SinglePassScanOptimized :323-386,
// CollectHistograms histogram.cc:317-339).  No CPU fallback: every error is reported.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/sjpeg_b200.h"
#include "host_codec.h"
#include "kernels.cuh"

using namespace sjb;

namespace {

enum { kMaxLanes = 4, kHeaderReserve = 2048, kWorstBitsPerBlock = 1696 };

struct DeviceBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  // grow-only; contents are not preserved
  cudaError_t Reserve(size_t need, bool zero = false) {
    if (need <= bytes) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    need = (need + 255) & ~static_cast<size_t>(255);
    cudaError_t e = cudaMalloc(&ptr, need);
    if (e != cudaSuccess) return e;
    bytes = need;
    if (zero) e = cudaMemset(ptr, 0, need);
    return e;
  }
  void Release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <class T> T* as() const { return static_cast<T*>(ptr); }
};

struct HostScratch {     // pinned, for small async up/downloads
  CodeTabs tabs;
  StreamInfo info;
  int32_t hist[2 * 64 * kHistoStride];
  uint32_t freq[2 * 272];
  uint8_t quant[2][64];
  uint8_t header[kHeaderReserve];
};

// one independent pipeline: a stream plus all the scratch one picture needs
struct Lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  DeviceBuffer pix, coef, nzmask, block_bits, tile_sums, tile_offsets, words, ff_sums, ff_offsets, out,
      small;   // small: StreamInfo | CodeTabs | hist | freq | quant
  HostScratch* host = nullptr;
  bool words_dirty = false;
  bool tabs_valid = false;      // d_tabs() mirrors host->tabs
  size_t last_size = 0;         // size of the JPEG left in 'out' by the last sjb_encode
  size_t header_len = 0;        // bytes of host->header mirrored at the start of 'out' (0 = none)
  float ms_f1 = 0, ms_entropy = 0, ms_total = 0;
  unsigned long long launches = 0;

  StreamInfo* d_info() const { return small.as<StreamInfo>(); }
  CodeTabs* d_tabs() const { return reinterpret_cast<CodeTabs*>(small.as<uint8_t>() + 256); }
  int32_t* d_hist() const { return reinterpret_cast<int32_t*>(small.as<uint8_t>() + 256 + 4096); }
  uint32_t* d_freq() const { return reinterpret_cast<uint32_t*>(small.as<uint8_t>() + 256 + 4096 + 68 * 1024); }
  uint8_t* d_quant() const { return small.as<uint8_t>() + 256 + 4096 + 68 * 1024 + 4096; }
};

}  // namespace

struct sjb_context {
  int device = 0;
  int sm_count = 0;
  Lane lanes[kMaxLanes];
  std::string err;
};

namespace {

#define CU(expr)                                                                   \
  do {                                                                             \
    cudaError_t e_ = (expr);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e_);               \
      return (e_ == cudaErrorMemoryAllocation) ? SJB_ERR_NOMEM : SJB_ERR_CUDA;     \
    }                                                                              \
  } while (0)

int InitLane(sjb_context* ctx, Lane* L) {
  if (L->stream) return SJB_OK;
  CU(cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking));
  for (auto& e : L->ev) CU(cudaEventCreate(&e));
  CU(cudaMallocHost(reinterpret_cast<void**>(&L->host), sizeof(HostScratch)));
  CU(L->small.Reserve(256 + 4096 + 68 * 1024 + 4096 + 256, true));
  return SJB_OK;
}

void DestroyLane(Lane* L) {
  if (L->stream) cudaStreamSynchronize(L->stream);
  for (DeviceBuffer* b : {&L->pix, &L->coef, &L->nzmask, &L->block_bits, &L->tile_sums, &L->tile_offsets,
                          &L->words, &L->ff_sums, &L->ff_offsets, &L->out, &L->small})
    b->Release();
  for (auto& e : L->ev) if (e) cudaEventDestroy(e);
  if (L->host) cudaFreeHost(L->host);
  if (L->stream) cudaStreamDestroy(L->stream);
  *L = Lane();
}

struct Plan {
  FrameGeometry g;
  sjb_params p;
  int pstep;
  bool adaptive, optimize, trellis;
  size_t stream_words;      // worst case, multiple of 4
  size_t out_capacity;      // worst case device output
  size_t nb_tiles, ff_tiles;
};

int MakePlan(sjb_context* ctx, int width, int height, long long stride, const sjb_params* params, Plan* plan) {
  if (params == nullptr) return SJB_ERR_ARG;
  plan->p = *params;
  sjb_params& p = plan->p;
  if (p.pix_fmt != SJB_PIX_RGB && p.pix_fmt != SJB_PIX_BGRA && p.pix_fmt != SJB_PIX_RGBA) return SJB_ERR_ARG;
  plan->pstep = (p.pix_fmt == SJB_PIX_RGB) ? 3 : 4;
  if (width <= 0 || height <= 0) return SJB_ERR_ARG;
  const long long astride = stride < 0 ? -stride : stride;
  if (astride < static_cast<long long>(plan->pstep) * width) return SJB_ERR_ARG;   // api.cc:36
  if (!MakeGeometry(p.yuv_mode, width, height, &plan->g)) return SJB_ERR_ARG;     // enc.cc:406
  p.method = std::min(8, std::max(0, p.method));                                  // enc.cc:122
  plan->adaptive = p.method >= 3;
  plan->optimize = p.method != 0 && p.method != 3;
  plan->trellis = p.method >= 7;
  if (p.q_bias < 0 || p.q_bias > 255) return SJB_ERR_ARG;
  const size_t nb = plan->g.nb_blocks();
  plan->stream_words = ((nb * kWorstBitsPerBlock / 32 + 64) + 3) & ~static_cast<size_t>(3);
  plan->out_capacity = kHeaderReserve + 2 * plan->stream_words * 4 + 16;
  plan->nb_tiles = (nb + kTileBlocks - 1) / kTileBlocks;
  plan->ff_tiles = (plan->stream_words * 4 + kStuffTileBytes - 1) / kStuffTileBytes;
  (void)ctx;
  return SJB_OK;
}

int ReserveLane(sjb_context* ctx, Lane* L, const Plan& plan) {
  const size_t nb = plan.g.nb_blocks();
  CU(L->coef.Reserve(nb * 64 * sizeof(int16_t)));
  CU(L->nzmask.Reserve(nb * sizeof(uint32_t)));
  CU(L->block_bits.Reserve(nb * sizeof(uint32_t)));
  CU(L->tile_sums.Reserve(plan.nb_tiles * sizeof(uint32_t)));
  CU(L->tile_offsets.Reserve(plan.nb_tiles * sizeof(unsigned long long)));
  if (plan.stream_words * 4 + 64 > L->words.bytes) {
    CU(L->words.Reserve(plan.stream_words * 4 + 64, true));   // zeroed once, then self-cleaning
    L->words_dirty = false;
  }
  CU(L->ff_sums.Reserve(plan.ff_tiles * sizeof(uint32_t)));
  CU(L->ff_offsets.Reserve(plan.ff_tiles * sizeof(unsigned long long)));
  CU(L->out.Reserve(plan.out_capacity));
  return SJB_OK;
}

// Copies a host picture to the lane's pixel buffer; returns the device address of row 0.
int UploadPicture(sjb_context* ctx, Lane* L, const uint8_t* pix, const Plan& plan, long long stride,
                  const uint8_t** d_row0, long long* d_stride) {
  const size_t row_bytes = static_cast<size_t>(plan.pstep) * plan.g.width;
  const long long astride = stride < 0 ? -stride : stride;
  const int h = plan.g.height;
  if (static_cast<size_t>(astride) <= 2 * row_bytes + 64) {
    // one contiguous span, stride kept (sign included)
    const size_t span = static_cast<size_t>(astride) * (h - 1) + row_bytes;
    const uint8_t* lowest = (stride < 0) ? pix + stride * (h - 1) : pix;
    CU(L->pix.Reserve(span + 64));
    CU(cudaMemcpyAsync(L->pix.ptr, lowest, span, cudaMemcpyHostToDevice, L->stream));
    *d_row0 = L->pix.as<uint8_t>() + ((stride < 0) ? static_cast<size_t>(astride) * (h - 1) : 0);
    *d_stride = stride;
  } else {
    // sparse rows: gather into a tight pitch
    const size_t pitch = (row_bytes + 15) & ~static_cast<size_t>(15);
    CU(L->pix.Reserve(pitch * h + 64));
    if (stride > 0) {
      CU(cudaMemcpy2DAsync(L->pix.ptr, pitch, pix, static_cast<size_t>(stride), row_bytes, h,
                           cudaMemcpyHostToDevice, L->stream));
    } else {
      // bottom-up source: copy from the lowest address, rows come out reversed -> negative pitch
      CU(cudaMemcpy2DAsync(L->pix.ptr, pitch, pix + stride * (h - 1), static_cast<size_t>(astride), row_bytes, h,
                           cudaMemcpyHostToDevice, L->stream));
    }
    *d_row0 = L->pix.as<uint8_t>() + ((stride < 0) ? pitch * (h - 1) : 0);
    *d_stride = (stride < 0) ? -static_cast<long long>(pitch) : static_cast<long long>(pitch);
  }
  return SJB_OK;
}

void LaunchF1(Lane* L, const ImageDesc& img, const FrameGeometry& g, bool raw, const QuantTabs& qt) {
  int16_t* coef = L->coef.as<int16_t>();
  uint32_t* nz = L->nzmask.as<uint32_t>();
  const int mx_full = g.width / g.mcu_size, my_full = g.height / g.mcu_size;
  int mx_fast = 0, my_fast = 0;
  if (F1FastEligible(img)) {
    mx_fast = (g.yuv_mode == kYuv420) ? mx_full : (mx_full & ~1);
    my_fast = my_full;
    if (mx_fast == 0) my_fast = 0;
    if (my_fast == 0) mx_fast = 0;
  }
  if (mx_fast > 0) {
    LaunchF1Fast(img, mx_fast, 0, my_fast, raw, qt, coef, nz, L->stream);
    ++L->launches;
  }
  if (mx_fast < g.mcus_x && my_fast > 0) {        // columns right of the fast region
    LaunchF1Generic(img, mx_fast, 0, g.mcus_x, my_fast, raw, qt, coef, nz, L->stream);
    ++L->launches;
  }
  if (my_fast < g.mcus_y) {                        // rows below it
    LaunchF1Generic(img, 0, my_fast, g.mcus_x, g.mcus_y, raw, qt, coef, nz, L->stream);
    ++L->launches;
  }
}

// Device pipeline for one picture already in device memory.  On return the JPEG is in L->out
// (header included) and L->host->info.out_size holds its size (after the stream is synchronised,
// which this function does only when it has to look at intermediate results).
int EncodeOnLane(sjb_context* ctx, Lane* L, const uint8_t* d_row0, long long d_stride, const Plan& plan,
                 bool timed) {
  const FrameGeometry& g = plan.g;
  HostScratch* H = L->host;
  ImageDesc img;
  img.pix = d_row0;
  img.stride = d_stride;
  img.width = g.width;
  img.height = g.height;
  img.yuv_mode = g.yuv_mode;
  img.pix_fmt = plan.p.pix_fmt;
  img.mcus_x = g.mcus_x;
  img.mcus_y = g.mcus_y;

  uint8_t quant[2][64], min_quant[2][64];
  memcpy(quant, plan.p.quant, sizeof(quant));
  memcpy(min_quant, plan.p.min_quant, sizeof(min_quant));
  QuantTabs qt;
  for (int i = 0; i < 2; ++i) {
    if (!FinalizeQuantizer(quant[i], min_quant[i], plan.p.q_bias, &qt.m[i])) {
      ctx->err = "quantiser entry outside the range of the fused quantise form";
      return SJB_ERR_ARG;
    }
  }
  const size_t nb = g.nb_blocks();
  int16_t* coef = L->coef.as<int16_t>();
  uint32_t* nz = L->nzmask.as<uint32_t>();

  if (L->words_dirty) {
    CU(cudaMemsetAsync(L->words.ptr, 0, L->words.bytes, L->stream));
    L->words_dirty = false;
  }
  if (timed) CU(cudaEventRecord(L->ev[0], L->stream));

  HuffSpec spec[4];   // dc0 dc1 ac0 ac1
  for (int i = 0; i < 4; ++i) DefaultHuffSpec(i >= 2, i & 1, &spec[i]);
  CodeTabs tabs;
  memset(&tabs, 0, sizeof(tabs));
  for (int c = 0; c < 2; ++c) {
    CodesFromSpec(spec[c], tabs.dc[c]);
    CodesFromSpec(spec[2 + c], tabs.ac[c]);
  }
  // Pinned staging (tables, header) is only rewritten when its content changes, and then only
  // after the stream has drained: an earlier async copy may still be reading it.
  auto upload_tabs = [&]() -> int {
    if (L->tabs_valid && memcmp(&H->tabs, &tabs, sizeof(tabs)) == 0) return SJB_OK;
    CU(cudaStreamSynchronize(L->stream));
    H->tabs = tabs;
    CU(cudaMemcpyAsync(L->d_tabs(), &H->tabs, sizeof(CodeTabs), cudaMemcpyHostToDevice, L->stream));
    L->tabs_valid = true;
    return SJB_OK;
  };

  if (plan.adaptive) {
    // enc.cc:425-429 : histogram pass over unquantised coefficients, matrices re-derived on host
    LaunchF1(L, img, g, /*raw=*/true, qt);
    CU(cudaMemsetAsync(L->d_hist(), 0, sizeof(H->hist), L->stream));
    LaunchHistogram(coef, nb, g.mcu_blocks, g.luma_blocks, L->d_hist(), L->stream);
    L->launches += 1;
    CU(cudaMemcpyAsync(H->hist, L->d_hist(), sizeof(H->hist), cudaMemcpyDeviceToHost, L->stream));
    if (timed) CU(cudaEventRecord(L->ev[1], L->stream));
    CU(cudaStreamSynchronize(L->stream));
    AnalyseHistograms(H->hist, g.nb_comps, quant, min_quant, plan.p.qdelta_max_luma, plan.p.qdelta_max_chroma);
    for (int i = (g.nb_comps > 1 ? 1 : 0); i >= 0; --i) {
      if (!FinalizeQuantizer(quant[i], min_quant[i], plan.p.q_bias, &qt.m[i])) return SJB_ERR_ARG;
    }
    if (plan.trellis) {
      // rate model = default AC tables (enc.cc:334)
      memcpy(H->quant, quant, sizeof(quant));
      CU(cudaMemcpyAsync(L->d_quant(), H->quant, sizeof(quant), cudaMemcpyHostToDevice, L->stream));
      { const int rc = upload_tabs(); if (rc != SJB_OK) return rc; }
      LaunchTrellis(coef, nz, nb, g.mcu_blocks, g.luma_blocks, qt, L->d_quant(), L->d_tabs(), L->stream);
    } else {
      LaunchRequantize(coef, nz, nb, g.mcu_blocks, g.luma_blocks, qt, L->stream);
    }
    L->launches += 1;
  } else {
    LaunchF1(L, img, g, /*raw=*/false, qt);
    if (timed) CU(cudaEventRecord(L->ev[1], L->stream));
  }

  if (plan.optimize) {
    // enc.cc:344-374 : symbol statistics -> optimal tables
    CU(cudaMemsetAsync(L->d_freq(), 0, sizeof(H->freq), L->stream));
    LaunchSymbolStats(coef, nz, nb, g.mcu_blocks, g.luma_blocks, L->d_freq(), L->stream);
    L->launches += 1;
    CU(cudaMemcpyAsync(H->freq, L->d_freq(), sizeof(H->freq), cudaMemcpyDeviceToHost, L->stream));
    CU(cudaStreamSynchronize(L->stream));
    const int nb_tables = (g.nb_comps == 1) ? 1 : 2;
    for (int c = 0; c < nb_tables; ++c) {
      OptimalHuffSpec(H->freq + 272 * c + 256, 12, &spec[c]);
      OptimalHuffSpec(H->freq + 272 * c, 256, &spec[2 + c]);
      CodesFromSpec(spec[c], tabs.dc[c]);
      CodesFromSpec(spec[2 + c], tabs.ac[c]);
    }
  }
  { const int rc = upload_tabs(); if (rc != SJB_OK) return rc; }

  std::vector<uint8_t> header;
  header.reserve(1024);
  AppendHeaders(g, quant, spec, &header);
  if (header.size() > kHeaderReserve) return SJB_ERR_ARG;
  if (L->header_len != header.size() || memcmp(H->header, header.data(), header.size()) != 0) {
    CU(cudaStreamSynchronize(L->stream));
    memcpy(H->header, header.data(), header.size());
    L->header_len = header.size();
  }
  // the header is re-sent every time: 'out' may have been reallocated or overwritten
  CU(cudaMemcpyAsync(L->out.ptr, H->header, header.size(), cudaMemcpyHostToDevice, L->stream));

  L->words_dirty = true;
  LaunchBlockBits(coef, nz, nb, g.mcu_blocks, g.luma_blocks, L->d_tabs(), L->block_bits.as<uint32_t>(),
                  L->tile_sums.as<uint32_t>(), L->stream);
  LaunchScanTiles(L->tile_sums.as<uint32_t>(), plan.nb_tiles, L->tile_offsets.as<unsigned long long>(),
                  L->d_info(), L->stream);
  LaunchPack(coef, nz, nb, g.mcu_blocks, g.luma_blocks, L->d_tabs(), L->block_bits.as<uint32_t>(),
             L->tile_offsets.as<unsigned long long>(), L->words.as<uint32_t>(), L->stream);
  LaunchStuff(L->words.as<uint32_t>(), plan.stream_words, L->ff_sums.as<uint32_t>(),
              L->ff_offsets.as<unsigned long long>(), L->d_info(), L->out.as<uint8_t>(), header.size(),
              L->stream);
  L->launches += 6;
  CU(cudaGetLastError());
  L->words_dirty = false;   // the stuffing kernel zeroes every word it consumed
  if (timed) CU(cudaEventRecord(L->ev[2], L->stream));
  CU(cudaMemcpyAsync(&H->info, L->d_info(), sizeof(StreamInfo), cudaMemcpyDeviceToHost, L->stream));
  return SJB_OK;
}

int FinishTimings(sjb_context* ctx, Lane* L, bool adaptive) {
  CU(cudaEventSynchronize(L->ev[2]));
  CU(cudaEventElapsedTime(&L->ms_total, L->ev[0], L->ev[2]));
  if (!adaptive) {
    CU(cudaEventElapsedTime(&L->ms_f1, L->ev[0], L->ev[1]));
    CU(cudaEventElapsedTime(&L->ms_entropy, L->ev[1], L->ev[2]));
  } else {
    CU(cudaEventElapsedTime(&L->ms_f1, L->ev[0], L->ev[1]));
    L->ms_entropy = L->ms_total - L->ms_f1;
  }
  return SJB_OK;
}

int IsDevicePointer(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

uint32_t sjb_version(void) { return 0x000101; }   // tracks SJPEG_VERSION (sjpeg.h:28)

int sjb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int sjb_context_create(int device, sjb_context** out) {
  if (out == nullptr) return SJB_ERR_ARG;
  *out = nullptr;
  if (device < 0 || device >= sjb_device_count()) return SJB_ERR_CUDA;
  sjb_context* ctx = new (std::nothrow) sjb_context();
  if (ctx == nullptr) return SJB_ERR_NOMEM;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    delete ctx;
    return SJB_ERR_CUDA;
  }
  const int rc = InitLane(ctx, &ctx->lanes[0]);
  if (rc != SJB_OK) {
    DestroyLane(&ctx->lanes[0]);
    delete ctx;
    return rc;
  }
  *out = ctx;
  return SJB_OK;
}

void sjb_context_destroy(sjb_context* ctx) {
  if (ctx == nullptr) return;
  cudaSetDevice(ctx->device);
  for (auto& L : ctx->lanes) DestroyLane(&L);
  delete ctx;
}

const char* sjb_last_error(const sjb_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void sjb_quality_to_matrices(float quality, uint8_t out[2][64]) { QualityToMatrices(quality, out); }

void sjb_params_default(sjb_params* p, float quality, int method, int yuv_mode) {
  if (p == nullptr) return;
  memset(p, 0, sizeof(*p));
  p->yuv_mode = yuv_mode;
  p->method = method;
  p->pix_fmt = SJB_PIX_RGB;
  QualityToMatrices(quality, p->quant);
  memset(p->min_quant, 1, sizeof(p->min_quant));
  p->q_bias = 0x78;
  p->qdelta_max_luma = 12;
  p->qdelta_max_chroma = 1;
}

size_t sjb_max_output_size(int width, int height, int yuv_mode) {
  FrameGeometry g;
  if (!MakeGeometry(yuv_mode, width, height, &g)) return 0;
  return kHeaderReserve + 2 * (g.nb_blocks() * kWorstBitsPerBlock / 8 + 256) + 16;
}

void* sjb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void sjb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int sjb_encode(sjb_context* ctx, const uint8_t* pix, int pix_on_device, int width, int height,
               long long stride, const sjb_params* params, uint8_t* out, int out_on_device,
               size_t out_capacity, size_t* out_size) {
  if (ctx == nullptr || pix == nullptr || out_size == nullptr) return SJB_ERR_ARG;
  *out_size = 0;
  ctx->lanes[0].last_size = 0;
  ctx->err.clear();
  Plan plan;
  int rc = MakePlan(ctx, width, height, stride, params, &plan);
  if (rc != SJB_OK) return rc;
  CU(cudaSetDevice(ctx->device));
  Lane* L = &ctx->lanes[0];
  rc = ReserveLane(ctx, L, plan);
  if (rc != SJB_OK) return rc;
  const uint8_t* d_row0 = pix;
  long long d_stride = stride;
  if (!pix_on_device) {
    rc = UploadPicture(ctx, L, pix, plan, stride, &d_row0, &d_stride);
    if (rc != SJB_OK) return rc;
  }
  L->launches = 0;
  rc = EncodeOnLane(ctx, L, d_row0, d_stride, plan, /*timed=*/true);
  if (rc != SJB_OK) {
    L->words_dirty = true;
    return rc;
  }
  CU(cudaStreamSynchronize(L->stream));
  FinishTimings(ctx, L, plan.adaptive);
  const size_t size = static_cast<size_t>(L->host->info.out_size);
  *out_size = size;
  L->last_size = size;
  if (out == nullptr || size > out_capacity) return SJB_ERR_CAPACITY;
  CU(cudaMemcpyAsync(out, L->out.ptr, size, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                     L->stream));
  CU(cudaStreamSynchronize(L->stream));
  return SJB_OK;
}

int sjb_fetch_output(sjb_context* ctx, uint8_t* out, int out_on_device, size_t out_capacity) {
  if (ctx == nullptr || out == nullptr) return SJB_ERR_ARG;
  Lane* L = &ctx->lanes[0];
  const size_t size = L->last_size;
  if (size == 0) return SJB_ERR_ARG;
  if (size > out_capacity) return SJB_ERR_CAPACITY;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(out, L->out.ptr, size, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                     L->stream));
  CU(cudaStreamSynchronize(L->stream));
  return SJB_OK;
}

int sjb_encode_batch(sjb_context* ctx, int n, const uint8_t* const* pix, int pix_on_device, int width,
                     int height, long long stride, const sjb_params* params, uint8_t* const* out,
                     int out_on_device, size_t out_capacity, size_t* sizes) {
  if (ctx == nullptr || pix == nullptr || out == nullptr || sizes == nullptr || n < 0) return SJB_ERR_ARG;
  ctx->err.clear();
  Plan plan;
  int rc = MakePlan(ctx, width, height, stride, params, &plan);
  if (rc != SJB_OK) return rc;
  CU(cudaSetDevice(ctx->device));
  const int nl = std::min<int>(kMaxLanes, std::max(1, n));
  for (int l = 0; l < nl; ++l) {
    if ((rc = InitLane(ctx, &ctx->lanes[l])) != SJB_OK) return rc;
    if ((rc = ReserveLane(ctx, &ctx->lanes[l], plan)) != SJB_OK) return rc;
  }
  // software pipeline over the lanes: frame i runs on lane i % nl; a lane is drained (size read,
  // bytes copied out) right before it is reused.
  auto drain = [&](int i) -> int {
    Lane* L = &ctx->lanes[i % nl];
    CU(cudaStreamSynchronize(L->stream));
    const size_t size = static_cast<size_t>(L->host->info.out_size);
    sizes[i] = size;
    if (out[i] == nullptr || size > out_capacity) return SJB_ERR_CAPACITY;
    CU(cudaMemcpyAsync(out[i], L->out.ptr, size, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                       L->stream));
    return SJB_OK;
  };
  int first_err = SJB_OK;
  for (int i = 0; i < n; ++i) {
    Lane* L = &ctx->lanes[i % nl];
    if (i >= nl) {
      rc = drain(i - nl);
      if (rc != SJB_OK && first_err == SJB_OK) first_err = rc;
    }
    if (pix[i] == nullptr) return SJB_ERR_ARG;
    const uint8_t* d_row0 = pix[i];
    long long d_stride = stride;
    if (!pix_on_device) {
      if ((rc = UploadPicture(ctx, L, pix[i], plan, stride, &d_row0, &d_stride)) != SJB_OK) return rc;
    }
    if ((rc = EncodeOnLane(ctx, L, d_row0, d_stride, plan, false)) != SJB_OK) {
      L->words_dirty = true;
      return rc;
    }
  }
  for (int i = std::max(0, n - nl); i < n; ++i) {
    rc = drain(i);
    if (rc != SJB_OK && first_err == SJB_OK) first_err = rc;
  }
  for (int l = 0; l < nl; ++l) CU(cudaStreamSynchronize(ctx->lanes[l].stream));
  return first_err;
}

int sjb_stage_coefficients(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                           const sjb_params* params, int quantise, int16_t* coef, uint32_t* nzmask) {
  if (ctx == nullptr || pix == nullptr || coef == nullptr) return SJB_ERR_ARG;
  ctx->err.clear();
  Plan plan;
  int rc = MakePlan(ctx, width, height, stride, params, &plan);
  if (rc != SJB_OK) return rc;
  CU(cudaSetDevice(ctx->device));
  Lane* L = &ctx->lanes[0];
  if ((rc = ReserveLane(ctx, L, plan)) != SJB_OK) return rc;
  const uint8_t* d_row0;
  long long d_stride;
  if ((rc = UploadPicture(ctx, L, pix, plan, stride, &d_row0, &d_stride)) != SJB_OK) return rc;
  ImageDesc img = {d_row0, d_stride, width, height, plan.g.yuv_mode, plan.p.pix_fmt, plan.g.mcus_x, plan.g.mcus_y};
  uint8_t quant[2][64], min_quant[2][64];
  memcpy(quant, plan.p.quant, sizeof(quant));
  memcpy(min_quant, plan.p.min_quant, sizeof(min_quant));
  QuantTabs qt;
  for (int i = 0; i < 2; ++i) {
    if (!FinalizeQuantizer(quant[i], min_quant[i], plan.p.q_bias, &qt.m[i])) return SJB_ERR_ARG;
  }
  LaunchF1(L, img, plan.g, quantise == 0, qt);
  CU(cudaGetLastError());
  const size_t nb = plan.g.nb_blocks();
  CU(cudaMemcpyAsync(coef, L->coef.ptr, nb * 64 * sizeof(int16_t), cudaMemcpyDeviceToHost, L->stream));
  if (quantise && nzmask) {
    CU(cudaMemcpyAsync(nzmask, L->nzmask.ptr, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, L->stream));
  }
  CU(cudaStreamSynchronize(L->stream));
  return SJB_OK;
}

int sjb_stage_histogram(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                        const sjb_params* params, int32_t* counts) {
  if (ctx == nullptr || pix == nullptr || counts == nullptr) return SJB_ERR_ARG;
  ctx->err.clear();
  Plan plan;
  int rc = MakePlan(ctx, width, height, stride, params, &plan);
  if (rc != SJB_OK) return rc;
  CU(cudaSetDevice(ctx->device));
  Lane* L = &ctx->lanes[0];
  if ((rc = ReserveLane(ctx, L, plan)) != SJB_OK) return rc;
  const uint8_t* d_row0;
  long long d_stride;
  if ((rc = UploadPicture(ctx, L, pix, plan, stride, &d_row0, &d_stride)) != SJB_OK) return rc;
  ImageDesc img = {d_row0, d_stride, width, height, plan.g.yuv_mode, plan.p.pix_fmt, plan.g.mcus_x, plan.g.mcus_y};
  QuantTabs qt;
  memset(&qt, 0, sizeof(qt));
  LaunchF1(L, img, plan.g, true, qt);
  CU(cudaMemsetAsync(L->d_hist(), 0, sizeof(L->host->hist), L->stream));
  LaunchHistogram(L->coef.as<int16_t>(), plan.g.nb_blocks(), plan.g.mcu_blocks, plan.g.luma_blocks, L->d_hist(),
                  L->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(counts, L->d_hist(), sizeof(L->host->hist), cudaMemcpyDeviceToHost, L->stream));
  CU(cudaStreamSynchronize(L->stream));
  return SJB_OK;
}

int sjb_stage_symbol_stats(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                           const sjb_params* params, uint32_t* freq_ac, uint32_t* freq_dc) {
  if (ctx == nullptr || pix == nullptr || freq_ac == nullptr || freq_dc == nullptr) return SJB_ERR_ARG;
  ctx->err.clear();
  Plan plan;
  int rc = MakePlan(ctx, width, height, stride, params, &plan);
  if (rc != SJB_OK) return rc;
  CU(cudaSetDevice(ctx->device));
  Lane* L = &ctx->lanes[0];
  if ((rc = ReserveLane(ctx, L, plan)) != SJB_OK) return rc;
  const uint8_t* d_row0;
  long long d_stride;
  if ((rc = UploadPicture(ctx, L, pix, plan, stride, &d_row0, &d_stride)) != SJB_OK) return rc;
  ImageDesc img = {d_row0, d_stride, width, height, plan.g.yuv_mode, plan.p.pix_fmt, plan.g.mcus_x, plan.g.mcus_y};
  uint8_t quant[2][64], min_quant[2][64];
  memcpy(quant, plan.p.quant, sizeof(quant));
  memcpy(min_quant, plan.p.min_quant, sizeof(min_quant));
  QuantTabs qt;
  for (int i = 0; i < 2; ++i) {
    if (!FinalizeQuantizer(quant[i], min_quant[i], plan.p.q_bias, &qt.m[i])) return SJB_ERR_ARG;
  }
  LaunchF1(L, img, plan.g, false, qt);
  CU(cudaMemsetAsync(L->d_freq(), 0, sizeof(L->host->freq), L->stream));
  LaunchSymbolStats(L->coef.as<int16_t>(), L->nzmask.as<uint32_t>(), plan.g.nb_blocks(), plan.g.mcu_blocks,
                    plan.g.luma_blocks, L->d_freq(), L->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(L->host->freq, L->d_freq(), sizeof(L->host->freq), cudaMemcpyDeviceToHost, L->stream));
  CU(cudaStreamSynchronize(L->stream));
  for (int c = 0; c < 2; ++c) {
    memcpy(freq_ac + 256 * c, L->host->freq + 272 * c, 256 * sizeof(uint32_t));
    memcpy(freq_dc + 12 * c, L->host->freq + 272 * c + 256, 12 * sizeof(uint32_t));
  }
  return SJB_OK;
}

int sjb_last_timings(const sjb_context* ctx, float ms[3]) {
  if (ctx == nullptr || ms == nullptr) return SJB_ERR_ARG;
  ms[0] = ctx->lanes[0].ms_f1;
  ms[1] = ctx->lanes[0].ms_entropy;
  ms[2] = ctx->lanes[0].ms_total;
  return SJB_OK;
}

int sjb_bench_device(sjb_context* ctx, int n, const uint8_t* const* dev_pix, int width, int height,
                     long long stride, const sjb_params* params, int iters, float* total_ms,
                     float* f1_ms, size_t* jpeg_bytes, unsigned long long* launches) {
  if (ctx == nullptr || dev_pix == nullptr || n <= 0 || iters <= 0 || total_ms == nullptr) return SJB_ERR_ARG;
  ctx->err.clear();
  Plan plan;
  int rc = MakePlan(ctx, width, height, stride, params, &plan);
  if (rc != SJB_OK) return rc;
  CU(cudaSetDevice(ctx->device));
  const int nl = std::min<int>(kMaxLanes, n);
  for (int l = 0; l < nl; ++l) {
    if ((rc = InitLane(ctx, &ctx->lanes[l])) != SJB_OK) return rc;
    if ((rc = ReserveLane(ctx, &ctx->lanes[l], plan)) != SJB_OK) return rc;
    ctx->lanes[l].launches = 0;
  }
  CU(cudaDeviceSynchronize());
  cudaEvent_t t0, t1;
  CU(cudaEventCreate(&t0));
  CU(cudaEventCreate(&t1));
  Lane* L0 = &ctx->lanes[0];
  CU(cudaEventRecord(t0, L0->stream));
  for (int l = 1; l < nl; ++l) CU(cudaStreamWaitEvent(ctx->lanes[l].stream, t0, 0));
  double f1_sum = 0;
  int f1_n = 0;
  for (int it = 0; it < iters; ++it) {
    for (int i = 0; i < n; ++i) {
      Lane* L = &ctx->lanes[i % nl];
      const bool timed = (i % nl == 0) && !plan.adaptive;
      if ((rc = EncodeOnLane(ctx, L, dev_pix[i], stride, plan, timed)) != SJB_OK) return rc;
      if (timed && it == iters - 1 && i + nl >= n) {
        // sample the fused kernel's time on lane 0 once per run (events add no sync)
      }
    }
  }
  for (int l = 1; l < nl; ++l) {
    CU(cudaEventRecord(ctx->lanes[l].ev[3], ctx->lanes[l].stream));
    CU(cudaStreamWaitEvent(L0->stream, ctx->lanes[l].ev[3], 0));
  }
  CU(cudaEventRecord(t1, L0->stream));
  CU(cudaEventSynchronize(t1));
  CU(cudaEventElapsedTime(total_ms, t0, t1));
  if (!plan.adaptive) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, L0->ev[0], L0->ev[1]) == cudaSuccess) {
      f1_sum += ms;
      ++f1_n;
    }
  }
  if (f1_ms) *f1_ms = f1_n ? static_cast<float>(f1_sum / f1_n) : 0.f;
  if (jpeg_bytes) *jpeg_bytes = static_cast<size_t>(L0->host->info.out_size);
  if (launches) {
    *launches = 0;
    for (int l = 0; l < nl; ++l) *launches += ctx->lanes[l].launches;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  return SJB_OK;
}

int sjb_bench_f1(sjb_context* ctx, int n, const uint8_t* const* dev_pix, int width, int height,
                 long long stride, const sjb_params* params, int iters, float* ms_per_launch) {
  if (ctx == nullptr || dev_pix == nullptr || n <= 0 || iters <= 0 || ms_per_launch == nullptr) return SJB_ERR_ARG;
  ctx->err.clear();
  Plan plan;
  int rc = MakePlan(ctx, width, height, stride, params, &plan);
  if (rc != SJB_OK) return rc;
  CU(cudaSetDevice(ctx->device));
  Lane* L = &ctx->lanes[0];
  if ((rc = ReserveLane(ctx, L, plan)) != SJB_OK) return rc;
  uint8_t quant[2][64], min_quant[2][64];
  memcpy(quant, plan.p.quant, sizeof(quant));
  memcpy(min_quant, plan.p.min_quant, sizeof(min_quant));
  QuantTabs qt;
  for (int i = 0; i < 2; ++i) {
    if (!FinalizeQuantizer(quant[i], min_quant[i], plan.p.q_bias, &qt.m[i])) return SJB_ERR_ARG;
  }
  CU(cudaStreamSynchronize(L->stream));
  CU(cudaEventRecord(L->ev[0], L->stream));
  for (int it = 0; it < iters; ++it) {
    for (int i = 0; i < n; ++i) {
      ImageDesc img = {dev_pix[i], stride, width, height, plan.g.yuv_mode, plan.p.pix_fmt, plan.g.mcus_x, plan.g.mcus_y};
      LaunchF1(L, img, plan.g, false, qt);
    }
  }
  CU(cudaEventRecord(L->ev[1], L->stream));
  CU(cudaEventSynchronize(L->ev[1]));
  CU(cudaGetLastError());
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, L->ev[0], L->ev[1]));
  *ms_per_launch = ms / (static_cast<float>(n) * iters);
  return SJB_OK;
}

}  // extern "C"
