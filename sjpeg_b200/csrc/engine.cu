// engine.cu -- host orchestration behind the C ABI (include/sjpeg_b200.h): contexts, device
// scratch, and the kernel-launch sequence that replaces the reference's per-MCU driver loops
// (Encoder::Encode enc.cc:391-448, SinglePassScan :276-307, SinglePassScanOptimized :323-386,
// CollectHistograms histogram.cc:317-339).  No CPU fallback: every error is reported.
//
// Unit of work = a GROUP of up to kMaxGroup pictures with identical geometry and settings,
// processed by one launch of each kernel (kernels.cuh).  A context owns up to kMaxLanes lanes
// (stream + scratch for one group); groups of a batch rotate over the lanes so that the copies
// and kernels of different groups overlap.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/sjpeg_b200.h"
#include "host_codec.h"
#include "host_pool.h"
#include "host_stager.h"
#include "kernels.cuh"
#include "nccl_dyn.h"
#include "sharp.cuh"

using namespace sjb;

namespace {

// Concurrent sharp conversions of a batch.  A 4K conversion is 4 iterations x a cluster of 4 CTAs = 16 SMs
// of 148 and is bound by its row latency, so more of them in flight is free: 16 x 4K, device-resident,
// 10.8 ms with 4 at a time, 6.2 ms with 8 (21.6 Gpix/s, above what PCIe delivers).
#ifndef SJB_SHARP_STREAMS
#define SJB_SHARP_STREAMS 8
#endif
enum { kSharpStreams = SJB_SHARP_STREAMS };
// Lanes (stream + scratch of one group) per context.  16-frame 4K device pipeline, gen B: 2 lanes 464, 4 lanes
// 477, 6 lanes 487, 8 lanes 490 Gpix/s (gen A 223 / 222 / 228 / 227): the tail of one group's kernels overlaps
// the head of the next ones'.  Lanes are only allocated when a batch has that many groups.
#ifndef SJB_MAX_LANES
#define SJB_MAX_LANES 6
#endif
// Batches that arrive from host memory keep to four: their uploads share the link, and with six streams
// copying at once every group's pixels land later (e2e 16 x 4K 18.0 -> 16.8 Gpix/s with six).
enum { kHostLanes = 4 };
enum { kMaxLanes = SJB_MAX_LANES, kHeaderReserve = 2048, kWorstBitsPerBlock = 1696, kHeadCopyBytes = 1 << 20 };
// Pictures per launch = this budget / coefficient bytes per picture (at most kMaxGroup = 16): 16
// pictures at 4K and 1080p, 4 at 8K.  Measured at 4K: 8 pictures per launch against 4 shortened the
// tail of the F1 grid (6.8 instead of 3.4 waves of CTAs: 13.8 -> 12.6 us per picture) and amortised
// the entropy stage's launches (13.6 -> 10.4 us per picture); 16 against 8 gave another 3.5 % on the
// whole pipeline.  What the entropy kernel really reads back -- bitmaps and the non-zero sectors --
// still fits the 126 MB L2 for photographic content.  SJB_GROUP_BUDGET_MB overrides the default.
size_t GroupCoefBudget() {
  static const size_t v = [] {
    const char* e = getenv("SJB_GROUP_BUDGET_MB");
    const long mb = e ? atol(e) : 400;
    return static_cast<size_t>(mb > 0 ? mb : 400) << 20;
  }();
  return v;
}

struct DeviceBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  // grow-only; contents are not preserved.  *grew reports a reallocation.
  cudaError_t Reserve(size_t need, bool zero = false, bool* grew = nullptr) {
    if (need <= bytes) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    need = (need + 255) & ~static_cast<size_t>(255);
    cudaError_t e = cudaMalloc(&ptr, need);
    if (e != cudaSuccess) return e;
    bytes = need;
    if (grew) *grew = true;
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

struct HostScratch {     // pinned, for small async up/downloads; one slot per picture of a group
  CodeTabs tabs[kMaxGroup];
  QuantTabs qtabs[kMaxGroup];
  StreamInfo info[kMaxGroup];
  int32_t hist[kMaxGroup][2 * 64 * kHistoStride];
  uint32_t freq[kMaxGroup][2 * 272];
  uint8_t quant[kMaxGroup][2][64];
  int aq_fail[kMaxGroup];
  CodeTabs def_tabs[kMaxGroup];              // written once, before their only upload
  uint8_t header[kMaxGroup][kHeaderReserve];
};

// device mirror of the small per-picture records
struct SmallLayout {
  StreamInfo info[kMaxGroup];
  CodeTabs tabs[kMaxGroup];
  QuantTabs qtabs[kMaxGroup];
  int32_t hist[kMaxGroup][2 * 64 * kHistoStride];
  uint32_t freq[kMaxGroup][2 * 272];
  uint8_t quant[kMaxGroup][2][64];
  int aq_fail[kMaxGroup];                    // A1: a derived matrix entry the fused quantiser cannot express (right behind
                                             // quant, as in HostScratch: the two travel to the host in one copy)
  uint32_t trellis_sort[kMaxGroup][128];
  CodeTabs def_tabs[kMaxGroup];              // the default tables, for the trellis' rate model (enc.cc:334); written once
  AqFit aq_fit[kMaxGroup][2][64];            // A1 scratch
};

// one independent pipeline: a stream plus all the scratch one group needs
struct Lane {
  cudaStream_t stream = nullptr;
  cudaStream_t hi = nullptr;        // SJB_ES_PRIORITY: highest-priority stream for E + S
  cudaEvent_t hi_ev = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t phase_ev = nullptr;   // end of the current phase of the group on this lane (GroupJob)
  // timed groups: begin/end of each kernel stage -- 0 F1, 1 H1, 2 Q1 or T1, 3 S1, 4 E, 5 S
  cudaEvent_t kev[6][2] = {};
  bool kev_set[6] = {false, false, false, false, false, false};
  DeviceBuffer pix, coef, nzmask, words, out, state, small, raw, perm;   // perm: block order of the trellis
  HostScratch* host = nullptr;
  GroupBuffers gb = {};
  int group_capacity = 0;        // pictures the buffers are laid out for
  size_t pix_pitch = 0;
  bool words_dirty = false;
  bool def_tabs_valid = false;   // d_small()->def_tabs hold the default code tables
  int tabs_valid = 0;            // number of leading device tabs[] slots that mirror host->tabs
  int header_valid = 0;          // same for the header bytes at the start of each out slot
  unsigned header_len[kMaxGroup] = {0};
  size_t last_size = 0;          // size of the JPEG left in out slot 0 by the last sjb_encode
  float ms_f1 = 0, ms_entropy = 0, ms_total = 0;
  unsigned long long launches = 0;
  int timed_frames = 0;          // pictures in the last timed group
  SmallLayout* d_small() const { return small.as<SmallLayout>(); }
};

}  // namespace

// More hardware work queues than the default 8, so that the streams of a context (lanes, uploads,
// collectives) do not share queues with each other or with the host program's own streams.  Only
// effective if the library is loaded before the CUDA context exists; never overrides the user.
namespace {
struct MoreConnections {
  MoreConnections() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_more_connections;
}  // namespace

struct sjb_context {
  int device = 0;
  int sm_count = 0;
  Lane lanes[kMaxLanes];
  std::string err;
  sjb_search* search = nullptr;   // armed by sjb_context_set_search for the next single encode
  // whole-picture passes in front of the block pipeline (sharp.cu)
  DeviceBuffer sharp_scratch, sharp_planes, sharp_tabs, risk_table, risk_sums;
  // batches in AUTO / SHARP mode: every picture resident, planes of the SHARP subset, one scratch + stream per
  // concurrent conversion
  DeviceBuffer batch_pix, batch_planes, sharp_scratch_n[kSharpStreams];
  cudaStream_t sharp_stream[kSharpStreams] = {};
  size_t risk_host_cap = 0;                  // unsigned long longs in risk_host
  unsigned long long* risk_host = nullptr;   // pinned, 3 sums
  int risk_table_version = 0;                // version of the process-wide table held in risk_table
  // where sjb_bench_device left each picture of its last round: {lane, slot, turn}; a picture can be
  // fetched (sjb_bench_output) while no later group has reused its lane
  struct BenchSlot { int lane, slot, turn; };
  std::vector<BenchSlot> bench_slots;
  int bench_last_turn[kMaxLanes];
  // Size queries (sjb_encode with out == NULL, what the drop-in facade does) copy the head of the
  // JPEG into this pinned buffer in the same stream as the sizes, so that a typical file needs one
  // synchronisation and no second device-to-host copy (sjb_fetch_output reads it from here).
  uint8_t* head_copy = nullptr;              // pinned, kHeadCopyBytes
  size_t head_valid = 0;                     // bytes of lane 0's out slot 0 mirrored in head_copy
  HostStager stager;                         // threaded upload of pageable pictures (host_stager.h)
  HostPool pool;                             // per-picture host analysis of a group, in parallel (host_pool.h)
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
// The C ABI promises that no exception leaves the library (include/sjpeg_b200.h): the entry points
// are function-try-blocks; the only exceptions the host code can meet are failed allocations of
// its small std::vector / std::string temporaries.
#define SJB_NOTHROW_END                                   \
  catch (const std::bad_alloc&) { return SJB_ERR_NOMEM; } \
  catch (...) { return SJB_ERR_CUDA; }
#define RC(expr)                   \
  do {                             \
    const int rc_ = (expr);        \
    if (rc_ != SJB_OK) return rc_; \
  } while (0)

int InitLane(sjb_context* ctx, Lane* L) {
  if (L->stream) return SJB_OK;
  CU(cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking));
  for (auto& e : L->ev) CU(cudaEventCreate(&e));
  CU(cudaEventCreateWithFlags(&L->phase_ev, cudaEventDisableTiming));
  for (auto& pair : L->kev) for (auto& e : pair) CU(cudaEventCreate(&e));
  CU(cudaMallocHost(reinterpret_cast<void**>(&L->host), sizeof(HostScratch)));
  memset(L->host, 0, sizeof(HostScratch));
  CU(L->small.Reserve(sizeof(SmallLayout), true));
  return SJB_OK;
}

void DestroyLane(Lane* L) {
  if (L->stream) cudaStreamSynchronize(L->stream);
  if (L->hi) {
    cudaStreamSynchronize(L->hi);
    cudaStreamDestroy(L->hi);
    L->hi = nullptr;
  }
  if (L->hi_ev) {
    cudaEventDestroy(L->hi_ev);
    L->hi_ev = nullptr;
  }
  for (DeviceBuffer* b : {&L->pix, &L->coef, &L->nzmask, &L->words, &L->out, &L->state, &L->small, &L->raw, &L->perm}) b->Release();
  for (auto& e : L->ev) if (e) cudaEventDestroy(e);
  if (L->phase_ev) cudaEventDestroy(L->phase_ev);
  for (auto& pair : L->kev) for (auto& e : pair) if (e) cudaEventDestroy(e);
  if (L->host) cudaFreeHost(L->host);
  if (L->stream) cudaStreamDestroy(L->stream);
  *L = Lane();
}

struct Plan {
  FrameGeometry g;
  sjb_params p;
  int pstep;
  bool adaptive, optimize, trellis;
  size_t stream_words;      // worst case per picture, multiple of 4
  size_t out_capacity;      // worst case per picture
  size_t nb_tiles, ff_tiles;
  int group;                // pictures per launch
};

int MakePlan(int width, int height, long long stride, const sjb_params* params, Plan* plan) {
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
  // the candidate steps are q0 - 12 .. q0 + 12: the reference asserts the limits stay inside (histogram.cc:179-183)
  // and indexes past its tables otherwise; here such settings are refused
  if (p.qdelta_max_luma > 12 || p.qdelta_max_chroma > 12) return SJB_ERR_ARG;
  // a quantiser step of 0 does not exist (the reference's matrices and minima start at 1): a lower bound of 0
  // means "none", and entries below their bound are raised to it by the quantiser set-up as in the reference
  for (int m = 0; m < 2; ++m) {
    for (int i = 0; i < 64; ++i) {
      if (p.min_quant[m][i] == 0) p.min_quant[m][i] = 1;
    }
  }
  const size_t nb = plan->g.nb_blocks();
  plan->stream_words = ((nb * kWorstBitsPerBlock / 32 + 64) + 3) & ~static_cast<size_t>(3);
  plan->out_capacity = (kHeaderReserve + 2 * plan->stream_words * 4 + 16 + 255) & ~static_cast<size_t>(255);
  plan->nb_tiles = (nb + kTileBlocks - 1) / kTileBlocks + 1;   // + the tile counter of the entropy kernel
  plan->ff_tiles = (plan->stream_words * 4 + kStuffTileBytes - 1) / kStuffTileBytes + 1;   // + the tile counter
  const size_t coef_bytes = nb * 128;
  plan->group = static_cast<int>(std::min<size_t>(kMaxGroup, std::max<size_t>(1, GroupCoefBudget() / coef_bytes)));
  return SJB_OK;
}

// (Re)lays out the lane's buffers for `frames` pictures of this plan.
int ReserveLane(sjb_context* ctx, Lane* L, const Plan& plan, int frames) {
  const size_t nb = plan.g.nb_blocks();
  frames = std::max(frames, 1);
  const size_t f = static_cast<size_t>(frames);
  bool out_grew = false, words_grew = false;
  const size_t coef_pitch = coef_padded_blocks(nb) * 64;   // sector-interleaved: whole groups of 4 blocks
  CU(L->coef.Reserve(f * coef_pitch * sizeof(int16_t)));
  CU(L->nzmask.Reserve(f * nb * sizeof(uint8_t) + 64));
  CU(L->words.Reserve(f * plan.stream_words * 4 + 64, true, &words_grew));   // zeroed once, then self-cleaning
  // zeroed once at allocation: the head copy of a single encode (FinishSingle) reads a fixed 64 KB+ of the
  // slot whatever the JPEG's size, and initcheck rightly flags reading bytes nobody ever wrote
  CU(L->out.Reserve(f * plan.out_capacity, true, &out_grew));
  CU(L->state.Reserve(f * (plan.nb_tiles + plan.ff_tiles) * sizeof(unsigned long long)));
  if (plan.trellis) CU(L->perm.Reserve(f * nb * sizeof(uint32_t)));
  GroupBuffers& gb = L->gb;
  const bool relayout = out_grew || gb.out_pitch != plan.out_capacity || gb.words_pitch != plan.stream_words ||
                        L->group_capacity != frames;
  gb.coef = L->coef.as<int16_t>();
  gb.coef_pitch = coef_pitch;
  gb.nzmask = L->nzmask.as<uint8_t>();
  gb.mask_pitch = nb;
  gb.words = L->words.as<uint32_t>();
  gb.words_pitch = plan.stream_words;
  gb.out = L->out.as<uint8_t>();
  gb.out_pitch = plan.out_capacity;
  gb.bit_state = L->state.as<unsigned long long>();
  gb.bit_state_pitch = plan.nb_tiles;
  gb.ff_state = gb.bit_state + f * plan.nb_tiles;
  gb.ff_state_pitch = plan.ff_tiles;
  SmallLayout* s = L->d_small();
  gb.info = s->info;
  gb.tabs = s->tabs;
  gb.qtabs = s->qtabs;
  gb.hist = &s->hist[0][0];
  gb.freq = &s->freq[0][0];
  gb.quant = &s->quant[0][0][0];
  gb.dc_init = nullptr;
  gb.bit_offsets = nullptr;
  if (relayout) L->header_valid = 0;   // out slots moved: headers must be sent again
  L->group_capacity = frames;
  return SJB_OK;
}


bool StagerEnabled() {
  static const bool on = [] {
    const char* e = getenv("SJPEG_B200_STAGER");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}

size_t PixSlotBytes(const Plan& plan, long long stride) {
  const size_t row_bytes = static_cast<size_t>(plan.pstep) * plan.g.width;
  const size_t astride = static_cast<size_t>(stride < 0 ? -stride : stride);
  const size_t h = plan.g.height;
  const size_t need =
      (astride <= 2 * row_bytes + 64) ? astride * (h - 1) + row_bytes : ((row_bytes + 15) & ~size_t(15)) * h;
  return (need + 64 + 255) & ~static_cast<size_t>(255);
}

int ReservePix(sjb_context* ctx, Lane* L, const Plan& plan, long long stride, int frames) {
  L->pix_pitch = PixSlotBytes(plan, stride);
  CU(L->pix.Reserve(L->pix_pitch * frames));
  return SJB_OK;
}

// Copies one host picture into slot `slot` of the lane's pixel buffer; returns the device
// address of its row 0.
int UploadPicture(sjb_context* ctx, Lane* L, const uint8_t* pix, const Plan& plan, long long stride, int slot,
                  const uint8_t** d_row0, long long* d_stride, cudaStream_t copy_stream = nullptr) {
  if (copy_stream == nullptr) copy_stream = L->stream;
  const size_t row_bytes = static_cast<size_t>(plan.pstep) * plan.g.width;
  const long long astride = stride < 0 ? -stride : stride;
  const int h = plan.g.height;
  uint8_t* base = L->pix.as<uint8_t>() + slot * L->pix_pitch;
  const uint8_t* lowest = (stride < 0) ? pix + stride * (h - 1) : pix;
  if (static_cast<size_t>(astride) <= 2 * row_bytes + 64) {
    // one contiguous span, stride kept (sign included)
    const size_t span = static_cast<size_t>(astride) * (h - 1) + row_bytes;
    bool staged = false;
    // From 4 MB up (below, the driver's own pageable path is as good).  Until the helper threads stayed
    // awake between uploads a lone call only paid off from 4K pictures up: waking them cost ~0.25 ms.
    const size_t min_bytes = HostStager::kMinBytes;
    if (span >= min_bytes && StagerEnabled()) {
      // malloc()ed memory (the normal case behind SjpegEncode): copy out through the pinned ring
      // with helper threads instead of the driver's single-threaded pageable path
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, lowest) == cudaSuccess && attr.type == cudaMemoryTypeUnregistered) {
        const cudaError_t e = ctx->stager.Upload(base, lowest, span, copy_stream);
        if (e == cudaSuccess) staged = true;
        else if (e != cudaErrorNotSupported) CU(e);
      } else {
        cudaGetLastError();
      }
    }
    if (!staged) CU(cudaMemcpyAsync(base, lowest, span, cudaMemcpyHostToDevice, copy_stream));
    *d_row0 = base + ((stride < 0) ? static_cast<size_t>(astride) * (h - 1) : 0);
    *d_stride = stride;
  } else {
    // sparse rows: gather into a tight pitch
    const size_t pitch = (row_bytes + 15) & ~static_cast<size_t>(15);
    CU(cudaMemcpy2DAsync(base, pitch, lowest, static_cast<size_t>(astride), row_bytes, h, cudaMemcpyHostToDevice,
                         copy_stream));
    *d_row0 = base + ((stride < 0) ? pitch * (h - 1) : 0);
    *d_stride = (stride < 0) ? -static_cast<long long>(pitch) : static_cast<long long>(pitch);
  }
  return SJB_OK;
}

void FillFrameSet(const Plan& plan, long long stride, FrameSet* fs) {
  memset(fs, 0, sizeof(*fs));
  const FrameGeometry& g = plan.g;
  fs->stride = stride;
  fs->width = g.width;
  fs->height = g.height;
  fs->yuv_mode = g.yuv_mode;
  fs->pix_fmt = plan.p.pix_fmt;
  fs->mcus_x = g.mcus_x;
  fs->mcus_y = g.mcus_y;
  fs->mcu_blocks = g.mcu_blocks;
  fs->luma_blocks = g.luma_blocks;
  fs->blocks_per_frame = static_cast<unsigned>(g.nb_blocks());
}

void LaunchF1(Lane* L, const FrameSet& fs, const FrameGeometry& g, bool raw, const QuantTabs& qt) {
  const int mx_full = g.width / g.mcu_size, my_full = g.height / g.mcu_size;
  int mx_fast = 0, my_fast = 0;
  if (F1FastEligible(fs)) {
    // row bytes of a partial tile must be a multiple of 16: any MCU count for packed 4:2:0 (48 or 64
    // bytes per MCU), an even one otherwise (8 / 24 bytes per MCU and plane)
    mx_fast = (g.yuv_mode == kYuv420 && !fs.planar) ? mx_full : (mx_full & ~1);
    my_fast = my_full;
    if (mx_fast == 0) my_fast = 0;
    if (my_fast == 0) mx_fast = 0;
  }
  if (mx_fast > 0) {
    LaunchF1Fast(fs, mx_fast, 0, my_fast, raw, qt, L->gb, L->stream);
    ++L->launches;
  }
  if (mx_fast < g.mcus_x && my_fast > 0) {        // columns right of the fast region
    LaunchF1Generic(fs, mx_fast, 0, g.mcus_x, my_fast, raw, qt, L->gb, L->stream);
    ++L->launches;
  }
  if (my_fast < g.mcus_y) {                        // rows below it
    LaunchF1Generic(fs, 0, my_fast, g.mcus_x, g.mcus_y, raw, qt, L->gb, L->stream);
    ++L->launches;
  }
}

bool MakeQuantTabs(const Plan& plan, uint8_t quant[2][64], uint8_t min_quant[2][64], QuantTabs* qt) {
  memcpy(quant, plan.p.quant, 128);
  memcpy(min_quant, plan.p.min_quant, 128);
  for (int i = 0; i < 2; ++i) {
    if (!FinalizeQuantizer(quant[i], min_quant[i], plan.p.q_bias, &qt->m[i])) return false;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// Device pipeline for one group whose pixels are already in device memory (fs.pix[]).
//
// Methods that optimise the Huffman tables need the host between kernels -- symbol counts -> optimal
// tables (entropy.cc:254-444); the adaptive ones did too until the histogram analysis
// (histogram.cc:126-315) became a kernel, and still want their matrices on the host for the DQT -- so a group
// is a small state machine (GroupJob): every phase enqueues work on the lane's stream, ends with an
// event, and the next phase starts by waiting for that event.  A single encode just runs the phases
// back to back; the batch entry points interleave the phases of several groups (one per lane) so
// that the calling thread is enqueueing the next group's upload and kernels while an earlier group's
// counters travel back, instead of idling in cudaStreamSynchronize with the GPU (and the PCIe link)
// waiting for it.  Per-picture host work of a phase is spread over the context's worker threads.
//   stage 0 -> [adaptive: F1 raw, H1, A1 (analysis on the device), Q1 or T1, matrices D2H | else: F1 quantised]
//   stage 1 -> (wait) adaptive methods without optimised tables: the matrices for the DQT have arrived
//   stage 2 -> (wait) OptimalHuffSpec per picture (the matrices arrived with the symbol counts)
//   finish  -> code tables + headers H2D, E, S, sizes D2H          => stage 3 (all enqueued)
// ---------------------------------------------------------------------------------------------
// brackets one kernel stage of a timed group with events (sjb_last_stage_timings)
struct StageTimer {
  Lane* L;
  int idx;
  bool on;
  cudaStream_t st;
  StageTimer(Lane* lane, int i, bool timed, cudaStream_t stream = nullptr)
      : L(lane), idx(i), on(timed), st(stream ? stream : lane->stream) {
    if (on) cudaEventRecord(L->kev[idx][0], st);
  }
  ~StageTimer() {
    if (on) {
      cudaEventRecord(L->kev[idx][1], st);
      L->kev_set[idx] = true;
    }
  }
};

// SJB_ES_PRIORITY=1: the entropy and stuffing kernels of a group go to a highest-priority stream of
// the lane, so that the block scheduler gives SM slots freed by another group's F1 to them first and
// the two kinds of kernels share SMs (F1 is issue bound, E latency bound) instead of following each other.
bool EsPriority() {
  static const bool v = [] {
    const char* e = getenv("SJB_ES_PRIORITY");
    return e != nullptr && atoi(e) != 0;
  }();
  return v;
}

struct GroupJob {
  Lane* L = nullptr;
  FrameSet fs;
  const Plan* plan = nullptr;
  bool timed = false;
  int stage = 3;                   // 3 = nothing pending
  uint8_t quant0[2][64], min_quant[2][64];
  QuantTabs qt;
  HuffSpec def_spec[4];            // dc0 dc1 ac0 ac1
  CodeTabs def_tabs;
  std::vector<uint8_t> quant;      // [n][2][64]: the matrices that go into each picture's DQT
  std::vector<HuffSpec> spec;      // [n][4]
  std::vector<CodeTabs> tabs;      // [n]
};

// Pinned staging (tables, headers) is only rewritten when its content changes, and then only after
// the stream has drained: an earlier async copy may still be reading it.  (Right after a phase's
// event wait the stream is idle and the synchronisation returns at once.)
int UploadCodeTabs(sjb_context* ctx, GroupJob* J) {
  Lane* L = J->L;
  HostScratch* H = L->host;
  const int n = J->fs.frames;
  bool same = L->tabs_valid >= n;
  for (int f = 0; same && f < n; ++f) same = memcmp(&H->tabs[f], &J->tabs[f], sizeof(CodeTabs)) == 0;
  if (same) return SJB_OK;
  CU(cudaStreamSynchronize(L->stream));
  for (int f = 0; f < n; ++f) H->tabs[f] = J->tabs[f];
  LaunchCopySmall(L->d_small()->tabs, H->tabs, n * sizeof(CodeTabs), L->stream);   // not the copy engine: kernels.cu
  L->tabs_valid = n;
  return SJB_OK;
}

// the default code tables as the trellis' rate model (enc.cc:334): a device copy of their own, so that
// the per-picture tables of the coder (d_small()->tabs) are only ever written with what the coder uses
int EnsureDefaultTabs(Lane* L, const CodeTabs& def_tabs, cudaStream_t st) {
  if (L->def_tabs_valid) return SJB_OK;
  for (int f = 0; f < kMaxGroup; ++f) L->host->def_tabs[f] = def_tabs;
  LaunchCopySmall(L->d_small()->def_tabs, L->host->def_tabs, sizeof(L->host->def_tabs), st);
  L->def_tabs_valid = true;
  return SJB_OK;
}

void FillAqParams(const Plan& plan, const uint8_t quant0[2][64], const uint8_t min_quant[2][64], AqParams* ap) {
  memcpy(ap->quant0, quant0, 128);
  memcpy(ap->min_quant, min_quant, 128);
  ap->qdelta_max[0] = plan.p.qdelta_max_luma;
  ap->qdelta_max[1] = plan.p.qdelta_max_chroma;
  ap->q_bias = plan.p.q_bias;
  ap->nb_comps = plan.g.nb_comps;
}

// Adaptive methods, everything after the histogram kernel up to the quantised coefficients, stream
// ordered on st: analysis on the device (A1), the matrices for the DQT on their way to the host
// (host->quant, host->aq_fail: valid once the stream has passed this point), Q1 or T1.
int EnqueueAdaptiveQuantise(sjb_context* ctx, Lane* L, const FrameSet& fs, const Plan& plan, const uint8_t quant0[2][64],
                            const uint8_t min_quant[2][64], const CodeTabs& def_tabs, bool run_kernels, cudaStream_t st) {
  SmallLayout* D = L->d_small();
  const int n = fs.frames;
  AqParams ap;
  FillAqParams(plan, quant0, min_quant, &ap);
  LaunchAnalyseHistograms(n, L->gb, ap, &D->aq_fit[0][0][0], D->aq_fail, st);      // also clears aq_fail[0..n)
  static_assert(offsetof(SmallLayout, aq_fail) == offsetof(SmallLayout, quant) + sizeof(D->quant) &&
                offsetof(HostScratch, aq_fail) == offsetof(HostScratch, quant) + sizeof(D->quant), "quant and aq_fail are copied together");
  CU(cudaMemcpyAsync(L->host->quant, D->quant, sizeof(D->quant) + sizeof(D->aq_fail), cudaMemcpyDeviceToHost, st));
  L->launches += kAnalyseLaunches;
  if (!run_kernels) return SJB_OK;
  if (plan.trellis) {
    RC(EnsureDefaultTabs(L, def_tabs, st));
    GroupBuffers gb = L->gb;
    gb.tabs = D->def_tabs;
    LaunchTrellis(fs, gb, nullptr, &D->trellis_sort[0][0], L->perm.as<uint32_t>(), plan.g.nb_blocks(), st);
    L->launches += kTrellisLaunches;
  } else {
    LaunchRequantize(fs, L->gb, nullptr, st);
    L->launches += 1;
  }
  CU(cudaGetLastError());
  return SJB_OK;
}

int FinishGroup(sjb_context* ctx, GroupJob* J) {
  Lane* L = J->L;
  const FrameSet& fs = J->fs;
  const FrameGeometry& g = J->plan->g;
  const int n = fs.frames;
  HostScratch* H = L->host;
  SmallLayout* D = L->d_small();
  const GroupBuffers& gb = L->gb;
  RC(UploadCodeTabs(ctx, J));
  // headers (host) -> start of each out slot
  {
    std::vector<std::vector<uint8_t> > headers(n);
    bool same = L->header_valid >= n;
    for (int f = 0; f < n; ++f) {
      headers[f].reserve(1024);
      AppendHeaders(g, reinterpret_cast<const uint8_t(*)[64]>(&J->quant[f * 128]), &J->spec[f * 4], &headers[f]);
      if (headers[f].size() > kHeaderReserve) return SJB_ERR_ARG;
      same = same && L->header_len[f] == headers[f].size() &&
             memcmp(H->header[f], headers[f].data(), headers[f].size()) == 0;
    }
    if (!same) {
      CU(cudaStreamSynchronize(L->stream));
      for (int f = 0; f < n; ++f) {
        memcpy(H->header[f], headers[f].data(), headers[f].size());
        L->header_len[f] = static_cast<unsigned>(headers[f].size());
        LaunchCopySmall(gb.out + f * gb.out_pitch, H->header[f], headers[f].size(), L->stream);
      }
      L->header_valid = n;
    }
  }
  L->words_dirty = true;
  cudaStream_t es = L->stream;
  if (EsPriority()) {
    if (L->hi == nullptr) {
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU(cudaStreamCreateWithPriority(&L->hi, cudaStreamNonBlocking, hi));
      CU(cudaEventCreateWithFlags(&L->hi_ev, cudaEventDisableTiming));
    }
    es = L->hi;
    CU(cudaEventRecord(L->hi_ev, L->stream));
    CU(cudaStreamWaitEvent(es, L->hi_ev, 0));
  }
  CU(cudaMemsetAsync(L->state.ptr, 0, L->state.bytes, es));   // look-back descriptors
  { StageTimer t(L, 4, J->timed, es); LaunchEntropyPack(fs, gb, es); }
  {
    StuffArgs sa;
    memset(&sa, 0, sizeof(sa));
    for (int f = 0; f < n; ++f) {
      sa.header_len[f] = L->header_len[f];
      sa.flags[f] = kStuffFirst | kStuffLast;
    }
    StageTimer t(L, 5, J->timed, es);
    LaunchStuff(fs, gb, sa, es);
  }
  if (es != L->stream) {
    CU(cudaEventRecord(L->hi_ev, es));
    CU(cudaStreamWaitEvent(L->stream, L->hi_ev, 0));
  }
  L->launches += 2;
  CU(cudaGetLastError());
  L->words_dirty = false;   // the stuffing kernel zeroes every word it consumed
  if (J->timed) CU(cudaEventRecord(L->ev[2], L->stream));
  CU(cudaMemcpyAsync(H->info, D->info, n * sizeof(StreamInfo), cudaMemcpyDeviceToHost, L->stream));
  J->stage = 3;
  return SJB_OK;
}

// coefficients are quantised: statistics pass for the optimised tables, or straight to the coder
int AfterQuantise(sjb_context* ctx, GroupJob* J) {
  Lane* L = J->L;
  const int n = J->fs.frames;
  if (!J->plan->optimize) return FinishGroup(ctx, J);
  // enc.cc:344-374 : symbol statistics -> optimal tables
  SmallLayout* D = L->d_small();
  CU(cudaMemsetAsync(D->freq, 0, n * sizeof(D->freq[0]), L->stream));
  { StageTimer t(L, 3, J->timed); LaunchSymbolStats(J->fs, L->gb, L->stream); }
  L->launches += 1;
  CU(cudaMemcpyAsync(L->host->freq, D->freq, n * sizeof(D->freq[0]), cudaMemcpyDeviceToHost, L->stream));
  CU(cudaEventRecord(L->phase_ev, L->stream));
  J->stage = 2;
  return SJB_OK;
}

int StartGroup(sjb_context* ctx, GroupJob* J, Lane* L, const FrameSet& fs, const Plan& plan, bool timed) {
  J->L = L;
  J->fs = fs;
  J->plan = &plan;
  J->timed = timed;
  J->stage = 3;
  const FrameGeometry& g = plan.g;
  const int n = fs.frames;
  SmallLayout* D = L->d_small();
  if (!MakeQuantTabs(plan, J->quant0, J->min_quant, &J->qt)) {
    ctx->err = "quantiser entry outside the range of the fused quantise form";
    return SJB_ERR_ARG;
  }
  if (L->words_dirty) {
    CU(cudaMemsetAsync(L->words.ptr, 0, L->words.bytes, L->stream));
    L->words_dirty = false;
  }
  if (timed) {
    CU(cudaEventRecord(L->ev[0], L->stream));
    for (bool& b : L->kev_set) b = false;
    L->timed_frames = n;
  }
  // default Huffman tables (entropy.cc:31-86)
  for (int i = 0; i < 4; ++i) DefaultHuffSpec(i >= 2, i & 1, &J->def_spec[i]);
  memset(&J->def_tabs, 0, sizeof(J->def_tabs));
  for (int c = 0; c < 2; ++c) {
    CodesFromSpec(J->def_spec[c], J->def_tabs.dc[c]);
    CodesFromSpec(J->def_spec[2 + c], J->def_tabs.ac[c]);
  }
  J->tabs.assign(n, J->def_tabs);
  J->quant.resize(static_cast<size_t>(n) * 128);
  for (int f = 0; f < n; ++f) memcpy(&J->quant[f * 128], J->quant0, 128);
  J->spec.resize(static_cast<size_t>(n) * 4);
  for (int f = 0; f < n; ++f) for (int i = 0; i < 4; ++i) J->spec[f * 4 + i] = J->def_spec[i];

  if (plan.adaptive) {
    // enc.cc:425-429 : histogram pass over unquantised coefficients; the matrices are re-derived ON
    // THE DEVICE (A1) and the coefficients quantised with them without the host in between.  The host
    // only needs the matrices for the DQT segment: they travel back behind the kernels and are picked
    // up at the next wait the method has anyway (symbol counts), or at one of their own (stage 1).
    { StageTimer t(L, 0, timed); LaunchF1(L, fs, g, /*raw=*/true, J->qt); }
    CU(cudaMemsetAsync(D->hist, 0, n * sizeof(D->hist[0]), L->stream));
    { StageTimer t(L, 1, timed); LaunchHistogram(fs, L->gb, L->stream); }
    L->launches += 1;
    if (timed) CU(cudaEventRecord(L->ev[1], L->stream));
    {
      StageTimer t(L, 2, timed);
      RC(EnqueueAdaptiveQuantise(ctx, L, fs, plan, J->quant0, J->min_quant, J->def_tabs, true, L->stream));
    }
    if (!plan.optimize) {
      CU(cudaEventRecord(L->phase_ev, L->stream));
      J->stage = 1;
      return SJB_OK;
    }
    return AfterQuantise(ctx, J);
  }
  { StageTimer t(L, 0, timed); LaunchF1(L, fs, g, /*raw=*/false, J->qt); }
  if (timed) CU(cudaEventRecord(L->ev[1], L->stream));
  return AfterQuantise(ctx, J);
}

// runs the next host phase of a job (waits for the device first); no-op once everything is enqueued
int AdvanceGroup(sjb_context* ctx, GroupJob* J) {
  if (J->stage >= 3) return SJB_OK;
  Lane* L = J->L;
  const Plan& plan = *J->plan;
  const FrameGeometry& g = plan.g;
  const int n = J->fs.frames;
  HostScratch* H = L->host;
  SmallLayout* D = L->d_small();
  CU(cudaEventSynchronize(L->phase_ev));
  if (plan.adaptive) {
    // the matrices A1 derived, for the headers
    for (int f = 0; f < n; ++f) {
      if (H->aq_fail[f]) {
        ctx->err = "adapted quantiser entry outside the range of the fused quantise form";
        return SJB_ERR_ARG;
      }
      memcpy(&J->quant[f * 128], H->quant[f], 128);
    }
  }
  if (J->stage == 1) return FinishGroup(ctx, J);
  // stage 2: optimal tables from the symbol counts
  const int nb_tables = (g.nb_comps == 1) ? 1 : 2;
  ctx->pool.ParallelFor(n, [&](int f) {
    for (int c = 0; c < nb_tables; ++c) {
      OptimalHuffSpec(H->freq[f] + 272 * c + 256, 12, &J->spec[f * 4 + c]);
      OptimalHuffSpec(H->freq[f] + 272 * c, 256, &J->spec[f * 4 + 2 + c]);
      CodesFromSpec(J->spec[f * 4 + c], J->tabs[f].dc[c]);
      CodesFromSpec(J->spec[f * 4 + 2 + c], J->tabs[f].ac[c]);
    }
  });
  return FinishGroup(ctx, J);
}

// one group, start to "everything enqueued": on return the JPEGs are (asynchronously) in the lane's
// out slots, header included, and host->info[] will hold their sizes once the stream has drained
int EncodeGroup(sjb_context* ctx, Lane* L, const FrameSet& fs, const Plan& plan, bool timed) {
  GroupJob job;
  RC(StartGroup(ctx, &job, L, fs, plan, timed));
  while (job.stage < 3) RC(AdvanceGroup(ctx, &job));
  return SJB_OK;
}

// Encoder::LoopScan (dichotomy.cc:113-205) for one picture on lane L.
int EncodeSearch(sjb_context* ctx, Lane* L, const FrameSet& fs, const Plan& plan, sjb_search* S) {
  const FrameGeometry& g = plan.g;
  HostScratch* H = L->host;
  SmallLayout* D = L->d_small();
  GroupBuffers gb = L->gb;
  const size_t nb = g.nb_blocks();
  const int nb_tables = (g.nb_comps == 1) ? 1 : 2;
  const int passes = std::min(20, std::max(1, S->passes));

  uint8_t quant[2][64], min_quant[2][64];
  QuantTabs qt;
  if (!MakeQuantTabs(plan, quant, min_quant, &qt)) return SJB_ERR_ARG;
  if (L->words_dirty) {
    CU(cudaMemsetAsync(L->words.ptr, 0, L->words.bytes, L->stream));
    L->words_dirty = false;
  }
  CU(cudaEventRecord(L->ev[0], L->stream));
  // unquantised coefficients -> raw buffer (kept for all passes)
  CU(L->raw.Reserve(coef_padded_blocks(nb) * 64 * sizeof(int16_t)));
  {
    GroupBuffers graw = gb;
    graw.coef = L->raw.as<int16_t>();
    Lane tmp_view;   // LaunchF1 only reads gb / stream / launches
    tmp_view.gb = graw;
    tmp_view.stream = L->stream;
    LaunchF1(&tmp_view, fs, g, /*raw=*/true, qt);
    L->launches += tmp_view.launches;
    tmp_view.stream = nullptr;
    if (plan.adaptive) {                                           // dichotomy.cc:117-121
      CU(cudaMemsetAsync(D->hist, 0, sizeof(D->hist[0]), L->stream));
      LaunchHistogram(fs, graw, L->stream);
      CU(cudaMemcpyAsync(H->hist, D->hist, sizeof(D->hist[0]), cudaMemcpyDeviceToHost, L->stream));
    }
  }
  CU(cudaEventRecord(L->ev[1], L->stream));
  CU(cudaStreamSynchronize(L->stream));
  const int16_t* raw = L->raw.as<int16_t>();

  // current Huffman tables: default until statistics are compiled (entropy.cc:84-86)
  HuffSpec spec[4];
  for (int i = 0; i < 4; ++i) DefaultHuffSpec(i >= 2, i & 1, &spec[i]);
  CodeTabs tabs;
  memset(&tabs, 0, sizeof(tabs));
  auto rebuild_codes = [&]() {
    for (int c = 0; c < nb_tables; ++c) {
      CodesFromSpec(spec[c], tabs.dc[c]);
      CodesFromSpec(spec[2 + c], tabs.ac[c]);
    }
  };
  for (int c = 0; c < 2; ++c) {
    CodesFromSpec(spec[c], tabs.dc[c]);
    CodesFromSpec(spec[2 + c], tabs.ac[c]);
  }
  auto upload_tabs = [&]() -> int {
    CU(cudaStreamSynchronize(L->stream));
    H->tabs[0] = tabs;
    CU(cudaMemcpyAsync(D->tabs, H->tabs, sizeof(CodeTabs), cudaMemcpyHostToDevice, L->stream));
    L->tabs_valid = 0;
    return SJB_OK;
  };
  auto upload_quant = [&]() -> int {
    CU(cudaStreamSynchronize(L->stream));
    H->qtabs[0] = qt;
    memcpy(H->quant[0], quant, 128);
    CU(cudaMemcpyAsync(D->qtabs, H->qtabs, sizeof(QuantTabs), cudaMemcpyHostToDevice, L->stream));
    CU(cudaMemcpyAsync(D->quant, H->quant, 128, cudaMemcpyHostToDevice, L->stream));
    return SJB_OK;
  };
  // StoreRunLevels (dichotomy.cc:80-111): quantise raw -> gb.coef with the current matrices (the
  // trellis prices with the AC codes in force), gather symbol statistics when optimising
  auto store_run_levels = [&]() -> int {
    RC(upload_quant());
    if (plan.trellis) {
      RC(upload_tabs());
      LaunchTrellis(fs, gb, raw, &D->trellis_sort[0][0], L->perm.as<uint32_t>(), nb, L->stream);
      L->launches += kTrellisLaunches - 1;
    } else {
      LaunchRequantize(fs, gb, raw, L->stream);
    }
    ++L->launches;
    if (plan.optimize) {
      CU(cudaMemsetAsync(D->freq, 0, sizeof(D->freq[0]), L->stream));
      LaunchSymbolStats(fs, gb, L->stream);
      ++L->launches;
      CU(cudaMemcpyAsync(H->freq, D->freq, sizeof(D->freq[0]), cudaMemcpyDeviceToHost, L->stream));
      CU(cudaStreamSynchronize(L->stream));
    }
    CU(cudaGetLastError());
    return SJB_OK;
  };
  auto compile_stats = [&]() {                                     // entropy.cc:432-444
    for (int c = 0; c < nb_tables; ++c) {
      OptimalHuffSpec(H->freq[0] + 272 * c + 256, 12, &spec[c]);
      OptimalHuffSpec(H->freq[0] + 272 * c, 256, &spec[2 + c]);
    }
  };
  auto header_bits = [&]() -> size_t {                             // dichotomy.cc:210-243
    size_t size = 20 + S->header_extra_bytes;
    size += nb_tables * 65 + 2 + 2;
    size += 8 + 3 * g.nb_comps + 2;
    size += 6 + 2 * g.nb_comps + 2;
    size += 2;
    for (int c = 0; c < nb_tables; ++c) {
      size += 2 + 3 + 16 + spec[c].nb_syms;
      size += 2 + 3 + 16 + spec[2 + c].nb_syms;
    }
    return size * 8;
  };

  uint8_t opt_quants[2][64];
  memcpy(opt_quants, quant, 128);
  float best = 0.f, best_result = 0.f;
  int best_pass = 0;
  bool last_is_best = false;
  for (int p = 0; p < passes; ++p) {
    if (S->begin_pass) S->begin_pass(S->user, p);
    for (int c = 0; c < 2; ++c) {
      S->next_matrix(S->user, c, quant[c]);
      if (!FinalizeQuantizer(quant[c], min_quant[c], plan.p.q_bias, &qt.m[c])) return SJB_ERR_ARG;
    }
    if (plan.adaptive) {
      AnalyseHistograms(H->hist[0], g.nb_comps, quant, min_quant, plan.p.qdelta_max_luma, plan.p.qdelta_max_chroma);
      for (int c = (g.nb_comps > 1 ? 1 : 0); c >= 0; --c) {
        if (!FinalizeQuantizer(quant[c], min_quant[c], plan.p.q_bias, &qt.m[c])) return SJB_ERR_ARG;
      }
    }
    float result;
    if (S->for_size) {
      RC(store_run_levels());
      size_t size;
      if (plan.optimize) {
        compile_stats();                                           // dichotomy.cc:150-153
        rebuild_codes();                                           // ComputeSize: InitCodes(false)
        size = header_bits();
        for (int q = 0; q < nb_tables; ++q) {                      // EntropySize, entropy.cc:230-245
          const uint32_t* f = H->freq[0] + 272 * q;
          for (int len = 0; len < 12; ++len) if (f[256 + len]) size += static_cast<size_t>(f[256 + len]) * ((tabs.dc[q][len] & 0xff) + len);
          for (int sym = 0; sym < 256; ++sym) if (f[sym]) size += static_cast<size_t>(f[sym]) * ((tabs.ac[q][sym] & 0xff) + (sym & 0x0f));
        }
      } else {
        // exact bit count incl. the 0x00 after every complete 0xFF byte (BitCounter, bit_writer.h:292-365)
        rebuild_codes();
        size = header_bits();
        RC(upload_tabs());
        CU(cudaMemsetAsync(L->state.ptr, 0, L->state.bytes, L->stream));
        LaunchEntropyPack(fs, gb, L->stream);
        StuffArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.header_len[0] = kHeaderReserve;          // scratch area of the out slot
        sa.flags[0] = kStuffFirst | kStuffKeepWords;
        LaunchStuff(fs, gb, sa, L->stream);
        L->launches += 2;
        CU(cudaMemcpyAsync(H->info, D->info, sizeof(StreamInfo), cudaMemcpyDeviceToHost, L->stream));
        CU(cudaMemsetAsync(L->words.ptr, 0, L->words.bytes, L->stream));
        CU(cudaStreamSynchronize(L->stream));
        CU(cudaGetLastError());
        L->header_valid = 0;
        size += static_cast<size_t>(H->info[0].total_bits) + 8 * static_cast<size_t>(H->info[0].stuffed_bytes);
      }
      result = size / 8.f;
    } else {
      RC(upload_quant());
      unsigned long long* d_err = reinterpret_cast<unsigned long long*>(&D->info[1]);   // spare record
      CU(cudaMemsetAsync(d_err, 0, sizeof(unsigned long long), L->stream));
      LaunchQuantError(fs, gb, raw, d_err - 0, L->stream);
      ++L->launches;
      CU(cudaMemcpyAsync(&H->info[1], &D->info[1], sizeof(StreamInfo), cudaMemcpyDeviceToHost, L->stream));
      CU(cudaStreamSynchronize(L->stream));
      CU(cudaGetLastError());
      const unsigned long long err = H->info[1].total_bits;
      const unsigned long long size = 64ull * nb;
      result = (err > 0 && size > 0) ? static_cast<float>(4.3429448f * log(size / (err / 255. / 255.))) : 99.f;
    }
    last_is_best = (p == 0 || fabs(result - S->target) < best);
    if (last_is_best) {
      memcpy(opt_quants, quant, 128);
      best = static_cast<float>(fabs(result - S->target));
      best_pass = p;
      best_result = result;
    }
    if (S->update(S->user, result)) break;
  }
  S->best_pass = best_pass;
  S->best_result = best_result;

  // transfer back the kept matrices and finish (dichotomy.cc:176-203)
  memcpy(quant, opt_quants, 128);
  for (int c = 0; c < 2; ++c) {
    if (!FinalizeQuantizer(quant[c], min_quant[c], plan.p.q_bias, &qt.m[c])) return SJB_ERR_ARG;
  }
  if (!S->for_size || !last_is_best) {
    RC(store_run_levels());
    if (plan.optimize) compile_stats();
  }
  rebuild_codes();
  RC(upload_tabs());
  std::vector<uint8_t> header;
  AppendHeaders(g, quant, spec, &header);
  if (header.size() > kHeaderReserve) return SJB_ERR_ARG;
  CU(cudaStreamSynchronize(L->stream));
  memcpy(H->header[0], header.data(), header.size());
  L->header_len[0] = static_cast<unsigned>(header.size());
  L->header_valid = 0;
  CU(cudaMemcpyAsync(gb.out, H->header[0], header.size(), cudaMemcpyHostToDevice, L->stream));
  L->words_dirty = true;
  CU(cudaMemsetAsync(L->state.ptr, 0, L->state.bytes, L->stream));
  LaunchEntropyPack(fs, gb, L->stream);
  StuffArgs sa;
  memset(&sa, 0, sizeof(sa));
  sa.header_len[0] = L->header_len[0];
  sa.flags[0] = kStuffFirst | kStuffLast;
  LaunchStuff(fs, gb, sa, L->stream);
  L->launches += 2;
  CU(cudaGetLastError());
  L->words_dirty = false;
  CU(cudaEventRecord(L->ev[2], L->stream));
  CU(cudaMemcpyAsync(H->info, D->info, sizeof(StreamInfo), cudaMemcpyDeviceToHost, L->stream));
  return SJB_OK;
}

int FinishTimings(sjb_context* ctx, Lane* L) {
  CU(cudaEventSynchronize(L->ev[2]));
  CU(cudaEventElapsedTime(&L->ms_total, L->ev[0], L->ev[2]));
  CU(cudaEventElapsedTime(&L->ms_f1, L->ev[0], L->ev[1]));
  L->ms_entropy = L->ms_total - L->ms_f1;
  return SJB_OK;
}

// single host/device picture on lane 0, result left in out slot 0
int EncodeSingle(sjb_context* ctx, const uint8_t* pix, int pix_on_device, long long stride, const Plan& plan,
                 bool timed) {
  Lane* L = &ctx->lanes[0];
  RC(ReserveLane(ctx, L, plan, 1));
  FrameSet fs;
  FillFrameSet(plan, stride, &fs);
  fs.frames = 1;
  fs.pix[0] = pix;
  if (!pix_on_device) {
    RC(ReservePix(ctx, L, plan, stride, 1));
    RC(UploadPicture(ctx, L, pix, plan, stride, 0, &fs.pix[0], &fs.stride));
  }
  sjb_search* search = ctx->search;
  ctx->search = nullptr;
  const int rc = search ? EncodeSearch(ctx, L, fs, plan, search) : EncodeGroup(ctx, L, fs, plan, timed);
  if (rc != SJB_OK) L->words_dirty = true;
  return rc;
}

// copies a host plane (rows of row_bytes, any stride sign) to dst; returns the device row 0
int UploadPlane(sjb_context* ctx, Lane* L, uint8_t* dst, const uint8_t* src, long long stride, size_t row_bytes,
                int rows, const uint8_t** d_row0, long long* d_stride) {
  const size_t pitch = (row_bytes + 15) & ~static_cast<size_t>(15);
  const long long astride = stride < 0 ? -stride : stride;
  const uint8_t* lowest = (stride < 0) ? src + stride * (rows - 1) : src;
  CU(cudaMemcpy2DAsync(dst, pitch, lowest, static_cast<size_t>(astride), row_bytes, rows, cudaMemcpyHostToDevice,
                       L->stream));
  *d_row0 = dst + ((stride < 0) ? pitch * (rows - 1) : 0);
  *d_stride = (stride < 0) ? -static_cast<long long>(pitch) : static_cast<long long>(pitch);
  return SJB_OK;
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

int sjb_context_create(int device, sjb_context** out) try {
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
} SJB_NOTHROW_END

void sjb_context_destroy(sjb_context* ctx) {
  if (ctx == nullptr) return;
  cudaSetDevice(ctx->device);
  for (auto& L : ctx->lanes) DestroyLane(&L);
  for (DeviceBuffer* b : {&ctx->sharp_scratch, &ctx->sharp_planes, &ctx->sharp_tabs, &ctx->risk_table, &ctx->risk_sums,
                          &ctx->batch_pix, &ctx->batch_planes}) {
    b->Release();
  }
  for (auto& b : ctx->sharp_scratch_n) b.Release();
  for (auto& st : ctx->sharp_stream) if (st) cudaStreamDestroy(st);
  if (ctx->risk_host) cudaFreeHost(ctx->risk_host);
  if (ctx->head_copy) cudaFreeHost(ctx->head_copy);
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
void* sjb_host_alloc_wc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void sjb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

namespace {
int EncodeAutoOrSharp(sjb_context* ctx, const uint8_t* pix, int pix_on_device, int width, int height,
                      long long stride, const sjb_params* params);
int FinishSingle(sjb_context* ctx, uint8_t* out, int out_on_device, size_t out_capacity, size_t* out_size);
}  // namespace

int sjb_encode(sjb_context* ctx, const uint8_t* pix, int pix_on_device, int width, int height,
               long long stride, const sjb_params* params, uint8_t* out, int out_on_device,
               size_t out_capacity, size_t* out_size) try {
  if (ctx == nullptr || pix == nullptr || out_size == nullptr) return SJB_ERR_ARG;
  *out_size = 0;
  ctx->lanes[0].last_size = 0;
  ctx->err.clear();
  if (params != nullptr && (params->yuv_mode == SJB_YUV_AUTO || params->yuv_mode == SJB_YUV_SHARP)) {
    RC(EncodeAutoOrSharp(ctx, pix, pix_on_device, width, height, stride, params));
  } else {
    Plan plan;
    RC(MakePlan(width, height, stride, params, &plan));
    CU(cudaSetDevice(ctx->device));
    ctx->lanes[0].launches = 0;
    RC(EncodeSingle(ctx, pix, pix_on_device, stride, plan, /*timed=*/true));
  }
  return FinishSingle(ctx, out, out_on_device, out_capacity, out_size);
} SJB_NOTHROW_END

int sjb_fetch_output(sjb_context* ctx, uint8_t* out, int out_on_device, size_t out_capacity) try {
  if (ctx == nullptr || out == nullptr) return SJB_ERR_ARG;
  Lane* L = &ctx->lanes[0];
  const size_t size = L->last_size;
  if (size == 0) return SJB_ERR_ARG;
  if (size > out_capacity) return SJB_ERR_CAPACITY;
  if (!out_on_device && size <= ctx->head_valid) {      // already on the host (FinishSingle)
    memcpy(out, ctx->head_copy, size);
    return SJB_OK;
  }
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(out, L->out.ptr, size, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                     L->stream));
  CU(cudaStreamSynchronize(L->stream));
  return SJB_OK;
} SJB_NOTHROW_END

}  // extern "C"

namespace {

// Pictures that come from host memory arrive at PCIe speed (one 4K picture per 0.47 ms), far
// slower than the kernels consume them: small groups start computing as soon as their pictures
// have landed and leave a short tail after the last copy.  Device-resident batches use the
// large groups that suit the kernels.
// Measured (B200, pinned host memory, Gpix/s): 16 x 4K  group 1/2/4/8 = 18.06/18.02/17.89/17.67 (m0),
// 17.81/17.68/17.58/17.21 (m4); 64 x 1080p = 17.30/17.41/17.43/17.33 (m0), 12.5/15.2/17.0/16.8 (m4: a
// group of two small pictures does not cover the host phase of the optimised tables).  So: two large
// pictures or four small ones per group.  SJB_HOST_BATCH_GROUP overrides.
int HostBatchGroup(const Plan& plan) {
  static const int forced = [] {
    const char* e = getenv("SJB_HOST_BATCH_GROUP");
    const int x = e ? atoi(e) : 0;
    return x > 0 ? x : 0;
  }();
  if (forced) return forced;
  return (static_cast<long long>(plan.g.width) * plan.g.height <= (4LL << 20)) ? 4 : 2;
}

// The batch pipeline over the context's lanes, independent of where the pictures come from:
// fill(k, L, &fs) makes the pictures [k * B, k * B + fs.frames) of the batch available on lane L
// (device pointers and strides into fs; uploads go on L's stream).  Picture j of the batch is
// delivered to out[index ? index[j] : j] / sizes[...].
//
// Software pipeline: group k runs on lane k % nl.  At step k the lane is first freed (the group
// that used it is finished and drained: sizes read, bytes copied out), group k is started on it
// (upload + first kernels enqueued), and then every older group still in flight is advanced by
// one host phase, oldest first -- so the waits for histograms / symbol counts of one group happen
// while the uploads and kernels of the younger ones are already queued.
template <class Fill>
int RunBatch(sjb_context* ctx, const Plan& plan, long long stride, int n, int B, Fill fill, const int* index,
             uint8_t* const* out, int out_on_device, size_t out_capacity, size_t* sizes, int max_lanes = kMaxLanes) {
  const int groups = (n + B - 1) / B;
  const int nl = std::min<int>(std::min<int>(kMaxLanes, max_lanes), std::max(1, groups));
  for (int l = 0; l < nl; ++l) {
    RC(InitLane(ctx, &ctx->lanes[l]));
    RC(ReserveLane(ctx, &ctx->lanes[l], plan, B));
  }
  std::vector<GroupJob> jobs(nl);
  auto where = [&](int j) { return index ? index[j] : j; };
  for (int j = 0; j < n; ++j) sizes[where(j)] = 0;
  auto drain = [&](int k) -> int {
    Lane* L = &ctx->lanes[k % nl];
    GroupJob* J = &jobs[k % nl];
    while (J->stage < 3) RC(AdvanceGroup(ctx, J));
    CU(cudaStreamSynchronize(L->stream));
    int rc = SJB_OK;
    for (int f = 0; f < B && k * B + f < n; ++f) {
      const int i = where(k * B + f);
      const size_t size = static_cast<size_t>(L->host->info[f].out_size);
      sizes[i] = size;
      if (out[i] == nullptr || size > out_capacity) {
        rc = SJB_ERR_CAPACITY;
        continue;
      }
      CU(cudaMemcpyAsync(out[i], L->gb.out + f * L->gb.out_pitch, size,
                         out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, L->stream));
    }
    return rc;
  };
  int first_err = SJB_OK;
  // A failure in the middle of the pipeline must not leave copies in flight that read the caller's
  // (pinned) inputs or write the caller's outputs after we return: every lane is drained first.
  auto abort_batch = [&](int rc) -> int {
    for (int l = 0; l < nl; ++l) {
      if (ctx->lanes[l].stream) cudaStreamSynchronize(ctx->lanes[l].stream);
      ctx->lanes[l].words_dirty = true;
    }
    cudaGetLastError();
    return rc;
  };
  for (int k = 0; k < groups; ++k) {
    Lane* L = &ctx->lanes[k % nl];
    if (k >= nl) {
      const int rc = drain(k - nl);
      if (rc != SJB_OK && rc != SJB_ERR_CAPACITY) return abort_batch(rc);
      if (rc != SJB_OK && first_err == SJB_OK) first_err = rc;
    }
    FrameSet fs;
    FillFrameSet(plan, stride, &fs);
    fs.frames = std::min(B, n - k * B);
    int rc = fill(k, L, &fs);
    if (rc == SJB_OK) rc = StartGroup(ctx, &jobs[k % nl], L, fs, plan, false);
    for (int j = std::max(0, k - nl + 1); j < k && rc == SJB_OK; ++j) rc = AdvanceGroup(ctx, &jobs[j % nl]);
    if (rc != SJB_OK) return abort_batch(rc);
  }
  for (int k = std::max(0, groups - nl); k < groups; ++k) {
    const int rc = drain(k);
    if (rc != SJB_OK && rc != SJB_ERR_CAPACITY) return abort_batch(rc);
    if (rc != SJB_OK && first_err == SJB_OK) first_err = rc;
  }
  for (int l = 0; l < nl; ++l) {
    if (cudaStreamSynchronize(ctx->lanes[l].stream) != cudaSuccess) {
      ctx->err = "cudaStreamSynchronize failed at the end of the batch";
      return abort_batch(SJB_ERR_CUDA);
    }
  }
  return first_err;
}

int EnsureSharpTabs(sjb_context* ctx, const uint32_t** g2l, const uint32_t** l2g);
int EnsureScoreTable(sjb_context* ctx);

// SJB_YUV_AUTO / SJB_YUV_SHARP for a batch of packed RGB pictures (what SjpegCompress and a default
// EncoderParam ask for, api.cc:83-101, encoders.cc:546-568).  All pictures are made resident first;
// AUTO runs the riskiness analyser on every picture (one launch each, one wait for all the sums) and
// the batch is then split by the mode each picture got: the 4:2:0 / 4:4:4 / 4:0:0 subsets go through
// the ordinary batch pipeline from their device copies, the SHARP subset is converted -- up to
// kSharpStreams (8) pictures at a time, each conversion on its own stream with its own scratch, because a
// single conversion only occupies a fifth of the SMs -- into planes that then go through the planar
// 4:2:0 pipeline.
int EncodeBatchAutoOrSharp(sjb_context* ctx, int n, const uint8_t* const* pix, int pix_on_device, int width, int height,
                           long long stride, const sjb_params* params, uint8_t* const* out, int out_on_device,
                           size_t out_capacity, size_t* sizes) {
  if (params->pix_fmt != SJB_PIX_RGB) return SJB_ERR_ARG;   // api.cc:208,235: callers convert to RGB first
  if (width > 65535 || height > 65535) return SJB_ERR_ARG;
  sjb_params p420 = *params;
  p420.yuv_mode = SJB_YUV_420;
  Plan plan;
  RC(MakePlan(width, height, stride, &p420, &plan));
  RC(InitLane(ctx, &ctx->lanes[0]));
  Lane* L0 = &ctx->lanes[0];
  // 1. every picture resident on the device
  std::vector<const uint8_t*> d_pix(n);
  long long d_stride = stride;
  if (pix_on_device) {
    for (int i = 0; i < n; ++i) d_pix[i] = pix[i];
  } else {
    const size_t slot = PixSlotBytes(plan, stride);
    CU(ctx->batch_pix.Reserve(slot * n));
    Lane view;                                   // UploadPicture reads pix / pix_pitch / stream only
    view.pix = ctx->batch_pix;
    view.pix_pitch = slot;
    view.stream = L0->stream;
    int rc = SJB_OK;
    for (int i = 0; i < n && rc == SJB_OK; ++i) rc = UploadPicture(ctx, &view, pix[i], plan, stride, i, &d_pix[i], &d_stride);
    view.pix = DeviceBuffer();                   // not ours to release
    view.stream = nullptr;
    RC(rc);
  }
  // 2. the mode of every picture
  std::vector<int> mode(n, params->yuv_mode);
  if (params->yuv_mode == SJB_YUV_AUTO) {
    RC(EnsureScoreTable(ctx));
    CU(ctx->risk_sums.Reserve(static_cast<size_t>(n) * 3 * sizeof(unsigned long long)));
    if (ctx->risk_host_cap < static_cast<size_t>(n) * 3) {
      if (ctx->risk_host) cudaFreeHost(ctx->risk_host);
      ctx->risk_host = nullptr;
      ctx->risk_host_cap = 0;
      CU(cudaMallocHost(reinterpret_cast<void**>(&ctx->risk_host), static_cast<size_t>(n) * 3 * sizeof(unsigned long long)));
      ctx->risk_host_cap = static_cast<size_t>(n) * 3;
    }
    unsigned long long* d_sums = ctx->risk_sums.as<unsigned long long>();
    for (int i = 0; i < n; ++i) {
      CU(LaunchRiskiness(d_pix[i], d_stride, width, height, ctx->risk_table.as<uint8_t>(), d_sums + 3 * i, ctx->sm_count, L0->stream));
    }
    CU(cudaMemcpyAsync(ctx->risk_host, d_sums, static_cast<size_t>(n) * 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       L0->stream));
    CU(cudaStreamSynchronize(L0->stream));
    for (int i = 0; i < n; ++i) {
      float risk;
      mode[i] = RiskinessDecision(ctx->risk_host[3 * i], ctx->risk_host[3 * i + 1], ctx->risk_host[3 * i + 2], width, height, &risk);
    }
  } else {
    CU(cudaStreamSynchronize(L0->stream));     // the uploads: the lanes below read the pictures from their own streams
  }
  // 3. the block-pipeline modes, one sub-batch each
  int result = SJB_OK;
  for (int m : {SJB_YUV_420, SJB_YUV_444, SJB_YUV_400}) {
    std::vector<int> idx;
    for (int i = 0; i < n; ++i) if (mode[i] == m) idx.push_back(i);
    if (idx.empty()) continue;
    sjb_params pm = *params;
    pm.yuv_mode = m;
    Plan pl;
    RC(MakePlan(width, height, d_stride, &pm, &pl));
    const int cnt = static_cast<int>(idx.size());
    const int B = std::max(1, std::min(pl.group, cnt));
    auto fill = [&](int k, Lane*, FrameSet* fs) -> int {
      fs->stride = d_stride;
      for (int f = 0; f < fs->frames; ++f) fs->pix[f] = d_pix[idx[k * B + f]];
      return SJB_OK;
    };
    const int rc = RunBatch(ctx, pl, d_stride, cnt, B, fill, idx.data(), out, out_on_device, out_capacity, sizes);
    if (rc != SJB_OK && rc != SJB_ERR_CAPACITY) return rc;
    if (rc != SJB_OK) result = rc;
  }
  // 4. the SHARP subset: conversions on kSharpStreams streams, then the planar 4:2:0 pipeline
  std::vector<int> idx;
  for (int i = 0; i < n; ++i) if (mode[i] == SJB_YUV_SHARP) idx.push_back(i);
  if (!idx.empty()) {
    const int cnt = static_cast<int>(idx.size());
    const uint32_t *g2l, *l2g;
    RC(EnsureSharpTabs(ctx, &g2l, &l2g));
    // planes with 16-byte aligned pitches, so that the planar fast kernel takes them
    const size_t ypitch = (static_cast<size_t>(width) + 15) & ~size_t(15);
    const size_t cw = (static_cast<size_t>(width) + 1) / 2, ch = (static_cast<size_t>(height) + 1) / 2;
    const bool tight = ypitch == static_cast<size_t>(width) && (cw & 15) == 0;   // the conversion writes tightly packed planes
    const size_t ybytes = (static_cast<size_t>(width) * height + 255) & ~size_t(255), cbytes = (cw * ch + 255) & ~size_t(255);
    const size_t per = ybytes + 2 * cbytes;
    (void)tight;
    CU(ctx->batch_planes.Reserve(per * cnt));
    SharpLayout lay;
    const size_t scratch = (width > 4 && height > 4) ? SharpScratchBytes(width, height, &lay) : 0;
    const int ns = std::min<int>(kSharpStreams, cnt);
    for (int s = 0; s < ns; ++s) {
      if (ctx->sharp_stream[s] == nullptr) CU(cudaStreamCreateWithFlags(&ctx->sharp_stream[s], cudaStreamNonBlocking));
      if (scratch) CU(ctx->sharp_scratch_n[s].Reserve(scratch));
    }
    for (int j = 0; j < cnt; ++j) {
      uint8_t* y = ctx->batch_planes.as<uint8_t>() + per * j;
      int launches = 0;
      CU(LaunchSharpYuv(d_pix[idx[j]], d_stride, width, height, ctx->sharp_scratch_n[j % ns].as<uint8_t>(), g2l, l2g, y, y + ybytes,
                        y + ybytes + cbytes, ctx->sharp_stream[j % ns], &launches));
    }
    for (int s = 0; s < ns; ++s) CU(cudaStreamSynchronize(ctx->sharp_stream[s]));
    sjb_params pm = *params;
    pm.yuv_mode = SJB_YUV_420;
    Plan pl;
    RC(MakePlan(width, height, 3LL * width, &pm, &pl));
    const int B = std::max(1, std::min(pl.group, cnt));
    auto fill = [&](int k, Lane*, FrameSet* fs) -> int {
      fs->planar = 1;
      fs->uv_step = 1;
      fs->stride = width;
      fs->stride_u = fs->stride_v = static_cast<long long>(cw);
      for (int f = 0; f < fs->frames; ++f) {
        const uint8_t* y = ctx->batch_planes.as<uint8_t>() + per * (k * B + f);
        fs->pix[f] = y;
        fs->pix_u[f] = y + ybytes;
        fs->pix_v[f] = y + ybytes + cbytes;
      }
      return SJB_OK;
    };
    const int rc = RunBatch(ctx, pl, width, cnt, B, fill, idx.data(), out, out_on_device, out_capacity, sizes);
    if (rc != SJB_OK && rc != SJB_ERR_CAPACITY) return rc;
    if (rc != SJB_OK) result = rc;
  }
  return result;
}

}  // namespace

extern "C" {

int sjb_encode_batch(sjb_context* ctx, int n, const uint8_t* const* pix, int pix_on_device, int width,
                     int height, long long stride, const sjb_params* params, uint8_t* const* out,
                     int out_on_device, size_t out_capacity, size_t* sizes) try {
  if (ctx == nullptr || pix == nullptr || out == nullptr || sizes == nullptr || n < 0 || params == nullptr) return SJB_ERR_ARG;
  ctx->err.clear();
  ctx->lanes[0].last_size = 0;
  for (int i = 0; i < n; ++i) if (pix[i] == nullptr) return SJB_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  if (params->yuv_mode == SJB_YUV_AUTO || params->yuv_mode == SJB_YUV_SHARP) {
    if (n == 0) return SJB_OK;
    return EncodeBatchAutoOrSharp(ctx, n, pix, pix_on_device, width, height, stride, params, out, out_on_device,
                                  out_capacity, sizes);
  }
  Plan plan;
  RC(MakePlan(width, height, stride, params, &plan));
  if (n == 0) return SJB_OK;
  const int B = std::max(1, std::min(pix_on_device ? plan.group : std::min(plan.group, HostBatchGroup(plan)), n));
  if (!pix_on_device) {
    const int groups = (n + B - 1) / B;
    for (int l = 0; l < std::min<int>(kHostLanes, groups); ++l) {
      RC(InitLane(ctx, &ctx->lanes[l]));
      RC(ReservePix(ctx, &ctx->lanes[l], plan, stride, B));
    }
  }
  auto fill = [&](int k, Lane* L, FrameSet* fs) -> int {
    for (int f = 0; f < fs->frames; ++f) {
      fs->pix[f] = pix[k * B + f];
      if (!pix_on_device) {
        long long ds = stride;
        RC(UploadPicture(ctx, L, pix[k * B + f], plan, stride, f, &fs->pix[f], &ds));
        fs->stride = ds;
      }
    }
    return SJB_OK;
  };
  return RunBatch(ctx, plan, stride, n, B, fill, nullptr, out, out_on_device, out_capacity, sizes,
                  pix_on_device ? kMaxLanes : kHostLanes);
} SJB_NOTHROW_END

// A batch of planar / semi-planar pictures of one geometry (sjb_encode_planar for the layouts):
// y[i] / u[i] / v[i] are the planes of picture i, strides common to the batch.
int sjb_encode_planar_batch(sjb_context* ctx, int n, const uint8_t* const* y, long long y_stride, const uint8_t* const* u,
                            long long u_stride, const uint8_t* const* v, long long v_stride, int uv_step, int on_device,
                            int width, int height, const sjb_params* params, uint8_t* const* out, int out_on_device,
                            size_t out_capacity, size_t* sizes) try {
  if (ctx == nullptr || y == nullptr || out == nullptr || sizes == nullptr || params == nullptr || n < 0) return SJB_ERR_ARG;
  ctx->err.clear();
  ctx->lanes[0].last_size = 0;
  const int mode = params->yuv_mode;
  if (mode != SJB_YUV_420 && mode != SJB_YUV_444 && mode != SJB_YUV_400) return SJB_ERR_ARG;
  if (width <= 0 || height <= 0) return SJB_ERR_ARG;
  if (uv_step != 1 && !(uv_step == 2 && mode == SJB_YUV_420)) return SJB_ERR_ARG;
  const int cw = (mode == SJB_YUV_420) ? (width + 1) / 2 : width, ch = (mode == SJB_YUV_420) ? (height + 1) / 2 : height;
  auto absll = [](long long a) { return a < 0 ? -a : a; };
  if (absll(y_stride) < width) return SJB_ERR_ARG;
  if (mode != SJB_YUV_400) {
    if (u == nullptr || v == nullptr) return SJB_ERR_ARG;
    if (absll(u_stride) < static_cast<long long>(uv_step) * cw || absll(v_stride) < static_cast<long long>(uv_step) * cw)
      return SJB_ERR_ARG;
  }
  for (int i = 0; i < n; ++i) {
    if (y[i] == nullptr || (mode != SJB_YUV_400 && (u[i] == nullptr || v[i] == nullptr))) return SJB_ERR_ARG;
  }
  sjb_params p = *params;
  p.pix_fmt = SJB_PIX_RGB;
  Plan plan;
  RC(MakePlan(width, height, 3LL * width, &p, &plan));
  if (n == 0) return SJB_OK;
  CU(cudaSetDevice(ctx->device));
  const int B = std::max(1, std::min(on_device ? plan.group : std::min(plan.group, HostBatchGroup(plan)), n));
  const size_t ypitch = (static_cast<size_t>(width) + 15) & ~size_t(15);
  const size_t crow = static_cast<size_t>(mode == SJB_YUV_420 ? uv_step : 1) * cw;
  const size_t cpitch = (crow + 15) & ~size_t(15);
  const size_t ybytes = ypitch * height, cbytes = (mode == SJB_YUV_400) ? 0 : cpitch * ch;
  const size_t slot = (ybytes + 2 * cbytes + 255) & ~size_t(255);
  if (!on_device) {
    const int groups = (n + B - 1) / B;
    for (int l = 0; l < std::min<int>(kHostLanes, groups); ++l) {
      RC(InitLane(ctx, &ctx->lanes[l]));
      CU(ctx->lanes[l].pix.Reserve(slot * B));
    }
  }
  auto fill = [&](int k, Lane* L, FrameSet* fs) -> int {
    fs->planar = 1;
    fs->uv_step = (mode == SJB_YUV_420) ? uv_step : 1;
    fs->stride = y_stride;
    fs->stride_u = u_stride;
    fs->stride_v = v_stride;
    for (int f = 0; f < fs->frames; ++f) {
      const int i = k * B + f;
      fs->pix[f] = y[i];
      fs->pix_u[f] = (mode == SJB_YUV_400) ? nullptr : u[i];
      fs->pix_v[f] = (mode == SJB_YUV_400) ? nullptr : v[i];
      if (on_device) continue;
      uint8_t* base = L->pix.as<uint8_t>() + slot * f;
      RC(UploadPlane(ctx, L, base, y[i], y_stride, width, height, &fs->pix[f], &fs->stride));
      if (mode == SJB_YUV_400) continue;
      if (uv_step == 2) {
        const uint8_t* first = (u[i] < v[i]) ? u[i] : v[i];
        const uint8_t* d0;
        long long ds;
        RC(UploadPlane(ctx, L, base + ybytes, first, u_stride, crow, ch, &d0, &ds));
        fs->pix_u[f] = d0 + (u[i] - first);
        fs->pix_v[f] = d0 + (v[i] - first);
        fs->stride_u = fs->stride_v = ds;
      } else {
        RC(UploadPlane(ctx, L, base + ybytes, u[i], u_stride, crow, ch, &fs->pix_u[f], &fs->stride_u));
        RC(UploadPlane(ctx, L, base + ybytes + cbytes, v[i], v_stride, crow, ch, &fs->pix_v[f], &fs->stride_v));
      }
    }
    return SJB_OK;
  };
  return RunBatch(ctx, plan, y_stride, n, B, fill, nullptr, out, out_on_device, out_capacity, sizes,
                  on_device ? kMaxLanes : kHostLanes);
} SJB_NOTHROW_END

}  // extern "C"

namespace {

// planar input up to the end of the kernel sequence (no wait, nothing copied out)
int EncodePlanarOnLane(sjb_context* ctx, const uint8_t* y, long long y_stride, const uint8_t* u, long long u_stride,
                       const uint8_t* v, long long v_stride, int uv_step, int on_device, int width, int height,
                       const sjb_params* params) {
  const int mode = params->yuv_mode;
  if (mode != SJB_YUV_420 && mode != SJB_YUV_444 && mode != SJB_YUV_400) return SJB_ERR_ARG;
  if (width <= 0 || height <= 0) return SJB_ERR_ARG;
  if (uv_step != 1 && !(uv_step == 2 && mode == SJB_YUV_420)) return SJB_ERR_ARG;
  const int cw = (mode == SJB_YUV_420) ? (width + 1) / 2 : width, ch = (mode == SJB_YUV_420) ? (height + 1) / 2 : height;
  auto absll = [](long long a) { return a < 0 ? -a : a; };
  if (absll(y_stride) < width) return SJB_ERR_ARG;                       // encoders.cc:349,428,499
  if (mode != SJB_YUV_400) {
    if (u == nullptr || v == nullptr) return SJB_ERR_ARG;
    if (absll(u_stride) < static_cast<long long>(uv_step) * cw || absll(v_stride) < static_cast<long long>(uv_step) * cw)
      return SJB_ERR_ARG;
  }
  sjb_params p = *params;
  p.pix_fmt = SJB_PIX_RGB;
  Plan plan;
  RC(MakePlan(width, height, 3LL * width, &p, &plan));   // geometry / method; the stride is not used
  CU(cudaSetDevice(ctx->device));
  Lane* L = &ctx->lanes[0];
  RC(ReserveLane(ctx, L, plan, 1));
  FrameSet fs;
  FillFrameSet(plan, y_stride, &fs);
  fs.frames = 1;
  fs.planar = 1;
  fs.uv_step = (mode == SJB_YUV_420) ? uv_step : 1;
  fs.pix[0] = y;
  fs.pix_u[0] = u;
  fs.pix_v[0] = v;
  fs.stride_u = u_stride;
  fs.stride_v = v_stride;
  if (!on_device) {
    const size_t ypitch = (static_cast<size_t>(width) + 15) & ~size_t(15);
    const size_t crow = static_cast<size_t>(fs.uv_step) * cw;
    const size_t cpitch = (crow + 15) & ~size_t(15);
    const size_t ybytes = ypitch * height, cbytes = cpitch * ch;
    CU(L->pix.Reserve(ybytes + 2 * cbytes + 256));
    uint8_t* base = L->pix.as<uint8_t>();
    RC(UploadPlane(ctx, L, base, y, y_stride, width, height, &fs.pix[0], &fs.stride));
    if (mode != SJB_YUV_400) {
      if (uv_step == 2) {
        // one interleaved plane: copy it once, from whichever of u / v comes first in memory
        const uint8_t* first = (u < v) ? u : v;
        const uint8_t* d0;
        long long ds;
        RC(UploadPlane(ctx, L, base + ybytes, first, u_stride, crow, ch, &d0, &ds));
        fs.pix_u[0] = d0 + (u - first);
        fs.pix_v[0] = d0 + (v - first);
        fs.stride_u = fs.stride_v = ds;
      } else {
        RC(UploadPlane(ctx, L, base + ybytes, u, u_stride, crow, ch, &fs.pix_u[0], &fs.stride_u));
        RC(UploadPlane(ctx, L, base + ybytes + cbytes, v, v_stride, crow, ch, &fs.pix_v[0], &fs.stride_v));
      }
    }
  }
  sjb_search* search = ctx->search;
  ctx->search = nullptr;
  const int rc = search ? EncodeSearch(ctx, L, fs, plan, search) : EncodeGroup(ctx, L, fs, plan, /*timed=*/true);
  if (rc != SJB_OK) L->words_dirty = true;
  return rc;
}

// waits for lane 0, reports the size and copies the JPEG out (shared tail of the single-picture calls)
int FinishSingle(sjb_context* ctx, uint8_t* out, int out_on_device, size_t out_capacity, size_t* out_size) {
  Lane* L = &ctx->lanes[0];
  ctx->head_valid = 0;
  size_t head = 0;
  if (out == nullptr) {
    if (ctx->head_copy == nullptr && cudaMallocHost(reinterpret_cast<void**>(&ctx->head_copy), kHeadCopyBytes) != cudaSuccess) {
      cudaGetLastError();
      ctx->head_copy = nullptr;
    }
    if (ctx->head_copy != nullptr) {
      // about 2 bits per pixel's worth: what a typical file needs, small enough not to weigh on small pictures
      head = std::min<size_t>(std::min<size_t>(kHeadCopyBytes, std::max<size_t>(64 << 10, L->gb.mask_pitch * 16)),
                              L->gb.out_pitch);
      CU(cudaMemcpyAsync(ctx->head_copy, L->out.ptr, head, cudaMemcpyDeviceToHost, L->stream));
    }
  }
  CU(cudaStreamSynchronize(L->stream));
  ctx->head_valid = head;
  FinishTimings(ctx, L);
  const size_t size = static_cast<size_t>(L->host->info[0].out_size);
  *out_size = size;
  L->last_size = size;
  if (out == nullptr || size > out_capacity) return SJB_ERR_CAPACITY;
  CU(cudaMemcpyAsync(out, L->out.ptr, size, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                     L->stream));
  CU(cudaStreamSynchronize(L->stream));
  return SJB_OK;
}

// ---- whole-picture passes in front of the block pipeline (sharp.cu) -----------------------------
std::mutex g_table_mutex;
std::vector<uint8_t> g_score_table;      // process-wide copy of the caller's 343 x 343 table
int g_score_table_version = 0;           // bumped on every change; 0 = never set
bool g_score_table_env_checked = false;

// Where the 343 x 343 riskiness table comes from when the caller has not handed one over
// (sjb_set_score_table), tried once, in this order:
//   1. SJPEG_B200_SCORE_TABLE=<file holding the 117649 bytes>
//   2. sjpeg::kSharpnessScore of a reference libsjpeg loaded in the same process (dlsym: the
//      drop-in scenario of INTEGRATION.md, where our kernels replace the reference's loops)
//   3. sjpeg_score_table.bin next to libsjpeg_b200.so -- written at BUILD time by csrc/Makefile from
//      the reference's score_7.cc where that source tree is present (the table is a generated data
//      asset of the reference, jpeg_tools.cc:204-206; it is not reproduced in this repository)
// Without any of them sjb_riskiness / SJB_YUV_AUTO fail (SJB_ERR_ARG): never a silent other mode.
bool ReadTableFile(const char* path, std::vector<uint8_t>* t) {
  FILE* f = fopen(path, "rb");
  if (f == nullptr) return false;
  t->resize(kRiskTableBytes);
  const bool ok = fread(t->data(), 1, t->size(), f) == t->size();
  fclose(f);
  return ok;
}

void LoadScoreTableFromEnvLocked() {
  if (g_score_table_env_checked) return;
  g_score_table_env_checked = true;
  if (!g_score_table.empty()) return;
  std::vector<uint8_t> t;
  const char* path = getenv("SJPEG_B200_SCORE_TABLE");
  bool ok = path != nullptr && ReadTableFile(path, &t);
  if (!ok) {
    const void* sym = dlsym(RTLD_DEFAULT, "_ZN5sjpeg15kSharpnessScoreE");
    if (sym != nullptr) {
      t.assign(static_cast<const uint8_t*>(sym), static_cast<const uint8_t*>(sym) + kRiskTableBytes);
      ok = true;
    }
  }
  if (!ok) {
    Dl_info info;
    if (dladdr(reinterpret_cast<const void*>(&sjb_version), &info) != 0 && info.dli_fname != nullptr) {
      std::string dir(info.dli_fname);
      const size_t slash = dir.rfind('/');
      dir = (slash == std::string::npos) ? std::string(".") : dir.substr(0, slash);
      ok = ReadTableFile((dir + "/sjpeg_score_table.bin").c_str(), &t);
    }
  }
  if (ok) {
    g_score_table.swap(t);
    ++g_score_table_version;
  }
}

int EnsureSharpTabs(sjb_context* ctx, const uint32_t** g2l, const uint32_t** l2g) {
  if (ctx->sharp_tabs.ptr == nullptr) {
    uint32_t host[1024 + 34];
    MakeSharpGammaTables(host, host + 1024);
    CU(ctx->sharp_tabs.Reserve(sizeof(host)));
    CU(cudaMemcpy(ctx->sharp_tabs.ptr, host, sizeof(host), cudaMemcpyHostToDevice));
  }
  *g2l = ctx->sharp_tabs.as<uint32_t>();
  *l2g = *g2l + 1024;
  return SJB_OK;
}

int EnsureScoreTable(sjb_context* ctx) {
  std::lock_guard<std::mutex> lock(g_table_mutex);
  LoadScoreTableFromEnvLocked();
  if (g_score_table.empty()) {
    ctx->err = "no riskiness score table: call sjb_set_score_table(), set SJPEG_B200_SCORE_TABLE, load the reference "
               "libsjpeg alongside, or build with the reference sources present (sjpeg_score_table.bin)";
    return SJB_ERR_ARG;
  }
  if (ctx->risk_table_version != g_score_table_version) {
    CU(ctx->risk_table.Reserve(kRiskTableBytes));
    CU(cudaMemcpy(ctx->risk_table.ptr, g_score_table.data(), kRiskTableBytes, cudaMemcpyHostToDevice));
    ctx->risk_table_version = g_score_table_version;
  }
  if (ctx->risk_host == nullptr) {
    CU(cudaMallocHost(reinterpret_cast<void**>(&ctx->risk_host), 3 * sizeof(unsigned long long)));
    ctx->risk_host_cap = 3;
    CU(ctx->risk_sums.Reserve(256));
  }
  return SJB_OK;
}

// device address of a packed RGB picture: as given, or uploaded into lane 0's pixel buffer
int ResidentRgb(sjb_context* ctx, const uint8_t* rgb, int on_device, int width, int height, long long stride,
                const uint8_t** d_rgb, long long* d_stride) {
  if (rgb == nullptr || width <= 0 || height <= 0) return SJB_ERR_ARG;
  const long long astride = stride < 0 ? -stride : stride;
  if (astride < 3LL * width) return SJB_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  Lane* L = &ctx->lanes[0];
  *d_rgb = rgb;
  *d_stride = stride;
  if (!on_device) {
    sjb_params p;
    sjb_params_default(&p, 75.f, 0, SJB_YUV_420);
    Plan plan;
    RC(MakePlan(width, height, stride, &p, &plan));
    RC(ReservePix(ctx, L, plan, stride, 1));
    RC(UploadPicture(ctx, L, rgb, plan, stride, 0, d_rgb, d_stride));
  }
  return SJB_OK;
}

// riskiness of a device-resident picture (jpeg_tools.cc:177-236); waits for the three sums
int RiskinessOnDevice(sjb_context* ctx, const uint8_t* d_rgb, long long d_stride, int width, int height, int* mode,
                      float* risk) {
  RC(EnsureScoreTable(ctx));
  Lane* L = &ctx->lanes[0];
  unsigned long long* d_sums = ctx->risk_sums.as<unsigned long long>();
  CU(LaunchRiskiness(d_rgb, d_stride, width, height, ctx->risk_table.as<uint8_t>(), d_sums, ctx->sm_count, L->stream));
  L->launches += 1;
  CU(cudaMemcpyAsync(ctx->risk_host, d_sums, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, L->stream));
  CU(cudaStreamSynchronize(L->stream));
  *mode = RiskinessDecision(ctx->risk_host[0], ctx->risk_host[1], ctx->risk_host[2], width, height, risk);
  return SJB_OK;
}

// sharp conversion of a device-resident picture into the context's plane buffer (stream-ordered)
int SharpOnDevice(sjb_context* ctx, const uint8_t* d_rgb, long long d_stride, int width, int height, uint8_t** y,
                  uint8_t** u, uint8_t** v) {
  const uint32_t *g2l, *l2g;
  RC(EnsureSharpTabs(ctx, &g2l, &l2g));
  Lane* L = &ctx->lanes[0];
  const size_t cw = (static_cast<size_t>(width) + 1) / 2, ch = (static_cast<size_t>(height) + 1) / 2;
  const size_t ybytes = (static_cast<size_t>(width) * height + 255) & ~size_t(255);
  const size_t cbytes = (cw * ch + 255) & ~size_t(255);
  CU(ctx->sharp_planes.Reserve(ybytes + 2 * cbytes));
  SharpLayout lay;
  if (width > 4 && height > 4) CU(ctx->sharp_scratch.Reserve(SharpScratchBytes(width, height, &lay)));
  *y = ctx->sharp_planes.as<uint8_t>();
  *u = *y + ybytes;
  *v = *u + cbytes;
  int launches = 0;
  CU(LaunchSharpYuv(d_rgb, d_stride, width, height, ctx->sharp_scratch.as<uint8_t>(), g2l, l2g, *y, *u, *v, L->stream,
                    &launches));
  L->launches += launches;
  return SJB_OK;
}

// SJB_YUV_AUTO / SJB_YUV_SHARP for packed RGB (EncoderFactory, encoders.cc:546-568): the pixels are
// uploaded once; AUTO asks the analyser, SHARP converts on the device and feeds the planar encoder.
int EncodeAutoOrSharp(sjb_context* ctx, const uint8_t* pix, int pix_on_device, int width, int height,
                      long long stride, const sjb_params* params) {
  if (params->pix_fmt != SJB_PIX_RGB) return SJB_ERR_ARG;   // api.cc:208,235: callers convert to RGB first
  if (width > 65535 || height > 65535) return SJB_ERR_ARG;
  Lane* L = &ctx->lanes[0];
  const uint8_t* d_rgb;
  long long d_stride;
  RC(ResidentRgb(ctx, pix, pix_on_device, width, height, stride, &d_rgb, &d_stride));
  L->launches = 0;
  sjb_params p = *params;
  if (p.yuv_mode == SJB_YUV_AUTO) {
    float risk;
    RC(RiskinessOnDevice(ctx, d_rgb, d_stride, width, height, &p.yuv_mode, &risk));
  }
  if (p.yuv_mode != SJB_YUV_SHARP) {
    Plan plan;
    RC(MakePlan(width, height, d_stride, &p, &plan));
    return EncodeSingle(ctx, d_rgb, /*pix_on_device=*/1, d_stride, plan, /*timed=*/true);
  }
  uint8_t *y, *u, *v;
  RC(SharpOnDevice(ctx, d_rgb, d_stride, width, height, &y, &u, &v));
  p.yuv_mode = SJB_YUV_420;
  const long long cw = (width + 1) / 2;
  return EncodePlanarOnLane(ctx, y, width, u, cw, v, cw, 1, /*on_device=*/1, width, height, &p);
}

}  // namespace

extern "C" {

int sjb_encode_planar(sjb_context* ctx, const uint8_t* y, long long y_stride, const uint8_t* u,
                      long long u_stride, const uint8_t* v, long long v_stride, int uv_step, int on_device,
                      int width, int height, const sjb_params* params, uint8_t* out, int out_on_device,
                      size_t out_capacity, size_t* out_size) try {
  if (ctx == nullptr || y == nullptr || out_size == nullptr || params == nullptr) return SJB_ERR_ARG;
  *out_size = 0;
  ctx->lanes[0].last_size = 0;
  ctx->err.clear();
  ctx->lanes[0].launches = 0;
  RC(EncodePlanarOnLane(ctx, y, y_stride, u, u_stride, v, v_stride, uv_step, on_device, width, height, params));
  return FinishSingle(ctx, out, out_on_device, out_capacity, out_size);
} SJB_NOTHROW_END

int sjb_set_score_table(const uint8_t* table, size_t size) try {
  std::lock_guard<std::mutex> lock(g_table_mutex);
  if (table == nullptr) {
    g_score_table.clear();
    ++g_score_table_version;
    g_score_table_env_checked = true;   // an explicit clear is not undone by the environment
    return SJB_OK;
  }
  if (size != static_cast<size_t>(kRiskTableBytes)) return SJB_ERR_ARG;
  g_score_table.assign(table, table + size);
  ++g_score_table_version;
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_has_score_table(void) try {
  std::lock_guard<std::mutex> lock(g_table_mutex);
  LoadScoreTableFromEnvLocked();
  return g_score_table.empty() ? 0 : 1;
} SJB_NOTHROW_END

int sjb_riskiness(sjb_context* ctx, const uint8_t* rgb, int rgb_on_device, int width, int height, long long stride,
                  int* yuv_mode, float* risk) try {
  if (ctx == nullptr || yuv_mode == nullptr) return SJB_ERR_ARG;
  ctx->err.clear();
  const uint8_t* d_rgb;
  long long d_stride;
  RC(ResidentRgb(ctx, rgb, rgb_on_device, width, height, stride, &d_rgb, &d_stride));
  return RiskinessOnDevice(ctx, d_rgb, d_stride, width, height, yuv_mode, risk);
} SJB_NOTHROW_END

int sjb_sharp_yuv(sjb_context* ctx, const uint8_t* rgb, int rgb_on_device, int width, int height, long long stride,
                  uint8_t* y, uint8_t* u, uint8_t* v, int out_on_device) try {
  if (ctx == nullptr || y == nullptr || u == nullptr || v == nullptr) return SJB_ERR_ARG;
  ctx->err.clear();
  const uint8_t* d_rgb;
  long long d_stride;
  RC(ResidentRgb(ctx, rgb, rgb_on_device, width, height, stride, &d_rgb, &d_stride));
  uint8_t *dy, *du, *dv;
  RC(SharpOnDevice(ctx, d_rgb, d_stride, width, height, &dy, &du, &dv));
  Lane* L = &ctx->lanes[0];
  const cudaMemcpyKind kind = out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  const size_t cw = (static_cast<size_t>(width) + 1) / 2, ch = (static_cast<size_t>(height) + 1) / 2;
  CU(cudaMemcpyAsync(y, dy, static_cast<size_t>(width) * height, kind, L->stream));
  CU(cudaMemcpyAsync(u, du, cw * ch, kind, L->stream));
  CU(cudaMemcpyAsync(v, dv, cw * ch, kind, L->stream));
  CU(cudaStreamSynchronize(L->stream));
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_context_set_search(sjb_context* ctx, sjb_search* search) try {
  if (ctx == nullptr) return SJB_ERR_ARG;
  if (search && (search->next_matrix == nullptr || search->update == nullptr || search->passes < 1)) return SJB_ERR_ARG;
  ctx->search = search;
  return SJB_OK;
} SJB_NOTHROW_END

// ---- stage-level entry points ----------------------------------------------------------------
static int StageF1(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                   const sjb_params* params, bool raw, Plan* plan, FrameSet* fs) {
  ctx->err.clear();
  ctx->lanes[0].last_size = 0;
  RC(MakePlan(width, height, stride, params, plan));
  CU(cudaSetDevice(ctx->device));
  Lane* L = &ctx->lanes[0];
  RC(ReserveLane(ctx, L, *plan, 1));
  RC(ReservePix(ctx, L, *plan, stride, 1));
  FillFrameSet(*plan, stride, fs);
  fs->frames = 1;
  RC(UploadPicture(ctx, L, pix, *plan, stride, 0, &fs->pix[0], &fs->stride));
  uint8_t quant[2][64], min_quant[2][64];
  QuantTabs qt;
  if (!MakeQuantTabs(*plan, quant, min_quant, &qt)) return SJB_ERR_ARG;
  // the fast kernel skips all-zero chunks: give the dump a defined background
  CU(cudaMemsetAsync(L->coef.ptr, 0, coef_padded_blocks(plan->g.nb_blocks()) * 64 * sizeof(int16_t), L->stream));
  LaunchF1(L, *fs, plan->g, raw, qt);
  CU(cudaGetLastError());
  return SJB_OK;
}

int sjb_stage_coefficients(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                           const sjb_params* params, int quantise, int16_t* coef, uint8_t* nzmask) try {
  if (ctx == nullptr || pix == nullptr || coef == nullptr) return SJB_ERR_ARG;
  Plan plan;
  FrameSet fs;
  RC(StageF1(ctx, pix, width, height, stride, params, quantise == 0, &plan, &fs));
  Lane* L = &ctx->lanes[0];
  const size_t nb = plan.g.nb_blocks();
  // the device array is sector-interleaved (block_ops.cuh); the caller gets plain [nblocks][64]
  const size_t stored_count = coef_padded_blocks(nb) * 64;
  std::unique_ptr<int16_t[]> stored(new (std::nothrow) int16_t[stored_count]);
  if (!stored) return SJB_ERR_NOMEM;
  CU(cudaMemcpyAsync(stored.get(), L->coef.ptr, stored_count * sizeof(int16_t), cudaMemcpyDeviceToHost, L->stream));
  if (quantise && nzmask) {
    CU(cudaMemcpyAsync(nzmask, L->nzmask.ptr, nb * sizeof(uint8_t), cudaMemcpyDeviceToHost, L->stream));
  }
  CU(cudaStreamSynchronize(L->stream));
  for (size_t g = 0; g < nb; ++g) {
    const int16_t* src = stored.get() + coef_block_base(g);
    for (int s = 0; s < 4; ++s) memcpy(coef + g * 64 + 16 * s, src + s * kCoefSectorStride, 16 * sizeof(int16_t));
  }
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_stage_histogram(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                        const sjb_params* params, int32_t* counts) try {
  if (ctx == nullptr || pix == nullptr || counts == nullptr) return SJB_ERR_ARG;
  Plan plan;
  FrameSet fs;
  RC(StageF1(ctx, pix, width, height, stride, params, true, &plan, &fs));
  Lane* L = &ctx->lanes[0];
  SmallLayout* D = L->d_small();
  CU(cudaMemsetAsync(D->hist, 0, sizeof(D->hist[0]), L->stream));
  LaunchHistogram(fs, L->gb, L->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(counts, D->hist, sizeof(D->hist[0]), cudaMemcpyDeviceToHost, L->stream));
  CU(cudaStreamSynchronize(L->stream));
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_stage_adapted_matrices(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                               const sjb_params* params, uint8_t quant[2][64]) try {
  if (ctx == nullptr || pix == nullptr || quant == nullptr) return SJB_ERR_ARG;
  Plan plan;
  FrameSet fs;
  RC(StageF1(ctx, pix, width, height, stride, params, true, &plan, &fs));
  Lane* L = &ctx->lanes[0];
  SmallLayout* D = L->d_small();
  uint8_t quant0[2][64], min_quant[2][64];
  QuantTabs qt;
  if (!MakeQuantTabs(plan, quant0, min_quant, &qt)) return SJB_ERR_ARG;
  CU(cudaMemsetAsync(D->hist, 0, sizeof(D->hist[0]), L->stream));
  LaunchHistogram(fs, L->gb, L->stream);
  AqParams ap;
  FillAqParams(plan, quant0, min_quant, &ap);
  LaunchAnalyseHistograms(1, L->gb, ap, &D->aq_fit[0][0][0], D->aq_fail, L->stream);
  CU(cudaGetLastError());
  int fail = 0;
  CU(cudaMemcpyAsync(quant, D->quant, 128, cudaMemcpyDeviceToHost, L->stream));
  CU(cudaMemcpyAsync(&fail, D->aq_fail, sizeof(int), cudaMemcpyDeviceToHost, L->stream));
  CU(cudaStreamSynchronize(L->stream));
  return fail ? SJB_ERR_ARG : SJB_OK;
} SJB_NOTHROW_END

int sjb_stage_symbol_stats(sjb_context* ctx, const uint8_t* pix, int width, int height, long long stride,
                           const sjb_params* params, uint32_t* freq_ac, uint32_t* freq_dc) try {
  if (ctx == nullptr || pix == nullptr || freq_ac == nullptr || freq_dc == nullptr) return SJB_ERR_ARG;
  Plan plan;
  FrameSet fs;
  RC(StageF1(ctx, pix, width, height, stride, params, false, &plan, &fs));
  Lane* L = &ctx->lanes[0];
  SmallLayout* D = L->d_small();
  CU(cudaMemsetAsync(D->freq, 0, sizeof(D->freq[0]), L->stream));
  LaunchSymbolStats(fs, L->gb, L->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(L->host->freq[0], D->freq, sizeof(D->freq[0]), cudaMemcpyDeviceToHost, L->stream));
  CU(cudaStreamSynchronize(L->stream));
  for (int c = 0; c < 2; ++c) {
    memcpy(freq_ac + 256 * c, L->host->freq[0] + 272 * c, 256 * sizeof(uint32_t));
    memcpy(freq_dc + 12 * c, L->host->freq[0] + 272 * c + 256, 12 * sizeof(uint32_t));
  }
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_last_stage_timings(sjb_context* ctx, float ms[6], int* frames) {
  if (ctx == nullptr || ms == nullptr) return SJB_ERR_ARG;
  Lane* L = &ctx->lanes[0];
  if (L->stream == nullptr) return SJB_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(L->stream);
  for (int i = 0; i < 6; ++i) {
    float v = 0.f;
    if (!L->kev_set[i] || cudaEventElapsedTime(&v, L->kev[i][0], L->kev[i][1]) != cudaSuccess) v = -1.f;
    ms[i] = v;
  }
  cudaGetLastError();
  if (frames) *frames = L->timed_frames;
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
                     float* f1_ms, size_t* jpeg_bytes, unsigned long long* launches) try {
  if (ctx == nullptr || dev_pix == nullptr || n <= 0 || iters <= 0 || total_ms == nullptr) return SJB_ERR_ARG;
  ctx->err.clear();
  ctx->lanes[0].last_size = 0;
  Plan plan;
  RC(MakePlan(width, height, stride, params, &plan));
  CU(cudaSetDevice(ctx->device));
  const int B = std::max(1, std::min(plan.group, n));
  const int groups = (n + B - 1) / B;
  // groups rotate over the lanes ACROSS iterations, so that a batch of few large groups still keeps
  // all lanes (streams) busy: the tail of one group's kernels overlaps the head of the next one's
  const int nl = std::min<int>(kMaxLanes, groups * iters);
  for (int l = 0; l < nl; ++l) {
    RC(InitLane(ctx, &ctx->lanes[l]));
    RC(ReserveLane(ctx, &ctx->lanes[l], plan, B));
    ctx->lanes[l].launches = 0;
  }
  CU(cudaDeviceSynchronize());
  struct EventPair {      // destroyed on every return path
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() {
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
    }
  } span;
  CU(cudaEventCreate(&span.a));
  CU(cudaEventCreate(&span.b));
  const cudaEvent_t t0 = span.a, t1 = span.b;
  Lane* L0 = &ctx->lanes[0];
  CU(cudaEventRecord(t0, L0->stream));
  for (int l = 1; l < nl; ++l) CU(cudaStreamWaitEvent(ctx->lanes[l].stream, t0, 0));
  ctx->bench_slots.assign(n, sjb_context::BenchSlot{-1, -1, -1});
  for (int& t : ctx->bench_last_turn) t = -1;
  std::vector<GroupJob> jobs(nl);
  for (int it = 0, turn = 0; it < iters; ++it) {
    for (int k = 0; k < groups; ++k, ++turn) {
      Lane* L = &ctx->lanes[turn % nl];
      GroupJob* J = &jobs[turn % nl];
      while (J->stage < 3) RC(AdvanceGroup(ctx, J));       // the group that used this lane nl turns ago
      FrameSet fs;
      FillFrameSet(plan, stride, &fs);
      fs.frames = std::min(B, n - k * B);
      for (int f = 0; f < fs.frames; ++f) {
        fs.pix[f] = dev_pix[k * B + f];
        ctx->bench_slots[k * B + f] = sjb_context::BenchSlot{turn % nl, f, turn};
      }
      ctx->bench_last_turn[turn % nl] = turn;
      RC(StartGroup(ctx, J, L, fs, plan, /*timed=*/turn % nl == 0));
      // methods >= 1: one host phase of every older group in flight, oldest first (see sjb_encode_batch)
      for (int j = std::max(0, turn - nl + 1); j < turn; ++j) RC(AdvanceGroup(ctx, &jobs[j % nl]));
    }
  }
  for (int l = 0; l < nl; ++l) while (jobs[l].stage < 3) RC(AdvanceGroup(ctx, &jobs[l]));
  for (int l = 1; l < nl; ++l) {
    CU(cudaEventRecord(ctx->lanes[l].ev[3], ctx->lanes[l].stream));
    CU(cudaStreamWaitEvent(L0->stream, ctx->lanes[l].ev[3], 0));
  }
  CU(cudaEventRecord(t1, L0->stream));
  CU(cudaEventSynchronize(t1));
  CU(cudaEventElapsedTime(total_ms, t0, t1));
  if (f1_ms) {
    float ms = 0;
    *f1_ms = (cudaEventElapsedTime(&ms, L0->ev[0], L0->ev[1]) == cudaSuccess) ? ms : 0.f;
  }
  FinishTimings(ctx, L0);   // F1 / entropy split of the last group timed on lane 0 (sjb_last_timings)
  if (jpeg_bytes) *jpeg_bytes = static_cast<size_t>(L0->host->info[0].out_size);
  if (launches) {
    *launches = 0;
    for (int l = 0; l < nl; ++l) *launches += ctx->lanes[l].launches;
  }
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_bench_output(sjb_context* ctx, int index, uint8_t* out, size_t out_capacity, size_t* out_size) try {
  if (ctx == nullptr || out_size == nullptr || index < 0 || index >= static_cast<int>(ctx->bench_slots.size()))
    return SJB_ERR_ARG;
  const sjb_context::BenchSlot b = ctx->bench_slots[index];
  if (b.lane < 0 || ctx->bench_last_turn[b.lane] != b.turn) return SJB_ERR_ARG;   // overwritten by a later group
  CU(cudaSetDevice(ctx->device));
  Lane* L = &ctx->lanes[b.lane];
  CU(cudaStreamSynchronize(L->stream));
  const size_t size = static_cast<size_t>(L->host->info[b.slot].out_size);
  *out_size = size;
  if (out == nullptr || size > out_capacity) return SJB_ERR_CAPACITY;
  CU(cudaMemcpyAsync(out, L->gb.out + b.slot * L->gb.out_pitch, size, cudaMemcpyDeviceToHost, L->stream));
  CU(cudaStreamSynchronize(L->stream));
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_bench_f1(sjb_context* ctx, int n, const uint8_t* const* dev_pix, int width, int height,
                 long long stride, const sjb_params* params, int iters, float* ms_per_launch,
                 int* frames_per_launch) try {
  if (ctx == nullptr || dev_pix == nullptr || n <= 0 || iters <= 0 || ms_per_launch == nullptr) return SJB_ERR_ARG;
  ctx->err.clear();
  ctx->lanes[0].last_size = 0;
  Plan plan;
  RC(MakePlan(width, height, stride, params, &plan));
  CU(cudaSetDevice(ctx->device));
  Lane* L = &ctx->lanes[0];
  const int B = std::max(1, std::min(plan.group, n));
  const int groups = n / B;             // whole groups only, so that every launch is identical
  if (groups < 1) return SJB_ERR_ARG;
  RC(ReserveLane(ctx, L, plan, B));
  uint8_t quant[2][64], min_quant[2][64];
  QuantTabs qt;
  if (!MakeQuantTabs(plan, quant, min_quant, &qt)) return SJB_ERR_ARG;
  CU(cudaStreamSynchronize(L->stream));
  CU(cudaEventRecord(L->ev[0], L->stream));
  for (int it = 0; it < iters; ++it) {
    for (int k = 0; k < groups; ++k) {
      FrameSet fs;
      FillFrameSet(plan, stride, &fs);
      fs.frames = B;
      for (int f = 0; f < B; ++f) fs.pix[f] = dev_pix[k * B + f];
      LaunchF1(L, fs, plan.g, false, qt);
    }
  }
  CU(cudaEventRecord(L->ev[1], L->stream));
  CU(cudaEventSynchronize(L->ev[1]));
  CU(cudaGetLastError());
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, L->ev[0], L->ev[1]));
  *ms_per_launch = ms / (static_cast<float>(groups) * iters);
  if (frames_per_launch) *frames_per_launch = B;
  return SJB_OK;
} SJB_NOTHROW_END

// ---------------------------------------------------------------------------------------------
// Row stripes of pictures split across GPUs (BASELINE.json config 5).  A session
// holds n stripes (one per picture, same geometry) between the three phases; the caller (one
// process per GPU, sjpeg_b200/distributed.py) exchanges the DC predictors and the bit offsets
// between the phases.  Default Huffman tables only (method 0).
// ---------------------------------------------------------------------------------------------
}  // extern "C"

struct sjb_stripes {
  sjb_context* ctx = nullptr;
  Plan plan;
  int n = 0;
  long long stride = 0;
  std::vector<Lane*> sets;          // one buffer set per group of <= kMaxGroup stripes
  std::vector<FrameSet> fsets;
  DeviceBuffer dc;                  // [n][3] last DCs out, then [n][3] predictors in
  bool transformed = false, coded = false;
};

extern "C" {

int sjb_stripes_create(sjb_context* ctx, int n, int width, int stripe_height, const sjb_params* params,
                       sjb_stripes** out) try {
  if (ctx == nullptr || out == nullptr || n <= 0) return SJB_ERR_ARG;
  *out = nullptr;
  ctx->err.clear();
  sjb_stripes* s = new (std::nothrow) sjb_stripes();
  if (s == nullptr) return SJB_ERR_NOMEM;
  s->ctx = ctx;
  s->n = n;
  const int pstep = (params && params->pix_fmt != SJB_PIX_RGB) ? 4 : 3;
  int rc = MakePlan(width, stripe_height, static_cast<long long>(pstep) * width, params, &s->plan);
  if (rc == SJB_OK && s->plan.p.method != 0) rc = SJB_ERR_ARG;     // stripes: default tables only
  if (rc != SJB_OK) {
    delete s;
    return rc;
  }
  *out = s;
  return SJB_OK;
} SJB_NOTHROW_END

void sjb_stripes_destroy(sjb_stripes* s) {
  if (s == nullptr) return;
  cudaSetDevice(s->ctx->device);
  for (Lane* L : s->sets) {
    DestroyLane(L);
    delete L;
  }
  s->dc.Release();
  delete s;
}

int sjb_stripes_transform(sjb_stripes* s, const uint8_t* const* pix, int pix_on_device, long long stride,
                          int* last_dc) try {
  if (s == nullptr || pix == nullptr || last_dc == nullptr) return SJB_ERR_ARG;
  sjb_context* ctx = s->ctx;
  ctx->err.clear();
  Plan plan;
  RC(MakePlan(s->plan.g.width, s->plan.g.height, stride, &s->plan.p, &plan));   // validates the stride
  s->plan = plan;
  s->stride = stride;
  CU(cudaSetDevice(ctx->device));
  const int groups = (s->n + kMaxGroup - 1) / kMaxGroup;
  while (static_cast<int>(s->sets.size()) < groups) {
    Lane* L = new (std::nothrow) Lane();
    if (L == nullptr) return SJB_ERR_NOMEM;
    s->sets.push_back(L);
    RC(InitLane(ctx, L));
  }
  CU(s->dc.Reserve(static_cast<size_t>(s->n) * 6 * sizeof(int)));
  s->fsets.assign(groups, FrameSet());
  uint8_t quant[2][64], min_quant[2][64];
  QuantTabs qt;
  if (!MakeQuantTabs(plan, quant, min_quant, &qt)) return SJB_ERR_ARG;
  std::vector<int> host_dc(static_cast<size_t>(s->n) * 3);
  for (int k = 0; k < groups; ++k) {
    Lane* L = s->sets[k];
    const int frames = std::min<int>(kMaxGroup, s->n - k * kMaxGroup);
    RC(ReserveLane(ctx, L, plan, frames));
    FrameSet& fs = s->fsets[k];
    FillFrameSet(plan, stride, &fs);
    fs.frames = frames;
    if (!pix_on_device) RC(ReservePix(ctx, L, plan, stride, frames));
    for (int f = 0; f < frames; ++f) {
      const uint8_t* p = pix[k * kMaxGroup + f];
      if (p == nullptr) return SJB_ERR_ARG;
      fs.pix[f] = p;
      if (!pix_on_device) {
        long long ds = stride;
        RC(UploadPicture(ctx, L, p, plan, stride, f, &fs.pix[f], &ds));
        fs.stride = ds;
      }
    }
    LaunchF1(L, fs, plan.g, /*raw=*/false, qt);
    LaunchLastDc(fs, L->gb, s->dc.as<int>() + k * kMaxGroup * 3, L->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&host_dc[static_cast<size_t>(k) * kMaxGroup * 3], s->dc.as<int>() + k * kMaxGroup * 3,
                       frames * 3 * sizeof(int), cudaMemcpyDeviceToHost, L->stream));
  }
  for (int k = 0; k < groups; ++k) CU(cudaStreamSynchronize(s->sets[k]->stream));
  memcpy(last_dc, host_dc.data(), host_dc.size() * sizeof(int));
  s->transformed = true;
  s->coded = false;
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_stripes_code(sjb_stripes* s, const int* dc_pred, unsigned long long* bits) try {
  if (s == nullptr || dc_pred == nullptr || bits == nullptr || !s->transformed) return SJB_ERR_ARG;
  sjb_context* ctx = s->ctx;
  ctx->err.clear();
  CU(cudaSetDevice(ctx->device));
  const int groups = static_cast<int>(s->fsets.size());
  int* d_pred = s->dc.as<int>() + static_cast<size_t>(s->n) * 3;
  // default Huffman tables for every stripe
  HuffSpec spec[4];
  CodeTabs tabs;
  memset(&tabs, 0, sizeof(tabs));
  for (int i = 0; i < 4; ++i) DefaultHuffSpec(i >= 2, i & 1, &spec[i]);
  for (int c = 0; c < 2; ++c) {
    CodesFromSpec(spec[c], tabs.dc[c]);
    CodesFromSpec(spec[2 + c], tabs.ac[c]);
  }
  for (int k = 0; k < groups; ++k) {
    Lane* L = s->sets[k];
    const FrameSet& fs = s->fsets[k];
    CU(cudaStreamSynchronize(L->stream));
    for (int f = 0; f < fs.frames; ++f) L->host->tabs[f] = tabs;
    CU(cudaMemcpyAsync(L->d_small()->tabs, L->host->tabs, fs.frames * sizeof(CodeTabs), cudaMemcpyHostToDevice,
                       L->stream));
    L->tabs_valid = 0;
    CU(cudaMemcpyAsync(d_pred + k * kMaxGroup * 3, dc_pred + k * kMaxGroup * 3, fs.frames * 3 * sizeof(int),
                       cudaMemcpyHostToDevice, L->stream));
    if (L->words_dirty) {
      CU(cudaMemsetAsync(L->words.ptr, 0, L->words.bytes, L->stream));
      L->words_dirty = false;
    }
    CU(cudaMemsetAsync(L->state.ptr, 0, L->state.bytes, L->stream));
    GroupBuffers gb = L->gb;
    gb.dc_init = d_pred + k * kMaxGroup * 3;
    L->words_dirty = true;
    LaunchEntropyPack(fs, gb, L->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(L->host->info, L->d_small()->info, fs.frames * sizeof(StreamInfo), cudaMemcpyDeviceToHost,
                       L->stream));
  }
  for (int k = 0; k < groups; ++k) {
    Lane* L = s->sets[k];
    CU(cudaStreamSynchronize(L->stream));
    for (int f = 0; f < s->fsets[k].frames; ++f) bits[k * kMaxGroup + f] = L->host->info[f].total_bits;
  }
  s->coded = true;
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_stripes_finish(sjb_stripes* s, const unsigned long long* bit_offsets, int is_first, int is_last,
                       uint8_t* const* out, size_t out_capacity, size_t* sizes, unsigned char* head_byte,
                       unsigned char* tail_byte, unsigned char* tail_bits) try {
  if (s == nullptr || bit_offsets == nullptr || out == nullptr || sizes == nullptr || head_byte == nullptr ||
      tail_byte == nullptr || tail_bits == nullptr || !s->coded)
    return SJB_ERR_ARG;
  sjb_context* ctx = s->ctx;
  ctx->err.clear();
  CU(cudaSetDevice(ctx->device));
  const int groups = static_cast<int>(s->fsets.size());
  // A stripe that is neither the first nor the last and whose bits all fall inside the byte it
  // shares with its predecessor (offset % 8 + bits < 8: only a flat 4:0:0 stripe of one block can
  // be that short) owns no byte boundary at all; the (head, tail) hand-over cannot express three
  // stripes meeting in one byte, so such a partition is refused instead of mis-assembled.
  for (int k = 0; k < groups; ++k) {
    for (int f = 0; f < s->fsets[k].frames; ++f) {
      const unsigned long long off = bit_offsets[k * kMaxGroup + f] & 7;
      if (!is_first && off != 0 && off + s->sets[k]->host->info[f].total_bits < 8) {
        ctx->err = "row stripe shorter than the byte it shares with its neighbours: use fewer stripes";
        return SJB_ERR_ARG;
      }
    }
  }
  for (int k = 0; k < groups; ++k) {
    Lane* L = s->sets[k];
    const FrameSet& fs = s->fsets[k];
    StuffArgs sa;
    memset(&sa, 0, sizeof(sa));
    for (int f = 0; f < fs.frames; ++f) {
      sa.header_len[f] = 0;
      sa.shift[f] = static_cast<unsigned>(bit_offsets[k * kMaxGroup + f] & 7);
      sa.flags[f] = (is_first ? kStuffFirst : 0) | (is_last ? kStuffLast : 0) | kStuffKeepWords;
    }
    LaunchStuff(fs, L->gb, sa, L->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(L->host->info, L->d_small()->info, fs.frames * sizeof(StreamInfo), cudaMemcpyDeviceToHost,
                       L->stream));
    L->header_valid = 0;      // the out slots now start with scan bytes
    L->words_dirty = true;    // shifted reads cannot self-clean: rezero before the next use
  }
  int rc = SJB_OK;
  for (int k = 0; k < groups; ++k) {
    Lane* L = s->sets[k];
    CU(cudaStreamSynchronize(L->stream));
    for (int f = 0; f < s->fsets[k].frames; ++f) {
      const int i = k * kMaxGroup + f;
      const StreamInfo& info = L->host->info[f];
      sizes[i] = static_cast<size_t>(info.out_size);
      head_byte[i] = info.head_byte;
      tail_byte[i] = info.tail_byte;
      tail_bits[i] = info.tail_bits;
      if (out[i] == nullptr || sizes[i] > out_capacity) {
        rc = SJB_ERR_CAPACITY;
        continue;
      }
      CU(cudaMemcpyAsync(out[i], L->gb.out + f * L->gb.out_pitch, sizes[i], cudaMemcpyDeviceToHost, L->stream));
    }
  }
  for (int k = 0; k < groups; ++k) CU(cudaStreamSynchronize(s->sets[k]->stream));
  s->coded = false;
  s->transformed = false;
  return rc;
} SJB_NOTHROW_END

int sjb_picture_header(const sjb_params* params, int width, int height, uint8_t* out, size_t out_capacity,
                       size_t* out_size) try {
  if (params == nullptr || out_size == nullptr) return SJB_ERR_ARG;
  Plan plan;
  const int pstep = (params->pix_fmt != SJB_PIX_RGB) ? 4 : 3;
  RC(MakePlan(width, height, static_cast<long long>(pstep) * width, params, &plan));
  if (plan.p.method != 0) return SJB_ERR_ARG;   // optimised tables / adapted matrices are data dependent
  uint8_t quant[2][64], min_quant[2][64];
  QuantTabs qt;
  if (!MakeQuantTabs(plan, quant, min_quant, &qt)) return SJB_ERR_ARG;
  HuffSpec spec[4];
  for (int i = 0; i < 4; ++i) DefaultHuffSpec(i >= 2, i & 1, &spec[i]);
  std::vector<uint8_t> h;
  AppendHeaders(plan.g, quant, spec, &h);
  *out_size = h.size();
  if (out == nullptr || h.size() > out_capacity) return SJB_ERR_CAPACITY;
  memcpy(out, h.data(), h.size());
  return SJB_OK;
} SJB_NOTHROW_END

}  // extern "C"

#include "engine_stripes.inl"
