// host_stager.cc -- see host_stager.h
#include "host_stager.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <new>

namespace sjb {

namespace {
inline void CpuRelax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}
}  // namespace

bool HostStager::Start() {
  if (started_) return true;
  if (failed_) return false;
  failed_ = true;   // until everything below has worked
  if (cudaMallocHost(reinterpret_cast<void**>(&pinned_), kSlots * kChunk) != cudaSuccess) {
    cudaGetLastError();
    pinned_ = nullptr;
    return false;
  }
  for (auto& e : slot_free_) {
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
  }
  int created = 0;
  try {
    for (; created < kHelpers; ++created) helpers_[created] = std::thread(&HostStager::HelperLoop, this, created);
  } catch (...) {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      quit_ = true;
    }
    wake_.notify_all();
    for (int i = 0; i < created; ++i) helpers_[i].join();
    quit_ = false;
    return false;
  }
  failed_ = false;
  started_ = true;
  return true;
}

HostStager::~HostStager() {
  if (started_) {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      quit_ = true;
    }
    wake_.notify_all();
    for (auto& t : helpers_) t.join();
  }
  for (auto& e : slot_free_) {
    if (e) {
      cudaEventSynchronize(e);
      cudaEventDestroy(e);
    }
  }
  if (pinned_) cudaFreeHost(pinned_);
}

// cursor_ layout: [63:40] ticket, [39:20] pieces of the job, [19:0] next piece to claim.  Every
// decision is taken on ONE atomic word: a successful compare-exchange on a word with next < pieces
// proves that the job of that ticket still had an unclaimed piece at that instant, hence that the
// owner is still waiting for it and has not begun to overwrite the job fields for the next one.
namespace {
inline unsigned long long MakeCursor(unsigned ticket, unsigned pieces) {
  return (static_cast<unsigned long long>(ticket & 0xffffffu) << 40) | (static_cast<unsigned long long>(pieces) << 20);
}
inline unsigned CursorPieces(unsigned long long c) { return static_cast<unsigned>((c >> 20) & 0xfffffu); }
inline unsigned CursorNext(unsigned long long c) { return static_cast<unsigned>(c & 0xfffffu); }
}  // namespace

// Claims and copies pieces of the current job until none is left unclaimed.
void HostStager::Work() {
  for (;;) {
    unsigned long long c = cursor_.load(std::memory_order_acquire);
    const unsigned piece = CursorNext(c);
    if (piece >= CursorPieces(c)) return;
    if (!cursor_.compare_exchange_weak(c, c + 1, std::memory_order_acq_rel)) continue;
    const size_t b = static_cast<size_t>(piece) * kPiece;
    const size_t n = std::min(kPiece, job_bytes_ - b);
    memcpy(job_dst_ + b, job_src_ + b, n);
    pieces_done_.fetch_add(1, std::memory_order_release);
  }
}

void HostStager::HelperLoop(int) {
  for (;;) {
    {
      std::unique_lock<std::mutex> lock(mutex_);
      wake_.wait(lock, [&] { return active_ || quit_; });
      if (quit_) return;
    }
    for (;;) {
      while (spinning_.load(std::memory_order_acquire)) {
        Work();
        CpuRelax();
      }
      // The upload is over.  Stay awake for a short grace period: a caller that encodes picture after
      // picture is back within a few hundred microseconds, and waking three sleeping threads through
      // the condition variable cost a quarter of a 4K upload (0.93 ms against 0.70 ms; 8K pictures,
      // where the wake-up is amortised, ran at 38 GB/s against 27).  The wait yields now and then, so
      // an oversubscribed machine (many contexts, few cores) loses little to it.
      bool again = false;
      const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(kGraceUs);
      for (unsigned spin = 0;; ++spin) {
        if (spinning_.load(std::memory_order_acquire)) {
          again = true;
          break;
        }
        if ((spin & 63u) == 63u) {
          if (std::chrono::steady_clock::now() >= until) break;
          std::this_thread::yield();
        }
        CpuRelax();
      }
      if (!again) break;
    }
    // wait until active_ has been cleared (or a new upload has begun) before sleeping on it again
    while (true) {
      std::lock_guard<std::mutex> lock(mutex_);
      if (!active_ || quit_ || spinning_.load(std::memory_order_acquire)) break;
    }
  }
}

void HostStager::CopyChunk(uint8_t* dst, const uint8_t* src, size_t bytes) {
  // nobody reads the job fields now: every piece of the previous job was claimed AND finished
  job_dst_ = dst;
  job_src_ = src;
  job_bytes_ = bytes;
  const unsigned pieces = static_cast<unsigned>((bytes + kPiece - 1) / kPiece);
  pieces_done_.store(0, std::memory_order_relaxed);
  cursor_.store(MakeCursor(++ticket_, pieces), std::memory_order_release);
  Work();
  while (pieces_done_.load(std::memory_order_acquire) < pieces) CpuRelax();
}

namespace {
// uploads in flight in this process, over all stagers (one per context, i.e. per calling thread of the
// drop-in entry points)
std::atomic<int> g_uploads_in_flight{0};
struct InFlight {
  int n;
  InFlight() : n(g_uploads_in_flight.fetch_add(1, std::memory_order_acq_rel) + 1) {}
  ~InFlight() { g_uploads_in_flight.fetch_sub(1, std::memory_order_acq_rel); }
};
}  // namespace

cudaError_t HostStager::Upload(void* dst_device, const void* src_host, size_t bytes, cudaStream_t stream) {
  if (!Start()) return cudaErrorNotSupported;
  // The helpers are there to make ONE upload as fast as the link; when many threads upload at the same
  // time their own copies already saturate it, and three helpers each only fight over the cores (16
  // threads calling SjpegEncode on 16 cores: 15.2 Gpix/s with the callers copying alone, 10.1 with 64
  // copying threads; up to four callers the helpers pay: 13-15 against 10-12 Gpix/s).  So an upload gets
  // helpers as long as callers x (helpers + 1) fits the machine's hardware threads.
  const InFlight in_flight;
  static const int max_helped = [] {
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max(1, static_cast<int>(hw ? hw : 4u) / (kHelpers + 1));
  }();
  const bool helped = in_flight.n <= max_helped;
  if (helped) {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      active_ = true;
      spinning_.store(true, std::memory_order_release);
    }
    wake_.notify_all();
  }
  cudaError_t err = cudaSuccess;
  const uint8_t* src = static_cast<const uint8_t*>(src_host);
  uint8_t* dst = static_cast<uint8_t*>(dst_device);
  // SJB_STAGER_TRACE=1: where an upload's time goes (waiting for a ring slot / copying / CUDA calls), on stderr
  static const bool trace = getenv("SJB_STAGER_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_wait = 0, t_copy = 0, t_api = 0, t0 = trace ? now() : 0;
  const double t_begin = t0;
  for (size_t off = 0; off < bytes && err == cudaSuccess; off += kChunk, next_slot_ = (next_slot_ + 1) % kSlots) {
    const int slot = next_slot_;
    const size_t n = std::min(kChunk, bytes - off);
    uint8_t* stage = pinned_ + static_cast<size_t>(slot) * kChunk;
    err = cudaEventSynchronize(slot_free_[slot]);          // the DMA that last read this slot is done
    if (err != cudaSuccess) break;
    if (trace) { const double t = now(); t_wait += t - t0; t0 = t; }
    CopyChunk(stage, src + off, n);
    if (trace) { const double t = now(); t_copy += t - t0; t0 = t; }
    err = cudaMemcpyAsync(dst + off, stage, n, cudaMemcpyHostToDevice, stream);
    if (err == cudaSuccess) err = cudaEventRecord(slot_free_[slot], stream);
    if (trace) { const double t = now(); t_api += t - t0; t0 = t; }
  }
  if (trace) {
    fprintf(stderr, "[stager] %.1f MB in %.0f us: slot wait %.0f, copy %.0f, cuda calls %.0f; helped %d (in flight %d)\n", bytes / 1e6,
            now() - t_begin, t_wait, t_copy, t_api, helped ? 1 : 0, in_flight.n);
  }
  if (helped) {
    spinning_.store(false, std::memory_order_release);
    std::lock_guard<std::mutex> lock(mutex_);
    active_ = false;
  }
  return err;
}

}  // namespace sjb
