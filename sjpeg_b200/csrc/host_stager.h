// host_stager.h -- host-to-device upload of PAGEABLE memory through a pinned ring, with the
// host-side copy spread over a few helper threads.
//
// Why: the drop-in entry points (SjpegEncode, sjpeg::Encode; /root/reference/src/sjpeg.h:104,280)
// receive whatever buffer the caller has, normally malloc()ed.  cudaMemcpyAsync from pageable
// memory is staged by the driver on the calling thread at about 11-19 GB/s on the B200 boxes (24.9 MB
// of 4K RGB: 1.2 ms cache-warm, 2.3 ms out of DRAM), while the device pipeline for that picture takes 0.06 ms, so the copy IS
// the call.  Here the picture is cut into chunks; the caller and kHelpers helper threads copy
// each chunk into pinned memory together -- in kPiece pieces claimed from a shared cursor, so the
// caller never waits for a helper that is still waking up -- and the DMA of chunk c overlaps the
// host copy of chunk c+1.  Pinned / registered / managed sources never come here (engine.cu checks the
// pointer's type first).
//
// One stager per context; a context is used by one thread at a time (include/sjpeg_b200.h), so
// Upload() is never entered concurrently.  The helpers spin while an upload is running and for kGraceUs
// after it (the next upload of a stream of frames finds them awake), then sleep on a condition variable.
// Across contexts, uploads only get helpers while callers x (kHelpers + 1) fits the hardware threads.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace sjb {

class HostStager {
 public:
  // tools/micro/stager_bench.cu, uploads 0.25 ms apart, SOURCES ROTATING over more memory than the host's
  // last-level cache (a stream of frames arrives from DRAM; the same buffer re-sent stays cache-warm and
  // flatters every figure by 20-25 %), 4K / 8K picture:
  //   driver's pageable path 2.3 / 8.8 ms;  pinned source 0.458 / 1.80 ms
  //   3 helpers 0.71 / 2.62 ms (35 / 38 GB/s), 5 helpers 0.66-0.70 / 2.07-2.27 ms (up to 48 GB/s), 7 helpers 0.67 / 2.10 ms
  //   (4 MB chunks, 256 KB pieces; 2 MB chunks: 0.76-0.83 / 2.6-2.9 ms)
  // With the helpers asleep between uploads (no grace period) 5 and 7 helpers had been SLOWER than 3 at 4K:
  // more threads to wake.  A cache-warm 4K source: 0.575 ms with 3 helpers.
#ifndef SJB_STAGER_HELPERS
#define SJB_STAGER_HELPERS 5
#endif
#ifndef SJB_STAGER_CHUNK_KB
#define SJB_STAGER_CHUNK_KB 4096
#endif
#ifndef SJB_STAGER_PIECE_KB
#define SJB_STAGER_PIECE_KB 256
#endif
  enum { kHelpers = SJB_STAGER_HELPERS, kSlots = 4 };
  static constexpr size_t kChunk = static_cast<size_t>(SJB_STAGER_CHUNK_KB) << 10;   // bytes per pinned slot
  static constexpr size_t kPiece = static_cast<size_t>(SJB_STAGER_PIECE_KB) << 10;   // unit of work claimed by a thread
  static constexpr size_t kMinBytes = 4u << 20;    // below this the driver's own path is as good
#ifndef SJB_STAGER_GRACE_US
#define SJB_STAGER_GRACE_US 1000
#endif
  enum { kGraceUs = SJB_STAGER_GRACE_US };          // how long the helpers stay awake after an upload

  HostStager() = default;
  ~HostStager();
  HostStager(const HostStager&) = delete;
  HostStager& operator=(const HostStager&) = delete;

  // Asynchronous on `stream` like cudaMemcpyAsync, except that the SOURCE may be reused as soon
  // as the call returns (it has been copied out).  Returns cudaErrorNotSupported when the stager
  // could not be set up (no threads / no pinned memory): the caller falls back to cudaMemcpyAsync.
  cudaError_t Upload(void* dst_device, const void* src_host, size_t bytes, cudaStream_t stream);

 private:
  bool Start();
  void HelperLoop(int id);
  void CopyChunk(uint8_t* dst, const uint8_t* src, size_t bytes);

  bool started_ = false, failed_ = false;
  uint8_t* pinned_ = nullptr;                       // kSlots * kChunk
  cudaEvent_t slot_free_[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  std::thread helpers_[kHelpers];

  // sleeping / waking
  std::mutex mutex_;
  std::condition_variable wake_;
  bool active_ = false, quit_ = false;
  // One chunk job at a time; cursor_ packs (ticket, pieces, next piece), see host_stager.cc.
  std::atomic<unsigned long long> cursor_{0};
  std::atomic<unsigned> pieces_done_{0};
  std::atomic<bool> spinning_{false};
  unsigned ticket_ = 0;
  int next_slot_ = 0;                               // the ring goes on where the previous upload stopped
  void Work();
  uint8_t* job_dst_ = nullptr;
  const uint8_t* job_src_ = nullptr;
  size_t job_bytes_ = 0;
};

}  // namespace sjb
