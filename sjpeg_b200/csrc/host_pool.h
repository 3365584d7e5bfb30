// host_pool.h -- a few worker threads for the per-picture host work of a GROUP of pictures:
// AnalyseHistograms (histogram.cc:126-315, 0.05-0.12 ms per 4K picture) and OptimalHuffSpec
// (entropy.cc:254-430) are independent per picture, and with 8 pictures per launch one thread
// doing them in turn is slower than the device pipeline it feeds.  ParallelFor(n, fn) runs fn(0..n-1)
// on the caller plus the workers and returns when all are done; n <= 1 runs inline and never
// starts a thread.  One pool per context; a context is used by one thread at a time, so
// ParallelFor is never entered concurrently.  Workers sleep on a condition variable between calls.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace sjb {

class HostPool {
 public:
  HostPool() = default;
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      quit_ = true;
    }
    wake_.notify_all();
    for (auto& t : workers_) t.join();
  }
  HostPool(const HostPool&) = delete;
  HostPool& operator=(const HostPool&) = delete;

  template <class Fn>
  void ParallelFor(int n, Fn fn) {
    if (n <= 1 || !Start()) {
      for (int i = 0; i < n; ++i) fn(i);
      return;
    }
    std::function<void(int)> f(fn);
    {
      std::lock_guard<std::mutex> lock(mutex_);
      job_ = &f;
      job_n_ = n;
      next_.store(0, std::memory_order_relaxed);
      done_.store(0, std::memory_order_relaxed);
      ++generation_;
    }
    wake_.notify_all();
    Work(&f, n);
    // every index finished AND every worker that picked this job up has let go of it (a worker
    // holds the pointer to `f` from pick-up until it leaves Work): only then may `f` die and the
    // counters be reused
    std::unique_lock<std::mutex> lock(mutex_);
    finished_.wait(lock, [&] { return done_.load(std::memory_order_acquire) >= n && inside_ == 0; });
    job_ = nullptr;
  }

 private:
  bool Start() {
    if (started_) return !workers_.empty();
    started_ = true;
    unsigned hw = std::thread::hardware_concurrency();
    int want = static_cast<int>(hw > 1 ? hw - 1 : 0);
    if (want > 7) want = 7;
    try {
      for (int i = 0; i < want; ++i) workers_.emplace_back(&HostPool::Loop, this);
    } catch (...) {
    }
    return !workers_.empty();
  }
  void Work(std::function<void(int)>* f, int n) {
    for (;;) {
      const int i = next_.fetch_add(1, std::memory_order_acq_rel);
      if (i >= n) return;
      (*f)(i);
      if (done_.fetch_add(1, std::memory_order_acq_rel) + 1 >= n) {
        std::lock_guard<std::mutex> lock(mutex_);
        finished_.notify_all();
      }
    }
  }
  void Loop() {
    unsigned long long seen = 0;
    for (;;) {
      std::function<void(int)>* f;
      int n;
      {
        std::unique_lock<std::mutex> lock(mutex_);
        wake_.wait(lock, [&] { return quit_ || (generation_ != seen && job_ != nullptr); });
        if (quit_) return;
        seen = generation_;
        f = job_;
        n = job_n_;
        ++inside_;
      }
      Work(f, n);
      {
        std::lock_guard<std::mutex> lock(mutex_);
        --inside_;
        finished_.notify_all();
      }
    }
  }

  bool started_ = false;
  std::vector<std::thread> workers_;
  std::mutex mutex_;
  std::condition_variable wake_, finished_;
  bool quit_ = false;
  unsigned long long generation_ = 0;
  std::function<void(int)>* job_ = nullptr;
  int job_n_ = 0;
  int inside_ = 0;
  std::atomic<int> next_{0}, done_{0};
};

}  // namespace sjb
