// Small MPSC work queue: short spin for latency, then sleep on a condvar.
// (The reference uses unbounded flume channels and yield-spinning IO loops;
// see SURVEY.md §2.2.  Ours is bounded by the request pool feeding it.)
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>

#include "core/common.h"

namespace bnet {

template <typename T>
class WorkQ {
 public:
  void push(const T& v) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      q_.push_back(v);
      size_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_one();
  }
  // false when the queue was stopped and is empty
  bool pop(T* out, int spin_us) {
    if (spin_us > 0 && size_.load(std::memory_order_acquire) == 0 && !stopped_.load(std::memory_order_relaxed)) {
      uint64_t until = now_ns() + (uint64_t)spin_us * 1000ull;
      while (size_.load(std::memory_order_acquire) == 0 && now_ns() < until) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
    }
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !q_.empty() || stopped_.load(std::memory_order_relaxed); });
    if (q_.empty()) return false;
    *out = q_.front();
    q_.pop_front();
    size_.fetch_sub(1, std::memory_order_release);
    return true;
  }
  bool try_pop(T* out) {
    std::lock_guard<std::mutex> lk(mu_);
    if (q_.empty()) return false;
    *out = q_.front();
    q_.pop_front();
    size_.fetch_sub(1, std::memory_order_release);
    return true;
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stopped_.store(true);
    }
    cv_.notify_all();
  }
  size_t size() const { return size_.load(std::memory_order_acquire); }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<T> q_;
  std::atomic<size_t> size_{0};
  std::atomic<bool> stopped_{false};
};

}  // namespace bnet
