// A host-only stand-in for the handful of CUDA runtime calls csrc/staging.cpp makes, so that the pinned ring's
// PROTOCOL (slot queues, event hand-off, host-slot reuse, copies issued from two threads) can be stress-tested — and run
// under ThreadSanitizer — on a machine without a GPU ("TSAN-able C++ host code").
//
// Streams are FIFO worker threads, so copies really are asynchronous with respect to the caller: a producer that
// overwrites a pinned slot before the copy engine has read it (the round-1 bug) corrupts the batch here exactly as it
// would on the device.  Events follow CUDA semantics: cudaStreamWaitEvent / cudaEventSynchronize wait for the most
// recent cudaEventRecord issued BEFORE the call (no-op if there is none).  cudaMemcpyAsync sleeps a random time up to
// mock_jitter_us() before copying, to shake out orderings.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDisableTiming = 2 };

inline std::atomic<long>& mock_jitter_us() {
  static std::atomic<long> v{0};
  return v;
}

struct MockEvent {
  std::mutex mu;
  std::condition_variable cv;
  long recorded = 0, done = 0;
};
typedef MockEvent* cudaEvent_t;

struct MockStream {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::function<void()>> q;
  bool stopping = false;
  uint64_t rng = 0x9E3779B97F4A7C15ull;
  std::thread worker;
  MockStream() {
    worker = std::thread([this] {
      for (;;) {
        std::function<void()> op;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return stopping || !q.empty(); });
          if (q.empty()) return;
          op = std::move(q.front());
          q.pop_front();
        }
        op();
      }
    });
  }
  ~MockStream() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stopping = true;
    }
    cv.notify_all();
    if (worker.joinable()) worker.join();
  }
  void push(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> lk(mu);
      q.push_back(std::move(f));
    }
    cv.notify_all();
  }
};
typedef MockStream* cudaStream_t;

inline const char* cudaGetErrorString(cudaError_t) { return "mock CUDA error"; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = new MockStream();
  return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
  delete s;
  return cudaSuccess;
}
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
  *p = std::malloc(n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaMalloc(void** p, size_t n) { return cudaHostAlloc(p, n, 0); }
inline cudaError_t cudaFreeHost(void* p) {
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) { return cudaFreeHost(p); }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = new MockEvent();
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  long gen;
  {
    std::lock_guard<std::mutex> lk(e->mu);
    gen = ++e->recorded;
  }
  s->push([e, gen] {
    // notify under the lock: a waiter (e.g. the stack event of cudaStreamSynchronize) may destroy the event as soon as
    // it observes `done`, which it cannot do before this thread has left the critical section
    std::lock_guard<std::mutex> lk(e->mu);
    if (e->done < gen) e->done = gen;
    e->cv.notify_all();
  });
  return cudaSuccess;
}
inline void mock_event_wait(cudaEvent_t e, long gen) {
  std::unique_lock<std::mutex> lk(e->mu);
  e->cv.wait(lk, [&] { return e->done >= gen; });
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
  long gen;
  {
    std::lock_guard<std::mutex> lk(e->mu);
    gen = e->recorded;
  }
  if (gen > 0) s->push([e, gen] { mock_event_wait(e, gen); });
  return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t e) {
  long gen;
  {
    std::lock_guard<std::mutex> lk(e->mu);
    gen = e->recorded;
  }
  if (gen > 0) mock_event_wait(e, gen);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t s) {
  s->push([=] {
    const long j = mock_jitter_us().load();
    if (j > 0) {
      s->rng ^= s->rng << 13; s->rng ^= s->rng >> 7; s->rng ^= s->rng << 17;
      std::this_thread::sleep_for(std::chrono::microseconds((long)(s->rng % (uint64_t)(j + 1))));
    }
    std::memcpy(dst, src, n);
  });
  return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t s) {
  MockEvent e;
  cudaEventRecord(&e, s);
  return cudaEventSynchronize(&e);
}
