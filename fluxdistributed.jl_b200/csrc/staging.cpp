// Pinned staging ring: fixed cudaHostAlloc slots, a dedicated copy stream and event hand-off (replaces the
// reference's pageable gpu(shard) upload inside the loader closure, src/ddp_tasks.jl:278-283).
//
//   producer (a C++ thread for synthetic data, or a Python loader thread through acquire()/commit()):
//       fill pinned slot i -> ready queue
//   consumer (training thread), next(compute_stream):
//       1. release the slot handed out last time: record `consumed[i]` on the compute stream (its device buffers are
//          read by work queued so far) and wait on the HOST for `copied[i]` before the pinned slot goes back to the
//          producer — the async H2D copy must have read the host buffer before anyone overwrites it (round-1 race);
//       2. top up the copies in flight (up to `prefetch` slots ahead): cudaStreamWaitEvent(copy, consumed[s]) ->
//          cudaMemcpyAsync(H2D) of x and y on the copy stream -> record `copied[s]`;
//       3. hand out the oldest copy in flight: the compute stream waits on `copied[s]`.
// Copies are ALSO issued from commit() (the producer's thread) whenever fewer than `prefetch` are in flight, so a
// batch that becomes ready while the consumer is busy queueing step i starts its transfer immediately instead of at
// the next next() call.  The H2D copy of step i+1 (and i+2) therefore overlaps step i's compute whichever side is
// ahead.  All queue state (free / ready / in flight) is guarded by one mutex; the CUDA calls made under it are
// asynchronous enqueues.
#ifdef FD_STAGING_MOCK_CUDA
#include "mock/staging_mock_cuda.h"  // host-only protocol tests (tests/test_staging_mock.py); never defined for the product
#else
#include <cuda_runtime.h>
#endif
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "staging.h"

namespace py = pybind11;

namespace {

#define RT_CHECK(x)                                                                                 \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

class PinnedRing {
 public:
  PinnedRing(int device, size_t x_bytes, size_t y_bytes, int depth, int prefetch = 2)
      : device_(device), xb_(x_bytes), yb_(y_bytes), depth_(depth < 3 ? 3 : depth),
        prefetch_(prefetch < 1 ? 1 : (prefetch > depth - 1 ? depth - 1 : prefetch)) {
    RT_CHECK(cudaSetDevice(device_));
    RT_CHECK(cudaStreamCreateWithFlags(&copy_, cudaStreamNonBlocking));
    hx_.resize(depth_); hy_.resize(depth_); dx_.resize(depth_); dy_.resize(depth_);
    copied_.resize(depth_); consumed_.resize(depth_);
    for (int i = 0; i < depth_; ++i) {
      RT_CHECK(cudaHostAlloc(&hx_[i], xb_, cudaHostAllocDefault));
      RT_CHECK(cudaHostAlloc(&hy_[i], yb_ ? yb_ : 8, cudaHostAllocDefault));
      RT_CHECK(cudaMalloc(&dx_[i], xb_));
      RT_CHECK(cudaMalloc(&dy_[i], yb_ ? yb_ : 8));
      RT_CHECK(cudaEventCreateWithFlags(&copied_[i], cudaEventDisableTiming));
      RT_CHECK(cudaEventCreateWithFlags(&consumed_[i], cudaEventDisableTiming));
      free_.push_back(i);
    }
  }
  ~PinnedRing() {
    stop();
    cudaSetDevice(device_);
    cudaStreamSynchronize(copy_);
    for (int i = 0; i < depth_; ++i) {
      cudaFreeHost(hx_[i]); cudaFreeHost(hy_[i]); cudaFree(dx_[i]); cudaFree(dy_[i]);
      cudaEventDestroy(copied_[i]); cudaEventDestroy(consumed_[i]);
    }
    cudaStreamDestroy(copy_);
  }

  // ---- python-side producer protocol
  int acquire() {  // blocks until a host slot is free; returns its index
    py::gil_scoped_release nogil;
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !free_.empty() || stopping_; });
    if (stopping_) return -1;
    int s = free_.front();
    free_.pop_front();
    return s;
  }
  uintptr_t dev_x(int s) const { return (uintptr_t)dx_[s]; }
  uintptr_t dev_y(int s) const { return (uintptr_t)dy_[s]; }
  uintptr_t host_x(int s) const { return (uintptr_t)hx_[s]; }
  uintptr_t host_y(int s) const { return (uintptr_t)hy_[s]; }
  void commit(int s) {
    py::gil_scoped_release nogil;
    RT_CHECK(cudaSetDevice(device_));
    commit_slot(s);
  }

  // ---- built-in synthetic producer (uint8 NHWC pixels, int64 labels)
  void start_synthetic(long steps, int batch, int nclasses, uint64_t seed) {
    stop();
    stopping_ = false;
    bank_.resize(xb_ / (batch > 0 ? batch : 1) * 2 + 64);
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
    for (auto& b : bank_) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      b = (uint8_t)(s >> 32);
    }
    worker_ = std::thread([this, steps, batch, nclasses, seed] {
      cudaSetDevice(device_);
      uint64_t st = seed + 12345;
      const size_t img = xb_ / batch;
      for (long it = 0; it < steps && !stopping_; ++it) {
        int slot;
        {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return !free_.empty() || stopping_; });
          if (stopping_) return;
          slot = free_.front();
          free_.pop_front();
        }
        auto* x = (uint8_t*)hx_[slot];
        auto* y = (long long*)hy_[slot];
        for (int b = 0; b < batch; ++b) {
          st ^= st << 13; st ^= st >> 7; st ^= st << 17;
          size_t off = (size_t)(st % (bank_.size() - img));
          std::memcpy(x + (size_t)b * img, bank_.data() + off, img);
          if (yb_) y[b] = (long long)((st >> 20) % (uint64_t)nclasses);
        }
        try {
          commit_slot(slot);
        } catch (const std::exception&) {  // a failed enqueue ends the stream of batches; next() reports "stopped"
          std::lock_guard<std::mutex> lk(mu_);
          stopping_ = true;
          cv_.notify_all();
          return;
        }
      }
    });
  }

  // ---- consumer: returns (slot, dev_x, dev_y); the compute stream `stream` will wait for the copy
  py::tuple next(uintptr_t stream) {
    py::gil_scoped_release nogil;
    RT_CHECK(cudaSetDevice(device_));
    if (last_ >= 0) {
      // previous slot: its device buffers are consumed by work already queued on `stream`; its pinned buffers may be
      // refilled only once the copy engine has read them
      RT_CHECK(cudaEventRecord(consumed_[last_], (cudaStream_t)stream));
      RT_CHECK(cudaEventSynchronize(copied_[last_]));
      std::lock_guard<std::mutex> lk(mu_);
      free_.push_back(last_);
      last_ = -1;
      cv_.notify_all();
    }
    int s;
    {
      std::unique_lock<std::mutex> lk(mu_);
      issue_ready_locked();
      if (inflight_.empty()) {  // nothing ready yet: block for the producer (its commit() issues the copy)
        cv_.wait(lk, [&] { return !inflight_.empty() || !ready_.empty() || stopping_; });
        issue_ready_locked();
        if (inflight_.empty()) throw std::runtime_error("staging ring stopped");
      }
      s = inflight_.front();
      inflight_.pop_front();
      last_ = s;
      issue_ready_locked();  // the copy of the NEXT batch is queued before this batch's compute
    }
    RT_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, copied_[s], 0));
    const uintptr_t px = (uintptr_t)dx_[s], py_ = (uintptr_t)dy_[s];
    py::gil_scoped_acquire gil;
    return py::make_tuple(s, px, py_);
  }

  // wall-clock independent evidence of overlap for tests: copies issued so far, and copies that had ALREADY been
  // issued when their batch was handed out one step earlier (= prefetched ahead of the compute that precedes them)
  long copies_issued() const { return issued_.load(); }

  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stopping_ = true;
      cv_.notify_all();
    }
    if (worker_.joinable()) worker_.join();
  }
  size_t x_bytes() const { return xb_; }
  size_t y_bytes() const { return yb_; }
  int depth() const { return depth_; }

 private:
  void commit_slot(int s) {
    std::lock_guard<std::mutex> lk(mu_);
    ready_.push_back(s);
    issue_ready_locked();
    cv_.notify_all();
  }

  // issue H2D copies for ready slots (FIFO) until `prefetch_` are in flight; mu_ held, current device = device_
  void issue_ready_locked() {
    while ((int)inflight_.size() < prefetch_ && !ready_.empty()) {
      const int s = ready_.front();
      ready_.pop_front();
      RT_CHECK(cudaStreamWaitEvent(copy_, consumed_[s], 0));  // device buffers of slot s no longer read by compute
      RT_CHECK(cudaMemcpyAsync(dx_[s], hx_[s], xb_, cudaMemcpyHostToDevice, copy_));
      if (yb_) RT_CHECK(cudaMemcpyAsync(dy_[s], hy_[s], yb_, cudaMemcpyHostToDevice, copy_));
      RT_CHECK(cudaEventRecord(copied_[s], copy_));
      inflight_.push_back(s);
      ++issued_;
    }
  }

  int device_;
  size_t xb_, yb_;
  int depth_;
  int prefetch_;
  std::deque<int> inflight_;  // guarded by mu_ (producer and consumer both issue copies)
  std::atomic<long> issued_{0};
  cudaStream_t copy_{};
  std::vector<void*> hx_, hy_, dx_, dy_;
  std::vector<cudaEvent_t> copied_, consumed_;
  std::deque<int> free_, ready_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::thread worker_;
  std::atomic<bool> stopping_{false};
  std::vector<uint8_t> bank_;
  int last_ = -1;
};

}  // namespace

void fd_bind_staging(py::module_& m) {
  py::class_<PinnedRing>(m, "PinnedRing")
      .def(py::init<int, size_t, size_t, int, int>(), py::arg("device"), py::arg("x_bytes"), py::arg("y_bytes"),
           py::arg("depth") = 4, py::arg("prefetch") = 2)
      .def("copies_issued", &PinnedRing::copies_issued)
      .def("dev_x", &PinnedRing::dev_x)
      .def("dev_y", &PinnedRing::dev_y)
      .def("acquire", &PinnedRing::acquire)
      .def("host_x", &PinnedRing::host_x)
      .def("host_y", &PinnedRing::host_y)
      .def("commit", &PinnedRing::commit)
      .def("start_synthetic", &PinnedRing::start_synthetic, py::arg("steps"), py::arg("batch"), py::arg("nclasses"),
           py::arg("seed") = 0)
      .def("next", &PinnedRing::next)
      .def("stop", &PinnedRing::stop)
      .def_property_readonly("x_bytes", &PinnedRing::x_bytes)
      .def_property_readonly("y_bytes", &PinnedRing::y_bytes)
      .def_property_readonly("depth", &PinnedRing::depth);
}
