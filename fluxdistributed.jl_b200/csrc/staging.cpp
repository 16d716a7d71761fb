// Pinned staging ring with a producer thread, a dedicated copy stream and event hand-off.
//
//   producer thread (C++, no GIL):  fill pinned slot i  ->  ready queue
//   consumer (training thread):     next() pops a ready slot, enqueues cudaMemcpyAsync(H2D) of x and y on the
//                                   side stream, records `copied[i]`, makes the compute stream wait on it and
//                                   returns the device pointers; release happens when next() is called again
//                                   (`consumed[i]` recorded on the compute stream guards buffer reuse).
// The H2D copy of step i+1 therefore overlaps the compute of step i.  The built-in producer fabricates
// deterministic synthetic uint8 images + int64 labels (benchmark configs use synthetic data); a Python
// producer can instead fill slots through host_ptr()/commit().
#include <cuda_runtime.h>
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
  PinnedRing(int device, size_t x_bytes, size_t y_bytes, int depth)
      : device_(device), xb_(x_bytes), yb_(y_bytes), depth_(depth < 2 ? 2 : depth) {
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
  uintptr_t host_x(int s) const { return (uintptr_t)hx_[s]; }
  uintptr_t host_y(int s) const { return (uintptr_t)hy_[s]; }
  void commit(int s) {
    std::lock_guard<std::mutex> lk(mu_);
    ready_.push_back(s);
    cv_.notify_all();
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
        commit(slot);
      }
    });
  }

  // ---- consumer: returns (slot, dev_x, dev_y); the compute stream `stream` will wait for the copy
  py::tuple next(uintptr_t stream) {
    int s;
    {
      py::gil_scoped_release nogil;
      std::unique_lock<std::mutex> lk(mu_);
      if (last_ >= 0) {  // previous slot: its device buffers are consumed by work already queued on `stream`
        cudaSetDevice(device_);
        cudaEventRecord(consumed_[last_], (cudaStream_t)stream);
        free_.push_back(last_);
        last_ = -1;
        cv_.notify_all();
      }
      cv_.wait(lk, [&] { return !ready_.empty() || stopping_; });
      if (ready_.empty()) throw std::runtime_error("staging ring stopped");
      s = ready_.front();
      ready_.pop_front();
    }
    RT_CHECK(cudaSetDevice(device_));
    RT_CHECK(cudaStreamWaitEvent(copy_, consumed_[s], 0));
    RT_CHECK(cudaMemcpyAsync(dx_[s], hx_[s], xb_, cudaMemcpyHostToDevice, copy_));
    if (yb_) RT_CHECK(cudaMemcpyAsync(dy_[s], hy_[s], yb_, cudaMemcpyHostToDevice, copy_));
    RT_CHECK(cudaEventRecord(copied_[s], copy_));
    RT_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, copied_[s], 0));
    last_ = s;
    return py::make_tuple(s, (uintptr_t)dx_[s], (uintptr_t)dy_[s]);
  }

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
  int device_;
  size_t xb_, yb_;
  int depth_;
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
      .def(py::init<int, size_t, size_t, int>(), py::arg("device"), py::arg("x_bytes"), py::arg("y_bytes"),
           py::arg("depth") = 3)
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
