// Python module `_Cstaging_mock`: the product's PinnedRing (csrc/staging.cpp, compiled with FD_STAGING_MOCK_CUDA against
// staging_mock_cuda.h) plus the few hooks a host-only stress test needs: a mock "compute stream", a check op that runs ON
// that stream (so it sees the device buffer exactly when real compute would) and the jitter knob of the mock copy engine.
#include <pybind11/pybind11.h>

#include <atomic>

#include "../staging.h"
#include "staging_mock_cuda.h"

namespace py = pybind11;

static std::atomic<long> g_errors{0};
static std::atomic<long> g_checked{0};

#ifndef FD_MOCK_MODNAME
#define FD_MOCK_MODNAME _Cstaging_mock
#endif

PYBIND11_MODULE(FD_MOCK_MODNAME, m) {
  m.doc() = "host-only build of the pinned staging ring over a mock CUDA runtime (tests only)";
  fd_bind_staging(m);
  m.def("stream_create", []() { return (uintptr_t) new MockStream(); });
  m.def("stream_destroy", [](uintptr_t h) { delete reinterpret_cast<MockStream*>(h); });
  m.def("stream_sync", [](uintptr_t h) {
    py::gil_scoped_release nogil;
    cudaStreamSynchronize(reinterpret_cast<MockStream*>(h));
  });
  // queue "compute" on stream h: after `delay_us`, every byte of [ptr, ptr + n) must equal `expect`
  m.def("enqueue_check", [](uintptr_t h, uintptr_t ptr, size_t n, int expect, long delay_us) {
    reinterpret_cast<MockStream*>(h)->push([=] {
      if (delay_us > 0) std::this_thread::sleep_for(std::chrono::microseconds(delay_us));
      const unsigned char* p = reinterpret_cast<const unsigned char*>(ptr);
      long bad = 0;
      for (size_t i = 0; i < n; ++i) bad += (p[i] != (unsigned char)expect);
      if (bad) g_errors.fetch_add(1);
      g_checked.fetch_add(1);
    });
  });
  m.def("errors", []() { return g_errors.load(); });
  m.def("checked", []() { return g_checked.load(); });
  m.def("reset_counters", []() { g_errors = 0; g_checked = 0; });
  m.def("set_jitter_us", [](long us) { mock_jitter_us().store(us); });
}
