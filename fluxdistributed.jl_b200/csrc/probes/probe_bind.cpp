// Bindings of the hardware probes (module fluxdist._Cprobe): built next to the product extension but never linked
// into it; used by scripts/probe_*.py only.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <atomic>
#include <stdexcept>
#include <string>

int fd_probe_shift32(const void* A, const void* B, float* D, int rows, int shift, int sbo, int mn_major, cudaStream_t s);
int fd_probe_mma_rate(long long* out, int bn, int reps, int shift, int sbo, int grid, int flavor, cudaStream_t s);
int fd_probe_shift(const void* A, const void* B, float* D, int rows, int shift, int sbo, int base_off, int mn_major,
                   int nk16, cudaStream_t s);
int fd_probe_mma_rate2(long long* out, int bn, int form, int cta2, int reps, int grid, cudaStream_t s);
int fd_probe_tma_bw(const void* buf, long rows, long long* cycles, int grid, int boxes_per_region, int share, int reps,
                    int mc, cudaStream_t s);

// the probes use the product's error / launch-count hooks; this module carries its own copies
static std::string g_err;
void fd_set_error(const char* msg, const char* file, int line) { g_err = std::string(msg) + " (" + file + ":" + std::to_string(line) + ")"; }
const char* fd_last_error() { return g_err.c_str(); }
static std::atomic<long> g_n{0};
void fd_count_launch() { g_n.fetch_add(1); }
long fd_launch_count() { return g_n.load(); }

namespace {
inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
inline void check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + " failed: " + fd_last_error());
}
}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("probe_shift32", [](at::Tensor A, at::Tensor B, at::Tensor D, int shift, int sbo, bool mn_major) {
    c10::cuda::CUDAGuard guard(A.device());
    check(fd_probe_shift32(A.data_ptr(), B.data_ptr(), D.data_ptr<float>(), (int)A.size(0), shift, sbo, mn_major, cur_stream()),
          "probe_shift32");
  });
  m.def("probe_mma_rate", [](at::Tensor out, int bn, int reps, int shift, int sbo, int grid, int flavor) {
    c10::cuda::CUDAGuard guard(out.device());
    check(fd_probe_mma_rate(reinterpret_cast<long long*>(out.data_ptr<int64_t>()), bn, reps, shift, sbo, grid, flavor, cur_stream()),
          "probe_mma_rate");
  });
  m.def("probe_shift", [](at::Tensor A, at::Tensor B, at::Tensor D, int shift, int sbo, int base_off, bool mn_major, int nk16) {
    c10::cuda::CUDAGuard guard(A.device());
    check(fd_probe_shift(A.data_ptr(), B.data_ptr(), D.data_ptr<float>(), (int)A.size(0), shift, sbo, base_off, mn_major, nk16,
                         cur_stream()), "probe_shift");
  });
  m.def("probe_mma_rate2", [](at::Tensor out, int bn, int form, bool cta2, int reps, int grid) {
    c10::cuda::CUDAGuard guard(out.device());
    check(fd_probe_mma_rate2(reinterpret_cast<long long*>(out.data_ptr<int64_t>()), bn, form, cta2 ? 1 : 0, reps, grid, cur_stream()),
          "probe_mma_rate2");
  });
  m.def("probe_tma_bw", [](at::Tensor buf, at::Tensor cycles, int grid, int boxes_per_region, int share, int reps, bool mc) {
    c10::cuda::CUDAGuard guard(buf.device());
    check(fd_probe_tma_bw(buf.data_ptr(), (long)buf.size(0), reinterpret_cast<long long*>(cycles.data_ptr<int64_t>()), grid,
                          boxes_per_region, share, reps, mc ? 1 : 0, cur_stream()), "probe_tma_bw");
  });
}
