#include <mutex>
#include <string>

#include <atomic>

#include "common.cuh"

// Kernel classes launched with the PDL attribute when FLUXDIST_PDL is not set (see common.cuh).  Measured on B200,
// ResNet-34 b96 step (profiles/r2/pdl_ab.txt): the tcgen05 kernels as dependents (their barrier init / TMEM allocation /
// descriptor prefetch and launch latency hide behind the previous kernel's tail) gain 5.9 %; the memory-bound BN
// kernels as dependents LOSE (their CTAs become resident next to the running tensor-core kernel and wait there), so
// they only act as primaries (they execute launch_dependents, their own launches stay stream-ordered).
#define FD_PDL_DEFAULT_MASK ((1 << PDL_IGEMM) | (1 << PDL_IGEMM_WGRAD) | (1 << PDL_HALO))

static std::mutex g_err_mu;
static std::string g_err;

void fd_set_error(const char* msg, const char* file, int line) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = std::string(msg) + " (" + file + ":" + std::to_string(line) + ")";
}
const char* fd_last_error() {
  std::lock_guard<std::mutex> lk(g_err_mu);
  static thread_local std::string copy;
  copy = g_err;
  return copy.c_str();
}

static std::atomic<long> g_launches{0};
void fd_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long fd_launch_count() { return g_launches.load(std::memory_order_relaxed); }

#include <cstdlib>
// -1 = not read yet; the environment is consulted once, fd_set_pdl() overrides it (tests / A-B runs)
static std::atomic<int> g_pdl{-1};
int fd_pdl_mask() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("FLUXDIST_PDL");
    v = e ? (int)strtol(e, nullptr, 0) & 0xffff : FD_PDL_DEFAULT_MASK;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v;
}
void fd_set_pdl(int mask) { g_pdl.store(mask & 0xffff, std::memory_order_relaxed); }
