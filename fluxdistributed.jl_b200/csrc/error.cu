#include <mutex>
#include <string>

#include <atomic>

#include "common.cuh"

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
