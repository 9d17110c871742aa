// Error slot, launch accounting and device queries shared by every C-ABI entry point.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <map>
#include <utility>

#include "common.cuh"
#include "clover_b200.h"

namespace clv {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}

// SM count of the CURRENT device, cached per device (forward runs on the Python main thread, backward on autograd's
// per-device worker threads: the cache is a table of atomics, not one process-wide static).
int num_sms() {
  constexpr int MAX_DEV = 64;
  static std::atomic<int> cache[MAX_DEV];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: raise it only when a launch needs more
// than what was already granted for that (kernel, device), remembered under a mutex (init-once, read-mostly).
int ensure_dynamic_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> granted;
  if (bytes <= 48 * 1024) return 0;                                  // the default limit needs no opt-in
  int dev = 0;
  CLV_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  int& have = granted[std::make_pair(kernel, dev)];
  if (bytes <= have) return 0;
  CLV_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  have = bytes;
  return 0;
}

// Tunables: plain process-wide integers with production defaults, changed only through clv_set_tunable (tools / tests).
// No environment variable is read anywhere in the library.
static std::atomic<long long> g_tunables[TUNE_COUNT] = {};
static std::atomic<bool> g_tunable_set[TUNE_COUNT] = {};
static const char* const TUNE_NAMES[TUNE_COUNT] = {"gemm_bn256_min_units", "w7_pipe", "w7_bwd2", "w7_dbias_acc", "w7_dbias_acc_min_mb", "w7_fwd2", "gemm_tma_store", "w7_l2_hint", "gemm_spec", "w7_bwd_early", "gemm_box", "w7_fwd_early", "w7_fwd_dbg", "w7_fwd_pvsplit", "w7_fwd_qtile"};

long long tunable(int id, long long dflt) {
  return g_tunable_set[id].load(std::memory_order_relaxed) ? g_tunables[id].load(std::memory_order_relaxed) : dflt;
}

}  // namespace clv

extern "C" const char* clv_last_error(void) { return clv::g_err; }
extern "C" int clv_version(void) { return 100; }
extern "C" int clv_set_tunable(const char* name, long long value) {
  for (int i = 0; i < clv::TUNE_COUNT; ++i)
    if (name && !strcmp(name, clv::TUNE_NAMES[i])) {
      clv::g_tunables[i].store(value, std::memory_order_relaxed);
      clv::g_tunable_set[i].store(value != -1, std::memory_order_relaxed);     // -1 restores the built-in default
      return 0;
    }
  clv::set_error("clv_set_tunable: unknown tunable '%s'", name ? name : "(null)");
  return 1;
}
extern "C" long long clv_launch_count(void) { return clv::g_launches.load(std::memory_order_relaxed); }
