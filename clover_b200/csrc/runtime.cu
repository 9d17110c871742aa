// Error slot, launch accounting and device queries shared by every C-ABI entry point.
#include <atomic>
#include <cstdarg>
#include <cstdio>

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

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace clv

extern "C" const char* clv_last_error(void) { return clv::g_err; }
extern "C" int clv_version(void) { return 100; }
extern "C" long long clv_launch_count(void) { return clv::g_launches.load(std::memory_order_relaxed); }
