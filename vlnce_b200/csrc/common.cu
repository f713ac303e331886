#include "common.cuh"

#include <atomic>
#include <stdarg.h>

namespace vlnce {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return VLNCE_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return VLNCE_ERR_CUDA_BASE + (int)e;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace vlnce

extern "C" const char* vlnce_version(void) { return "vlnce_b200 0.1 (sm_100a)"; }
extern "C" const char* vlnce_last_error(void) { return vlnce::g_err; }
extern "C" long long vlnce_launch_count(void) { return vlnce::g_launches.load(); }
