#include "common.cuh"

#include <atomic>
#include <mutex>
#include <unordered_set>
#include <stdarg.h>
#include <stdlib.h>

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

static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("VLNCE_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
void set_pdl(int on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }

bool carveout_first_use(const void* kernel) {
  static std::mutex mu;
  static std::unordered_set<const void*> seen;
  static int enabled = -1;
  std::lock_guard<std::mutex> lock(mu);
  if (enabled < 0) {
    const char* e = getenv("VLNCE_CARVEOUT");
    enabled = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  if (!enabled) return false;
  return seen.insert(kernel).second;
}

static std::atomic<int> g_lstm_rows{-1};
int lstm_rows_per_cta(int R) {
  int v = g_lstm_rows.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("VLNCE_LSTM_NR");
    v = (e != nullptr && (e[0] == '1' || e[0] == '2' || e[0] == '4') && e[1] == 0) ? e[0] - '0' : 0;
    g_lstm_rows.store(v, std::memory_order_relaxed);
  }
  (void)R;
  return (v == 4 || v == 2) ? v : 1;
}
void set_lstm_rows(int nr) { g_lstm_rows.store((nr == 4 || nr == 2 || nr == 1) ? nr : 0, std::memory_order_relaxed); }

}  // namespace vlnce

extern "C" int vlnce_set_lstm_rows(int rows_per_cta) {
  vlnce::set_lstm_rows(rows_per_cta);
  return VLNCE_OK;
}
extern "C" const char* vlnce_version(void) { return "vlnce_b200 0.1 (sm_100a)"; }
extern "C" const char* vlnce_last_error(void) { return vlnce::g_err; }
extern "C" long long vlnce_launch_count(void) { return vlnce::g_launches.load(); }
namespace vlnce { void set_pdl(int on); }
extern "C" int vlnce_set_pdl(int on) {
  vlnce::set_pdl(on);
  return VLNCE_OK;
}
extern "C" long long vlnce_get_pdl(void) { return vlnce::pdl_enabled() ? 1 : 0; }

// Debug aid: one thread writes %globaltimer (ns) into *slot when the stream reaches this point — a timeline probe
// that can be captured into a CUDA graph on any branch (tools/timeline_probe.py).
namespace vlnce {
__global__ void stamp_kernel(long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = (long long)t;
}
}  // namespace vlnce
extern "C" int vlnce_debug_stamp(long long* slot, cudaStream_t stream) {
  VLNCE_CHECK_ARG(slot != nullptr, "debug_stamp: null slot");
  vlnce::stamp_kernel<<<1, 1, 0, stream>>>(slot);
  VLNCE_LAUNCH_CHECK("stamp_kernel");
  return VLNCE_OK;
}

// ---- workspace sizes (vlnce_<op>_workspace_bytes): host-only arithmetic, no CUDA call.  Every entry point
// that takes a workspace also works with a smaller one (fewer split-K / split-row partials) or none; these return the
// size with which the launch plan is never narrowed.
extern "C" long long vlnce_sgemm_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const long long splits = K / 64 < 1 ? 1 : (K / 64 > 32 ? 32 : K / 64);   // gemm.cu: <= 32 partials of one (M,N) tile set
  return 16384 /* arrival counters */ + splits * (long long)M * N * (long long)sizeof(float);
}
extern "C" long long vlnce_colsum_workspace_bytes(int M, int N) {
  if (M < 256 || N <= 0) return 0;
  const long long S = M / 128 > 64 ? 64 : M / 128;
  return 16384 + S * N * (long long)sizeof(float);
}
extern "C" long long vlnce_groupnorm_workspace_bytes(int N, int HW, int C, int groups) {
  if (N <= 0 || HW <= 0 || C <= 0 || groups <= 0) return 0;
  const long long S = ((long long)HW * C + 4095) / 4096;   // trunk_ops.cu: 4096-element slices of an image
  return (long long)N * S * groups * 2 * (long long)sizeof(float);
}
extern "C" long long vlnce_clip_grad_norm_workspace_bytes(long long n) {
  (void)n;
  return 1026 * (long long)sizeof(float);   // <= 1024 partial sums of squares, the norm, the clip factor
}
