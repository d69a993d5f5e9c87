// Library-level entry points of the C ABI: version, error string, struct sizes.
#include <stdarg.h>

#include "ols_common.cuh"

static thread_local char g_ols_error[512] = "";

void ols_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_ols_error, sizeof(g_ols_error), fmt, ap);
  va_end(ap);
}

#include <atomic>
static std::atomic<unsigned long long> g_ols_launches{0};
void ols_note_launch(int n) { g_ols_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
extern "C" uint64_t ols_kernel_launches(void) { return g_ols_launches.load(std::memory_order_relaxed); }

int ols_sm_count(void) {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 148;
  }
  if (dev >= 0 && dev < 64 && cache[dev].load(std::memory_order_relaxed) > 0) return cache[dev].load(std::memory_order_relaxed);
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = 148;
  }
  if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

extern "C" int ols_abi_version(void) { return OLS_ABI_VERSION; }

extern "C" const char* ols_last_error(void) { return g_ols_error; }

extern "C" size_t ols_struct_bytes(int which) {
  switch (which) {
    case 0: return sizeof(ols_log_t);
    case 1: return sizeof(ols_queues_t);
    case 2: return sizeof(ols_replicas_t);
    case 3: return sizeof(ols_step_out_t);
    case 4: return sizeof(ols_run_t);
    case 5: return sizeof(ols_queue_run_t);
    case 6: return sizeof(ols_exo_run_t);
    default: return 0;
  }
}
