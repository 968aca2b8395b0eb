// Library management: version, error text, launch accounting.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include "pu_common.cuh"

namespace pu {

static std::mutex g_err_mu;
static char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PU_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;  // measured neutral on B200 (2.02 vs 2.01 ms/step): off unless PU_PDL=1
  }
  return v == 1;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  count_launch(1);
  return PU_OK;
}

}  // namespace pu

extern "C" {

int pu_version(void) { return 100; }  // 0.1.0

const char* pu_last_error(void) { return pu::g_err; }

long long pu_launch_count(void) { return pu::g_launches.load(); }

void pu_reset_launch_count(void) { pu::g_launches.store(0); }

}  // extern "C"
