// audio8_b200 — C-ABI plumbing: version, thread-local error string, launch accounting.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -3;
  }
  return 0;
}
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("A8_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
static thread_local const unsigned long long* g_seed_src = nullptr;
const unsigned long long* seed_source() { return g_seed_src; }
}  // namespace a8

extern "C" int a8_set_seed_source(const void* dev_u64) {
  a8::g_seed_src = static_cast<const unsigned long long*>(dev_u64);
  return 0;
}
extern "C" int a8_version(void) { return A8_ABI_VERSION; }
extern "C" const char* a8_last_error(void) { return a8::g_err; }
extern "C" int64_t a8_launch_count(void) { return (int64_t)a8::g_launches.load(); }
extern "C" int a8_launch_count_add(int64_t n) {
  a8::g_launches.fetch_add(n, std::memory_order_relaxed);
  return 0;
}
