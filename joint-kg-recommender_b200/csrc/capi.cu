// Error plumbing and small queries of the kgrec_b200 C ABI.
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace kgrec {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}
}  // namespace kgrec

extern "C" int kgrec_abi_version(void) { return KGREC_ABI_VERSION; }
extern "C" const char* kgrec_last_error(void) { return kgrec::g_err; }
extern "C" int kgrec_sm_count(void) { return kgrec::sm_count(); }
