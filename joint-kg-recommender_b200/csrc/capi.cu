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
float l2_keep_fraction(double table_bytes) {
  static thread_local int cached_dev = -1;
  static thread_local double l2 = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1.f;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev) != cudaSuccess || v <= 0) v = 64 << 20;
    l2 = v;
    cached_dev = dev;
  }
  if (table_bytes <= 0) return 1.f;
  const double f = 0.6 * l2 / table_bytes;      // leave 40 % of L2 to the streams
  return static_cast<float>(f >= 1.0 ? 1.0 : (f < 0.05 ? 0.05 : f));
}
}  // namespace kgrec

extern "C" int kgrec_abi_version(void) { return KGREC_ABI_VERSION; }
extern "C" const char* kgrec_last_error(void) { return kgrec::g_err; }
extern "C" int kgrec_sm_count(void) { return kgrec::sm_count(); }
