// Library-level entry points: version, error string, device query.
#include "common.cuh"
#include "../../include/icepy4d_b200.h"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void i4d_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int i4d_num_sms() {
  static int sms = -1;
  if (sms < 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      sms = v;
    else
      return 148;
  }
  return sms;
}

extern "C" __attribute__((visibility("default"))) int i4d_version(void) { return I4D_VERSION; }
extern "C" __attribute__((visibility("default"))) const char* i4d_last_error(void) { return g_err; }
extern "C" __attribute__((visibility("default"))) int i4d_device_sm_count(void) {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return v;
}
