// C-ABI plumbing shared by all translation units.
#include "dwn_common.cuh"
#include <string.h>

static thread_local char g_err[1024] = "";

int dwn_fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}

int dwn_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

extern "C" const char* dwn_last_error() { return g_err; }
extern "C" int dwn_abi_version() { return 1; }
extern "C" int dwn_sm_count() { return dwn_num_sms(); }
