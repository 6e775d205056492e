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

static int g_sm_budget = 0;
int dwn_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return (g_sm_budget > 0 && g_sm_budget < n) ? g_sm_budget : n;
}
// Number of SMs the persistent kernels (GEMM tile loop, one-wave grids) size themselves for; 0 = all.  Data-parallel
// backward leaves a few SMs to the concurrently running NCCL kernels: a persistent CTA that cannot become resident
// next to an NCCL CTA would otherwise wait for the whole collective.
extern "C" int dwn_set_sm_budget(int n) {
  g_sm_budget = n > 0 ? n : 0;
  return 0;
}

extern "C" const char* dwn_last_error() { return g_err; }
extern "C" int dwn_abi_version() { return 1; }
extern "C" int dwn_sm_count() { return dwn_num_sms(); }
