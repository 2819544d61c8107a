// C-ABI plumbing shared by every entry point: thread-local error string, launch checks, version.
#include "m0_common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void m0_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int m0_check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return M0_OK;
  m0_set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return M0_ERR_CUDA;
}

static unsigned long long g_launches = 0;

int m0_check_launch(const char* what) {
  g_launches++;
  return m0_check_cuda(cudaGetLastError(), what);
}

extern "C" {

const char* m0_last_error(void) { return g_err; }

int m0_version(void) { return 100; }

// Number of kernel launches issued by this library so far (bench.py reports the per-step count).
unsigned long long m0_launch_count(void) { return g_launches; }

// Number of CUDA devices visible to the library (<= 0 means the product cannot run: there is no
// CPU path behind this ABI).
int m0_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int m0_device_sm_count(int device) {
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  return v;
}

}  // extern "C"
