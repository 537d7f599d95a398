// Error plumbing + version for the C-ABI (include/icl_b200.h).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void icl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int icl_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    icl_set_error("%s: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

ICL_API const char* icl_last_error(void) { return g_err; }
ICL_API int icl_version(void) { return 100; }

// Number of kernels launched through this library since load (bench.py's gpu_launches claim).
static unsigned long long g_launches = 0;
void icl_count_launch(int n) { g_launches += (unsigned long long)n; }
ICL_API unsigned long long icl_launch_count(void) { return g_launches; }
