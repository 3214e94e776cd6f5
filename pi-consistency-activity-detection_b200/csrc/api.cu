// error plumbing + misc entry points of the C ABI
#include "common.cuh"
#include "../../include/b200caps.h"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";
long long b2c_launches_add(long long n);

int b2c_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int b2c_cuda_check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}
B2C_API const char* b2c_last_error(void) { return g_err; }
B2C_API int b2c_version(void) { return 100; }
B2C_API long long b2c_launch_count(void) { return b2c_launches_add(0); }

// precision mode (see common.cuh)
static int g_precision = 0;
int b2c_precision() { return g_precision; }
B2C_API int b2c_set_precision(int32_t mode) {
  if (mode != 0 && mode != 1) return b2c_fail(-1, "set_precision: mode=%d (0 = bf16, 1 = fp32 activations / tf32 operands)", mode);
  g_precision = mode;
  return 0;
}
B2C_API int b2c_get_precision(void) { return g_precision; }
