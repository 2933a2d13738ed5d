// error.cu — thread-local error reporting for the C-ABI (tio_last_error).
#include <stdarg.h>

#include "common.cuh"

namespace tio {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace tio

extern "C" const char* tio_last_error(void) { return tio::g_error; }
extern "C" int tio_abi_version(void) { return TIO_ABI_VERSION; }
