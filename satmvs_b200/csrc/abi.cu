// abi.cu — version and thread-local error text of libsatmvs_b200.so
#include "common.cuh"
#include <cstring>

namespace satmvs {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace satmvs

extern "C" {
int satmvs_abi_version(void) { return SATMVS_ABI_VERSION; }
const char* satmvs_last_error(void) { return satmvs::g_error; }
}
