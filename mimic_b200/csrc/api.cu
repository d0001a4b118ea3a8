// Library-wide state: error string, implementation tag, device query.
#include <cstring>

#include "common.cuh"

namespace mimic {

static thread_local char g_err[512] = "";
static thread_local char g_impl[64] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void set_attn_impl(const char* name) {
  strncpy(g_impl, name, sizeof(g_impl) - 1);
  g_impl[sizeof(g_impl) - 1] = 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace mimic

extern "C" {
int mimic_abi_version(void) { return MIMIC_ABI_VERSION; }
const char* mimic_last_error(void) { return mimic::g_err; }
const char* mimic_last_attn_impl(void) { return mimic::g_impl; }
}
