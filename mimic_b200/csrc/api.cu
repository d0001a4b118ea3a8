// Library-wide state: error string, implementation tag, device query.
#include <cstring>

#include "attn_common.cuh"

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

static unsigned long long* g_trace = nullptr;
static size_t g_trace_cap = 0;
unsigned long long* debug_trace_buffer() { return g_trace; }
size_t debug_trace_capacity() { return g_trace_cap; }
void set_trace(void* buf, size_t n) { g_trace = static_cast<unsigned long long*>(buf); g_trace_cap = n; }

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

// L2 working-set budget for kernels that choose their work order by what stays cache-resident: a third of the device's
// L2 (both dies keep their own copy of lines that SMs of both dies read), queried once.
long long l2_budget_bytes() {
  static long long b = 0;
  if (b == 0) {
    int dev = 0, l2 = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev) != cudaSuccess || l2 <= 0)
      l2 = 126 << 20;
    b = (long long)l2 / 3;
  }
  return b;
}

}  // namespace mimic

extern "C" {
int mimic_abi_version(void) { return MIMIC_ABI_VERSION; }
const char* mimic_last_error(void) { return mimic::g_err; }
const char* mimic_last_attn_impl(void) { return mimic::g_impl; }
void mimic_debug_set_trace(void* device_u64_buffer, size_t n_entries) { mimic::set_trace(device_u64_buffer, n_entries); }
void mimic_debug_force_fwd_kernel(int kernel) { mimic::tc::force_fwd_kernel(kernel); }
}
