// Error plumbing of the C ABI: every entry point returns 0 on success, non-zero otherwise; the
// message of the most recent failure on the calling thread is available through dv3_last_error().
#include "common.cuh"
#include <stdarg.h>

namespace dv3 {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
unsigned long long launch_count() { return g_launches; }

int check_launch(const char* what) {
    __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
}  // namespace dv3

extern "C" {
const char* dv3_last_error(void) { return dv3::g_err; }
int dv3_abi_version(void) { return 1; }
long long dv3_launch_count(void) { return (long long)dv3::launch_count(); }
}
