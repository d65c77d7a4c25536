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

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
}  // namespace dv3

extern "C" {
const char* dv3_last_error(void) { return dv3::g_err; }
int dv3_abi_version(void) { return 1; }
}
