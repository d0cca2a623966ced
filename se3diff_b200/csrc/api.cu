// Error reporting, launch accounting and ABI version of libse3diff_b200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace se3 {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: %s", what, cudaGetErrorString(e));
        return SE3_ECUDA;
    }
    return SE3_OK;
}

void count_launch(int n) { g_launches += n; }

}  // namespace se3

extern "C" {
const char* se3_last_error(void) { return se3::g_err; }
int se3_abi_version(void) { return SE3_ABI_VERSION; }
int64_t se3_launch_count(void) { return se3::g_launches; }
void se3_launch_count_reset(void) { se3::g_launches = 0; }
}
