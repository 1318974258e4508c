// orb_common.cu -- error text + version of liborbslam2_dualcam_b200.so
#include "orb_common.h"

#include <string.h>

namespace orbhost {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaSuccess) return ORB_OK;
    const char* base = strrchr(file, '/');
    set_error("%s:%d: %s -> %s (%s)", base ? base + 1 : file, line, what, cudaGetErrorName(e), cudaGetErrorString(e));
    return ORB_E_CUDA;
}

}  // namespace orbhost

extern "C" {

const char* orb_last_error(void) { return orbhost::g_err; }
const char* orb_version(void) { return "orbslam2_dualcam_b200 0.1 (sm_100a)"; }

}  // extern "C"
