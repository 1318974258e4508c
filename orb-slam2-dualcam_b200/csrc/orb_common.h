// orb_common.h -- shared host-side plumbing of the C-ABI library: thread-local error text, CUDA call checking.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/orbslam2_dualcam_b200.h"

namespace orbhost {

void set_error(const char* fmt, ...);

// returns ORB_OK or ORB_E_CUDA (and records file:line + the CUDA error string)
int check_cuda(cudaError_t e, const char* what, const char* file, int line);

}  // namespace orbhost

#define ORB_CUDA(call)                                                              \
    do {                                                                            \
        int _rc = orbhost::check_cuda((call), #call, __FILE__, __LINE__);           \
        if (_rc != ORB_OK) return _rc;                                              \
    } while (0)

#define ORB_FAIL(code, ...)              \
    do {                                 \
        orbhost::set_error(__VA_ARGS__); \
        return (code);                   \
    } while (0)
