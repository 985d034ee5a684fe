// Shared host-side plumbing for libkami_b200: error reporting and CUDA call checking.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kami_b200.h"

namespace kb {

void set_error(const char* fmt, ...);
cudaStream_t main_stream();
int sm_count();
bool initialized();

#define KB_CUDA(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            kb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return KB_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

#define KB_REQUIRE_INIT()                                                    \
    do {                                                                     \
        if (!kb::initialized()) {                                            \
            int _r = kb_init(-1);                                            \
            if (_r != KB_OK) return _r;                                      \
        }                                                                    \
    } while (0)

#define KB_ARG(cond, msg)                     \
    do {                                      \
        if (!(cond)) {                        \
            kb::set_error("bad argument: %s", msg); \
            return KB_ERR_ARG;                \
        }                                     \
    } while (0)

}  // namespace kb
