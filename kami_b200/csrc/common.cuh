// Shared host-side plumbing for libkami_b200: error reporting and CUDA call checking.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kami_b200.h"

namespace kb {

void set_error(const char* fmt, ...);
cudaStream_t main_stream();  // the CALLING THREAD's stream on its current device (lib.cu: threading model)
int sm_count();
bool initialized();
int current_device();
int bind_device(int device);      // switch the calling thread to an initialised device
void adopt(cudaStream_t* last);   // object hand-over between host threads: wait for the stream that last touched it

#define KB_CUDA(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            kb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return KB_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl may become resident while the kernel before
// it on the stream is still running, run its own set-up (barrier init, TMEM allocation, zero fills, constant loads)
// and must execute pdl_wait() before it touches anything the earlier kernel writes or reads.  Every kernel of such a
// chain calls pdl_wait() unconditionally, so "my predecessor completed" also means all earlier kernels completed.
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int pdl_mask();  // KB_PDL_MASK: bit 0 select, bit 1 tower, bit 2 expand, bit 3 fused expand+select launched with the PDL attribute

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int which, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_mask() >> which) & 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#define KB_REQUIRE_INIT()                                                    \
    do {                                                                     \
        if (!kb::initialized()) {                                            \
            int _r = kb_init(-1);                                            \
            if (_r != KB_OK) return _r;                                      \
        }                                                                    \
    } while (0)

// entry points that take an object: run on the object's device, ordered after whatever last touched it
#define KB_BIND(obj)                                   \
    do {                                               \
        int _b = kb::bind_device((obj)->device);       \
        if (_b != KB_OK) return _b;                    \
        kb::adopt(&(obj)->last_stream);                \
    } while (0)

#define KB_ARG(cond, msg)                     \
    do {                                      \
        if (!(cond)) {                        \
            kb::set_error("bad argument: %s", msg); \
            return KB_ERR_ARG;                \
        }                                     \
    } while (0)

}  // namespace kb
