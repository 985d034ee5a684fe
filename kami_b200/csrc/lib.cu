// Library-level plumbing of libkami_b200: device binding, error strings, raw memory helpers.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace kb {

static thread_local char g_err[512] = "";
static bool g_init = false;
static int g_device = -1;
static int g_sms = 0;
static cudaStream_t g_stream = nullptr;
static char g_name[256] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
cudaStream_t main_stream() { return g_stream; }
int pdl_mask() {
    static int m = -1;
    if (m < 0) {
        const char* e = getenv("KB_PDL_MASK");
        m = e ? atoi(e) & 15 : 3;  // select + tower; expand launched plainly (measured: PDL on expand costs 8-10 us per step)
    }
    return m;
}
int sm_count() { return g_sms; }
bool initialized() { return g_init; }
int upload_tables();  // tree.cu

}  // namespace kb

using namespace kb;

extern "C" {

const char* kb_last_error(void) { return g_err; }

int kb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// device < 0: keep the current device (or LOCAL_RANK's when launched by torchrun).
int kb_init(int device) {
    if (g_init && (device < 0 || device == g_device)) return KB_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s): libkami_b200 has no CPU path", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        cudaGetLastError();
        return KB_ERR_CUDA;
    }
    if (device < 0) {
        const char* lr = getenv("LOCAL_RANK");
        device = lr ? atoi(lr) % n : 0;
    }
    KB_ARG(device < n, "device ordinal out of range");
    KB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    KB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d (%s) is sm_%d%d; libkami_b200 is built for sm_100a only", device, prop.name, prop.major, prop.minor);
        return KB_ERR_UNSUPPORTED;
    }
    g_sms = prop.multiProcessorCount;
    strncpy(g_name, prop.name, sizeof(g_name) - 1);
    if (!g_stream) KB_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    g_device = device;
    int r = upload_tables();
    if (r) return r;
    g_init = true;
    return KB_OK;
}

int kb_device_name(char* out, int cap) {
    KB_REQUIRE_INIT();
    KB_ARG(out && cap > 0, "out/cap");
    strncpy(out, g_name, cap - 1);
    out[cap - 1] = 0;
    return KB_OK;
}
int kb_sm_count(void) {
    if (!g_init && kb_init(-1) != KB_OK) return 0;
    return g_sms;
}

int kb_dev_alloc(void** out, size_t bytes) {
    KB_REQUIRE_INIT();
    KB_ARG(out, "out");
    KB_CUDA(cudaMalloc(out, bytes ? bytes : 16));
    KB_CUDA(cudaMemsetAsync(*out, 0, bytes ? bytes : 16, main_stream()));
    return KB_OK;
}
int kb_dev_free(void* ptr) {
    if (ptr) KB_CUDA(cudaFree(ptr));
    return KB_OK;
}
int kb_dev_upload(void* dst_dev, const void* src_host, size_t bytes) {
    KB_REQUIRE_INIT();
    KB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_dev_download(void* dst_host, const void* src_dev, size_t bytes) {
    KB_REQUIRE_INIT();
    KB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_dev_sync(void) {
    KB_REQUIRE_INIT();
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_host_alloc_pinned(void** out, size_t bytes) {
    KB_REQUIRE_INIT();
    KB_ARG(out, "out");
    KB_CUDA(cudaMallocHost(out, bytes ? bytes : 16));
    return KB_OK;
}
int kb_host_free_pinned(void* ptr) {
    if (ptr) KB_CUDA(cudaFreeHost(ptr));
    return KB_OK;
}

// timing helpers for bench.py: CUDA events on the library's own stream
static cudaEvent_t g_ev[2] = {nullptr, nullptr};
int kb_timer_start(void) {
    KB_REQUIRE_INIT();
    if (!g_ev[0]) {
        KB_CUDA(cudaEventCreate(&g_ev[0]));
        KB_CUDA(cudaEventCreate(&g_ev[1]));
    }
    KB_CUDA(cudaEventRecord(g_ev[0], main_stream()));
    return KB_OK;
}
int kb_timer_stop(float* ms) {
    KB_REQUIRE_INIT();
    KB_ARG(ms && g_ev[0], "ms / timer not started");
    KB_CUDA(cudaEventRecord(g_ev[1], main_stream()));
    KB_CUDA(cudaEventSynchronize(g_ev[1]));
    KB_CUDA(cudaEventElapsedTime(ms, g_ev[0], g_ev[1]));
    return KB_OK;
}
// writes `bytes` of a scratch buffer on the library stream (L2 flush between timed iterations)
static void* g_flush = nullptr;
static size_t g_flush_bytes = 0;
int kb_flush_l2(size_t bytes) {
    KB_REQUIRE_INIT();
    if (bytes > g_flush_bytes) {
        if (g_flush) cudaFree(g_flush);
        KB_CUDA(cudaMalloc(&g_flush, bytes));
        g_flush_bytes = bytes;
    }
    KB_CUDA(cudaMemsetAsync(g_flush, 0, bytes, main_stream()));
    return KB_OK;
}

}  // extern "C"
