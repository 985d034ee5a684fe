// Library-level plumbing of libkami_b200: device binding, error strings, raw memory helpers.
//
// Threading model (reference: kami/selfplay.cpp:25-31 runs `inference_threads` + `training_threads` host threads over one
// NN, nn.cpp:164-168): every host thread owns ONE CUDA stream per device (thread-local, created on first use), so the
// enqueues of different host threads never interleave on a stream and their kernels may overlap on the GPU.  Objects
// (kb_pool, kb_env, kb_trainer) are single-thread-owned like the reference's MCTS / Env; when an object moves to
// another thread the next call waits for the stream that last touched it.  kb_net is shared: its weights are guarded
// by a shared / exclusive lock inside the library (net.cu) and it holds no per-call state.
//
// Devices: kb_init(d) binds the CALLING THREAD to device d (first call per device uploads the tables).  Threads that
// never called kb_init adopt the first device the process initialised.  Every object remembers its device and its
// entry points switch the calling thread to it, so one process can drive all GPUs of a box, one host thread per GPU
// ("one host thread + CUDA stream set per GPU").
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include <cuda_profiler_api.h>

#include "common.cuh"

namespace kb {

constexpr int MAX_DEVICES = 16;
struct DeviceState {
    bool init = false;
    int sms = 0;
    char name[256] = "";
};
static DeviceState g_dev[MAX_DEVICES];
static std::mutex g_mu;
static int g_default_device = -1;
// streams are never destroyed: a thread that exits hands its streams back for the next thread, so a handle stored in
// an object (last_stream) always stays valid
static std::vector<cudaStream_t> g_free_streams[MAX_DEVICES];

struct ThreadCtx {
    int device = -1;
    cudaStream_t st[MAX_DEVICES] = {};
    cudaEvent_t ev[MAX_DEVICES][2] = {};
    ~ThreadCtx() {
        std::lock_guard<std::mutex> g(g_mu);
        for (int d = 0; d < MAX_DEVICES; ++d)
            if (st[d]) g_free_streams[d].push_back(st[d]);
    }
};
static thread_local ThreadCtx tl;
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
// the calling thread's stream on its current device
cudaStream_t main_stream() {
    const int d = tl.device < 0 ? 0 : tl.device;
    if (!tl.st[d]) {
        std::lock_guard<std::mutex> g(g_mu);
        if (!g_free_streams[d].empty()) {
            tl.st[d] = g_free_streams[d].back();
            g_free_streams[d].pop_back();
        } else if (cudaStreamCreateWithFlags(&tl.st[d], cudaStreamNonBlocking) != cudaSuccess) {
            tl.st[d] = nullptr;  // legacy stream: still correct, only slower
            cudaGetLastError();
        }
    }
    return tl.st[d];
}
int pdl_mask() {
    static int m = -1;
    if (m < 0) {
        const char* e = getenv("KB_PDL_MASK");
        m = e ? atoi(e) & 15 : 3;  // select + tower; expand launched plainly (measured: PDL on expand costs 8-10 us per step)
    }
    return m;
}
int sm_count() { return tl.device >= 0 ? g_dev[tl.device].sms : 0; }
bool initialized() { return tl.device >= 0; }
int current_device() { return tl.device; }
int upload_tables();  // tree.cu

// switch the calling thread to an (already initialised) device
int bind_device(int device) {
    if (tl.device == device) return KB_OK;
    if (device < 0 || device >= MAX_DEVICES || !g_dev[device].init) {
        set_error("object belongs to device %d, which this process has not initialised", device);
        return KB_ERR_STATE;
    }
    KB_CUDA(cudaSetDevice(device));
    tl.device = device;
    return KB_OK;
}
// an object last used on another thread's stream: wait for that stream before this thread touches it
void adopt(cudaStream_t* last) {
    cudaStream_t cur = main_stream();
    if (*last && *last != cur) {
        if (cudaStreamSynchronize(*last) != cudaSuccess) cudaGetLastError();
    }
    *last = cur;
}

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

// Binds the calling thread to `device` (device < 0: the process's first device, or LOCAL_RANK's when launched by
// torchrun, else 0).  The first call per device checks the architecture and uploads the tables.
int kb_init(int device) {
    if (device < 0 && tl.device >= 0) return KB_OK;
    if (device >= 0 && device == tl.device) return KB_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s): libkami_b200 has no CPU path", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        cudaGetLastError();
        return KB_ERR_CUDA;
    }
    if (device < 0) {
        if (g_default_device >= 0) device = g_default_device;
        else {
            const char* lr = getenv("LOCAL_RANK");
            device = lr ? atoi(lr) % n : 0;
        }
    }
    KB_ARG(device < n && device < MAX_DEVICES, "device ordinal out of range");
    KB_CUDA(cudaSetDevice(device));
    {
        std::lock_guard<std::mutex> g(g_mu);
        DeviceState& D = g_dev[device];
        if (!D.init) {
            cudaDeviceProp prop;
            KB_CUDA(cudaGetDeviceProperties(&prop, device));
            if (prop.major != 10) {
                set_error("device %d (%s) is sm_%d%d; libkami_b200 is built for sm_100a only", device, prop.name, prop.major, prop.minor);
                return KB_ERR_UNSUPPORTED;
            }
            D.sms = prop.multiProcessorCount;
            strncpy(D.name, prop.name, sizeof(D.name) - 1);
            int r = upload_tables();
            if (r) return r;
            D.init = true;
            if (g_default_device < 0) g_default_device = device;
        }
    }
    tl.device = device;
    return KB_OK;
}
int kb_current_device(void) { return tl.device; }

int kb_device_name(char* out, int cap) {
    KB_REQUIRE_INIT();
    KB_ARG(out && cap > 0, "out/cap");
    strncpy(out, g_dev[tl.device].name, cap - 1);
    out[cap - 1] = 0;
    return KB_OK;
}
int kb_sm_count(void) {
    if (tl.device < 0 && kb_init(-1) != KB_OK) return 0;
    return g_dev[tl.device].sms;
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
// pins a caller-owned buffer so that the host-pointer entry points (kb_net_infer, kb_pool_step_hostio, ...) copy
// straight from / into it instead of staging; unregister before freeing the buffer
int kb_host_register(void* ptr, size_t bytes) {
    KB_REQUIRE_INIT();
    KB_ARG(ptr && bytes, "ptr/bytes");
    KB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return KB_OK;
}
int kb_host_unregister(void* ptr) {
    if (ptr) KB_CUDA(cudaHostUnregister(ptr));
    return KB_OK;
}
int kb_host_free_pinned(void* ptr) {
    if (ptr) KB_CUDA(cudaFreeHost(ptr));
    return KB_OK;
}

// timing helpers for bench.py: CUDA events on the library's own stream
// (events live with the calling thread, like its stream: kb_timer_* brackets what THIS thread enqueued)
int kb_timer_start(void) {
    KB_REQUIRE_INIT();
    cudaEvent_t* g_ev = tl.ev[tl.device];
    if (!g_ev[0]) {
        KB_CUDA(cudaEventCreate(&g_ev[0]));
        KB_CUDA(cudaEventCreate(&g_ev[1]));
    }
    KB_CUDA(cudaEventRecord(g_ev[0], main_stream()));
    return KB_OK;
}
int kb_timer_stop(float* ms) {
    KB_REQUIRE_INIT();
    cudaEvent_t* g_ev = tl.ev[tl.device];
    KB_ARG(ms && g_ev[0], "ms / timer not started");
    KB_CUDA(cudaEventRecord(g_ev[1], main_stream()));
    KB_CUDA(cudaEventSynchronize(g_ev[1]));
    KB_CUDA(cudaEventElapsedTime(ms, g_ev[0], g_ev[1]));
    return KB_OK;
}
// capture window for `ncu --profile-from-start off` / nsys (tools/prof_step.py)
int kb_profiler_start(void) {
    KB_CUDA(cudaProfilerStart());
    return KB_OK;
}
int kb_profiler_stop(void) {
    KB_CUDA(cudaProfilerStop());
    return KB_OK;
}
// writes `bytes` of a scratch buffer on the library stream (L2 flush between timed iterations)
static void* g_flush[MAX_DEVICES] = {};
static size_t g_flush_bytes[MAX_DEVICES] = {};
int kb_flush_l2(size_t bytes) {
    KB_REQUIRE_INIT();
    std::lock_guard<std::mutex> g(g_mu);
    const int d = tl.device;
    if (bytes > g_flush_bytes[d]) {
        if (g_flush[d]) cudaFree(g_flush[d]);
        KB_CUDA(cudaMalloc(&g_flush[d], bytes));
        g_flush_bytes[d] = bytes;
    }
    KB_CUDA(cudaMemsetAsync(g_flush[d], 0, bytes, main_stream()));
    return KB_OK;
}

}  // extern "C"
