// Data-parallel NN::train mini-batches from ONE host process (BASELINE config 5): a kb_trainer
// replica per GPU, each differentiates its own rows of the batch, ONE ncclAllReduce(sum) of the flat fp32 gradient bucket
// over NVLink, the same plain-SGD step on every replica (nn.cpp:239-241 has no optimiser state, so the replicas stay
// bit-identical).  The reference trains on a single device; this is the collective the hot path needs and the only one.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library loads and runs self-play on boxes without it, and a
// process that already holds torch's bundled NCCL (bench.py under torchrun) shares that copy.  The calling host thread
// drives every GPU: libkami_b200 gives it one stream per device (lib.cu), so the replicas' kernels run concurrently.
#include <dlfcn.h>

#include <new>
#include <vector>

#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclFloat = 7, kNcclSum = 0, kNcclAvg = 4;  // ncclFloat32, ncclSum, ncclAvg (nccl.h)
struct Nccl {
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
Nccl* nccl() {
    static Nccl n;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            n.CommInitAll = (decltype(n.CommInitAll))dlsym(h, "ncclCommInitAll");
            n.CommDestroy = (decltype(n.CommDestroy))dlsym(h, "ncclCommDestroy");
            n.AllReduce = (decltype(n.AllReduce))dlsym(h, "ncclAllReduce");
            n.GroupStart = (decltype(n.GroupStart))dlsym(h, "ncclGroupStart");
            n.GroupEnd = (decltype(n.GroupEnd))dlsym(h, "ncclGroupEnd");
            n.GetErrorString = (decltype(n.GetErrorString))dlsym(h, "ncclGetErrorString");
            n.ok = n.CommInitAll && n.CommDestroy && n.AllReduce && n.GroupStart && n.GroupEnd;
        }
    }
    return &n;
}
#define KB_NCCL(expr)                                                                                              \
    do {                                                                                                           \
        ncclResult_t _r = (expr);                                                                                  \
        if (_r != 0) {                                                                                             \
            kb::set_error("%s failed: %s", #expr, nccl()->GetErrorString ? nccl()->GetErrorString(_r) : "NCCL error"); \
            return KB_ERR_CUDA;                                                                                    \
        }                                                                                                          \
    } while (0)

}  // namespace

struct kb_dp {
    std::vector<int> devices;
    std::vector<kb_trainer*> tr;
    std::vector<ncclComm_t> comm;
    std::vector<float*> grads, params;
    std::vector<size_t> stat_off, stat_cnt;  // BatchNorm running statistics inside the parameter vector (same for every replica)
    size_t n_floats = 0;
    int home = 0;  // device the calling thread was bound to at creation (restored after every call)
};

using namespace kb;

extern "C" {

int kb_dp_create(kb_dp** out, const int* devices, int n, int filters, int residuals, int max_batch_per_device) {
    KB_REQUIRE_INIT();
    KB_ARG(out && devices && n >= 1 && n <= 16, "out / devices / 1 <= n <= 16");
    if (!nccl()->ok) {
        set_error("libnccl.so.2 not found (or too old): data-parallel training needs NCCL");
        return KB_ERR_UNSUPPORTED;
    }
    kb_dp* d = new (std::nothrow) kb_dp();
    if (!d) return KB_ERR_ARG;
    d->home = current_device();
    d->devices.assign(devices, devices + n);
    int r = KB_OK;
    for (int i = 0; i < n && r == KB_OK; ++i) {
        if ((r = kb_init(devices[i]))) break;
        kb_trainer* t = nullptr;
        if ((r = kb_trainer_create(&t, filters, residuals, max_batch_per_device))) break;
        d->tr.push_back(t);
        void* g = nullptr;
        size_t nf = 0;
        if ((r = kb_trainer_grad_buffer(t, &g, &nf))) break;
        d->grads.push_back((float*)g);
        d->n_floats = nf;
        if ((r = kb_trainer_param_buffer(t, &g, &nf))) break;
        d->params.push_back((float*)g);
        if (i == 0) {
            d->stat_off.resize(256);
            d->stat_cnt.resize(256);
            int ns = 0;
            if ((r = kb_trainer_stat_ranges(t, d->stat_off.data(), d->stat_cnt.data(), 256, &ns))) break;
            d->stat_off.resize((size_t)ns);
            d->stat_cnt.resize((size_t)ns);
        }
    }
    if (r == KB_OK) {
        d->comm.resize((size_t)n);
        ncclResult_t nr = nccl()->CommInitAll(d->comm.data(), n, devices);
        if (nr != 0) {
            set_error("ncclCommInitAll failed: %s", nccl()->GetErrorString ? nccl()->GetErrorString(nr) : "NCCL error");
            d->comm.clear();
            r = KB_ERR_CUDA;
        }
    }
    kb_init(d->home);
    if (r != KB_OK) {
        for (kb_trainer* t : d->tr) kb_trainer_destroy(t);
        delete d;
        return r;
    }
    *out = d;
    return KB_OK;
}
int kb_dp_destroy(kb_dp* d) {
    if (!d) return KB_OK;
    for (kb_trainer* t : d->tr) kb_trainer_destroy(t);
    for (ncclComm_t c : d->comm) nccl()->CommDestroy(c);
    kb_init(d->home);
    delete d;
    return KB_OK;
}
int kb_dp_size(kb_dp* d) { return d ? (int)d->tr.size() : KB_ERR_ARG; }
kb_trainer* kb_dp_replica(kb_dp* d, int rank) { return d && rank >= 0 && rank < (int)d->tr.size() ? d->tr[(size_t)rank] : nullptr; }

int kb_dp_load_blob(kb_dp* d, const float* blob, size_t n_floats) {
    KB_ARG(d && blob, "dp / blob");
    int r = KB_OK;
    for (kb_trainer* t : d->tr)
        if ((r = kb_trainer_load_blob(t, blob, n_floats))) break;
    kb_init(d->home);
    return r;
}
int kb_dp_export_blob(kb_dp* d, int rank, float* blob, size_t n_floats) {
    KB_ARG(d && blob && rank >= 0 && rank < (int)d->tr.size(), "dp / blob / rank");
    int r = kb_trainer_export_blob(d->tr[(size_t)rank], blob, n_floats);
    kb_init(d->home);
    return r;
}

// One data-parallel mini-batch.  Rank r takes rows [r * batch_per_device, (r + 1) * batch_per_device) of the host arrays
// (the layout of NN::train's inputs, nn.h:67).  loss (optional) = the sum of the replicas' losses (NNModule::loss sums
// the policy term over the batch, nn.cpp:96-102).  lr as in nn.cpp:239: the all-reduced gradient is the SUM over all
// rows, exactly what one device would have accumulated for the policy term; grad_scale lets the caller average instead.
int kb_dp_step(kb_dp* d, const float* obs, const float* obs_p, const float* obs_v, int batch_per_device, float lr, float grad_scale, float* loss) {
    KB_ARG(d && obs && obs_p && obs_v && batch_per_device >= 1, "dp / arrays / batch");
    const int n = (int)d->tr.size();
    int r = KB_OK;
    std::vector<float> losses((size_t)n, 0.0f);
    // forward + backward on every GPU (the host-array form synchronises per replica: fine for the reference-shaped call;
    // hosts with resident batches use kb_dp_step_dev)
    for (int i = 0; i < n && r == KB_OK; ++i)
        r = kb_trainer_forward_backward(d->tr[(size_t)i], obs + (size_t)i * batch_per_device * KB_OBSIZE, obs_p + (size_t)i * batch_per_device * KB_PSIZE,
                                        obs_v + (size_t)i * batch_per_device, batch_per_device, &losses[(size_t)i]);
    if (r == KB_OK) r = kb_dp_allreduce_apply(d, lr, grad_scale);
    if (loss) {
        *loss = 0.0f;
        for (float l : losses) *loss += l;
    }
    kb_init(d->home);
    return r;
}
// The same with every replica's batch already resident on its GPU: nothing synchronises until the final wait.
int kb_dp_step_dev(kb_dp* d, const float* const* obs_dev, const float* const* obs_p_dev, const float* const* obs_v_dev, int batch_per_device, float lr,
                   float grad_scale) {
    KB_ARG(d && obs_dev && obs_p_dev && obs_v_dev && batch_per_device >= 1, "dp / arrays / batch");
    const int n = (int)d->tr.size();
    int r = KB_OK;
    for (int i = 0; i < n && r == KB_OK; ++i)
        r = kb_trainer_forward_backward_dev(d->tr[(size_t)i], obs_dev[i], obs_p_dev[i], obs_v_dev[i], batch_per_device, nullptr);
    if (r == KB_OK) r = kb_dp_allreduce_apply(d, lr, grad_scale);
    kb_init(d->home);
    return r;
}
// gradients of all replicas -> their sum on every replica (one NCCL all-reduce of the whole bucket, plus the few hundred
// floats of BatchNorm running statistics averaged in the same NCCL group), then SGD everywhere; returns when every GPU
// has finished
int kb_dp_allreduce_apply(kb_dp* d, float lr, float grad_scale) {
    KB_ARG(d, "dp");
    const int n = (int)d->tr.size();
    std::vector<cudaStream_t> st((size_t)n);
    for (int i = 0; i < n; ++i) {
        int r = kb_init(d->devices[(size_t)i]);
        if (r) return r;
        st[(size_t)i] = main_stream();  // the stream the replica's backward was enqueued on: the collective is ordered behind it
    }
    if (n > 1) {
        KB_NCCL(nccl()->GroupStart());
        for (int i = 0; i < n; ++i) {
            ncclResult_t nr = nccl()->AllReduce(d->grads[(size_t)i], d->grads[(size_t)i], d->n_floats, kNcclFloat, kNcclSum, d->comm[(size_t)i], st[(size_t)i]);
            if (nr != 0) {
                nccl()->GroupEnd();
                set_error("ncclAllReduce failed: %s", nccl()->GetErrorString ? nccl()->GetErrorString(nr) : "NCCL error");
                return KB_ERR_CUDA;
            }
        }
        // BatchNorm running statistics were updated from each replica's own rows: average them, so that the replicas stay
        // bit-identical and any of them can be exported (torch's DistributedDataParallel broadcasts rank 0's instead)
        for (size_t k = 0; k < d->stat_off.size(); ++k)
            for (int i = 0; i < n; ++i) {
                float* ptr = d->params[(size_t)i] + d->stat_off[k];
                ncclResult_t nr = nccl()->AllReduce(ptr, ptr, d->stat_cnt[k], kNcclFloat, kNcclAvg, d->comm[(size_t)i], st[(size_t)i]);
                if (nr != 0) {
                    nccl()->GroupEnd();
                    set_error("ncclAllReduce (statistics) failed: %s", nccl()->GetErrorString ? nccl()->GetErrorString(nr) : "NCCL error");
                    return KB_ERR_CUDA;
                }
            }
        KB_NCCL(nccl()->GroupEnd());
    }
    for (int i = 0; i < n; ++i) {
        int r = kb_trainer_apply_sgd(d->tr[(size_t)i], lr, grad_scale);
        if (r) return r;
    }
    for (int i = 0; i < n; ++i) {
        int r = kb_init(d->devices[(size_t)i]);
        if (r) return r;
        KB_CUDA(cudaStreamSynchronize(st[(size_t)i]));
    }
    return kb_init(d->home);
}

}  // extern "C"
