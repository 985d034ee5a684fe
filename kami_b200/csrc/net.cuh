// Internal interface of the network module (net.cu) used by the pool step (tree.cu).
//
// Threading contract (reference: NN::infer runs concurrently from `inference_threads` + arena threads under a shared
// lock, NN::read / NN::train take it exclusively, nn.cpp:164-168, 206, 226): a kb_net holds WEIGHTS ONLY.  Every
// caller brings its own activation workspace (NetWs: input planes, activations, logits scratch, NaN flag) and its
// own stream, so forwards of different host threads never share a buffer.  The weights are guarded by a
// shared / exclusive lock inside kb_net: every entry point that reads them holds it shared until its last kernel
// has completed, kb_net_load_blob holds it exclusively (and synchronises the device before it frees anything).
#pragma once
#include <cuda_runtime.h>
#include "../../include/kami_b200.h"
#include "layout.cuh"

namespace kb {

// Activation workspace of one caller (a tree pool, or one host-pointer inference call).  Grown by ws_reserve on the
// owner's stream only; never shared between streams that may run concurrently (disjoint item ranges of one
// workspace may: the groups of kb_pool_step / kb_pool_step_hostio).
struct NetWs {
    int cap_boards = 0;
    int slabs = 0;       // 64-channel slabs per item of X / Y (0: fused tower, X / Y / H not needed)
    uint4 *P = nullptr, *X = nullptr, *Y = nullptr, *H = nullptr;
    float* logits = nullptr;  // [cap][4672] scratch of the legal-gather mode (per-layer path)
    int logits_cap = 0;
    int* nan_flag = nullptr;
};
// makes sure the workspace fits `batch` boards of `net` (may synchronise `st`, free and allocate)
int ws_reserve(NetWs& ws, kb_net* net, int batch, cudaStream_t st);
void ws_free(NetWs& ws);
inline uint4* ws_planes(NetWs& ws, int item0 = 0) { return ws.P + (size_t)item0 * IN_SLABS * SLAB_U4; }  // input planes from item0 on

// shared lock on the weights for the duration of a call that enqueues forwards AND waits for them
void net_lock_shared(kb_net* net);
void net_unlock_shared(kb_net* net);
struct NetReadGuard {
    kb_net* net;
    explicit NetReadGuard(kb_net* n) : net(n) { net_lock_shared(net); }
    ~NetReadGuard() { net_unlock_shared(net); }
    NetReadGuard(const NetReadGuard&) = delete;
    NetReadGuard& operator=(const NetReadGuard&) = delete;
};
int net_device(kb_net* net);

// planes: bf16 tall-image input (layout.cuh) for `batch` boards; policy [batch][4672] fp32
// softmax over all logits; value256 [batch][256] fp32.  Asynchronous on `stream`.
// item0: first item of the activation workspace this forward may use, so that forwards of disjoint groups of boards
// can run concurrently on different streams.
int net_forward_async(kb_net* net, NetWs& ws, const void* planes, int batch, float* policy_dev, float* value256_dev, cudaStream_t stream,
                      int item0 = 0);
// Pool-step form (north_star kernel 3: softmax over the LEGAL moves only).  Board b's legal action
// codes are the first n_b uint16 at act_base + b*stride, n_b is the int at nact_base + b*stride.
// Writes prior[b][i] = exp(logit[a_i] - max_j logit[a_j]) for i < n_b (row pitch 128 floats); the
// normalisation by the sum over legal moves is MCTS::expand's own (mcts.h:273-276, 296), so the
// result equals the dense softmax followed by that renormalisation.  No [batch][4672] tensor is
// written.
int net_forward_legal_async(kb_net* net, NetWs& ws, const void* planes, int batch, const void* act_base, const void* nact_base, size_t stride,
                            float* prior_dev, float* value256_dev, cudaStream_t stream, int item0 = 0, const unsigned* sync = nullptr,
                            unsigned sync_target = 0);
// true when `net` runs as the single fused tower kernel (64 filters): only that kernel understands the split select's counters
bool net_is_fused(kb_net* net);
int net_launches_per_forward(kb_net* net);
// fp32 [n][64][30] observations -> bf16 tall-image planes (tree.cu)
int obs_to_tall_launch(const float* obs_dev, int n, void* planes, cudaStream_t st);
}  // namespace kb
