// Internal interface of the network module (net.cu) used by the pool step (tree.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/kami_b200.h"

namespace kb {
// planes: bf16 tall-image input (layout.cuh) for `batch` boards; policy [batch][4672] fp32
// softmax over all logits; value256 [batch][256] fp32.  Asynchronous on `stream`.
int net_forward_async(kb_net* net, const void* planes, int batch, float* policy_dev, float* value256_dev, cudaStream_t stream);
// Pool-step form (north_star kernel 3: softmax over the LEGAL moves only).  Board b's legal action
// codes are the first n_b uint16 at act_base + b*stride, n_b is the int at nact_base + b*stride.
// Writes prior[b][i] = exp(logit[a_i] - max_j logit[a_j]) for i < n_b (row pitch 128 floats); the
// normalisation by the sum over legal moves is MCTS::expand's own (mcts.h:273-276, 296), so the
// result equals the dense softmax followed by that renormalisation.  No [batch][4672] tensor is
// written.  Returns KB_ERR_UNSUPPORTED when the net has no legal-gather path.
int net_forward_legal_async(kb_net* net, const void* planes, int batch, const void* act_base, const void* nact_base, size_t stride,
                            float* prior_dev, float* value256_dev, cudaStream_t stream);
// makes sure the net's activation workspace fits `batch` boards (may allocate)
int net_reserve(kb_net* net, int batch);
// input plane buffer owned by the net for `batch` boards (pad pixels already zero)
void* net_input_planes(kb_net* net);
int net_launches_per_forward(kb_net* net);
// Group form: the forward of `batch` boards whose planes were written at net_group_planes(net, item0), using the
// activation workspace from item `item0` on, so forwards of disjoint groups can run concurrently on different
// streams.  Reserve the workspace for all groups (net_reserve) before the first group is launched.
int net_forward_group_async(kb_net* net, int item0, int batch, float* policy_dev, float* value256_dev, cudaStream_t stream);
void* net_group_planes(kb_net* net, int item0);
// fp32 [n][64][30] observations -> bf16 tall-image planes (tree.cu)
int obs_to_tall_launch(const float* obs_dev, int n, void* planes, cudaStream_t st);
}  // namespace kb
