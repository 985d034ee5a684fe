// Internal interface of the network module (net.cu) used by the pool step (tree.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/kami_b200.h"

namespace kb {
// planes: bf16 tall-image input (layout.cuh) for `batch` boards; policy [batch][4672] fp32
// softmax over all logits; value256 [batch][256] fp32.  Asynchronous on `stream`.
int net_forward_async(kb_net* net, const void* planes, int batch, float* policy_dev, float* value256_dev, cudaStream_t stream);
// makes sure the net's activation workspace fits `batch` boards (may allocate)
int net_reserve(kb_net* net, int batch);
// input plane buffer owned by the net for `batch` boards (pad pixels already zero)
void* net_input_planes(kb_net* net);
int net_launches_per_forward(kb_net* net);
// fp32 [n][64][30] observations -> bf16 tall-image planes (tree.cu)
int obs_to_tall_launch(const float* obs_dev, int n, void* planes, cudaStream_t st);
}  // namespace kb
