// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C shim over the UNMODIFIED reference network (kami/nn/nn.cpp, compiled from
// /root/reference where it lies) linked against the LibTorch inside the installed torch
// wheel.  Used to (a) pin oracle/nn_oracle.py, (b) generate tests/golden NN fixtures and
// (c) time the reference's CPU NN::infer in bench.py's reference arm.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include <torch/torch.h>

#define private public
#include "kami/nn/nn.h"
#undef private
#include "kami/options.h"

using namespace kami;

namespace {
struct Handle {
    NN* nn;
    std::vector<std::pair<std::string, torch::Tensor>> tensors;  // params then buffers
};
void collect(Handle* h) {
    h->tensors.clear();
    for (auto& kv : h->nn->mod->named_parameters()) h->tensors.push_back({kv.key(), kv.value()});
    for (auto& kv : h->nn->mod->named_buffers()) h->tensors.push_back({kv.key(), kv.value()});
}
}  // namespace

extern "C" {

void ref_nn_opt_set_int(const char* key, int v) { options::setInt(key, v); }
void ref_nn_set_threads(int n) { torch::set_num_threads(n); }
int ref_nn_get_threads() { return torch::get_num_threads(); }

// NN(8, 8, 30, 4672, force_cpu) after options filters/residuals and torch::manual_seed(seed)
// (kami/nn/nn.cpp:36-57, 107-128).
void* ref_nn_new(int filters, int residuals, uint64_t seed, int force_cpu) {
    options::setInt("filters", filters);
    options::setInt("residuals", residuals);
    torch::manual_seed(seed);
    Handle* h = new Handle();
    h->nn = new NN(8, 8, 30, 73 * 64, force_cpu != 0);
    collect(h);
    return h;
}
void ref_nn_free(void* p) {
    Handle* h = (Handle*)p;
    delete h->nn;
    delete h;
}
int ref_nn_is_cuda(void* p) { return ((Handle*)p)->nn->isCUDA() ? 1 : 0; }

// kami/nn/nn.cpp:155-187 (including the value memcpy quirk).
int ref_nn_infer(void* p, float* input, int batch, float* policy, float* value) {
    try {
        ((Handle*)p)->nn->infer(input, batch, policy, value);
    } catch (std::exception& e) {
        return -1;
    }
    return 0;
}

// Full [batch,256] value-head output + [batch,4672] policy, bypassing infer()'s memcpy so
// every FC output can be compared (forward = kami/nn/nn.cpp:59-91).
int ref_nn_forward_full(void* p, float* input, int batch, float* policy, float* value256) {
    Handle* h = (Handle*)p;
    torch::NoGradGuard guard;
    torch::Tensor in = torch::from_blob(input, {batch, 8, 8, 30}, torch::kCPU).to(h->nn->device, torch::kFloat32);
    std::vector<torch::Tensor> out = h->nn->mod->forward(in);
    torch::Tensor ph = out[0].cpu().contiguous(), vh = out[1].cpu().contiguous();
    memcpy(policy, ph.data_ptr<float>(), sizeof(float) * batch * 4672);
    memcpy(value256, vh.data_ptr<float>(), sizeof(float) * batch * 256);
    return 0;
}

// NN::train (kami/nn/nn.cpp:224-377) with the reference's own option keys: training_mlr (lr*1000),
// training_epochs, training_batchsize.  Leaves the module in eval mode with generation + 1.
int ref_nn_train(void* p, int n, float* inputs, float* obs_p, float* obs_v, int mlr, int epochs, int batchsize) {
    options::setInt("training_mlr", mlr);
    options::setInt("training_epochs", epochs);
    options::setInt("training_batchsize", batchsize);
    Handle* h = (Handle*)p;
    // NN::train prints progress with operator<<(int/float); inside a Python process whose libstdc++
    // locale facets come from another copy of the library that crashes, so mute the stream meanwhile
    std::cout.setstate(std::ios_base::failbit);
    try {
        h->nn->train(n, inputs, obs_p, obs_v, false);
    } catch (std::exception& e) {
        std::cout.clear();
        return -1;
    }
    std::cout.clear();
    collect(h);
    return h->nn->get_generation();
}

// NN::write / NN::read (kami/nn/nn.cpp:189-222): the reference's torch-archive checkpoint incl. "generation"
int ref_nn_write(void* p, const char* path) {
    std::cout.setstate(std::ios_base::failbit);
    try {
        ((Handle*)p)->nn->write(path);
    } catch (std::exception& e) {
        std::cout.clear();
        return -1;
    }
    std::cout.clear();
    return 0;
}
int ref_nn_read(void* p, const char* path) {
    Handle* h = (Handle*)p;
    try {
        h->nn->read(path);
    } catch (...) {
        return -1;
    }
    collect(h);
    return 0;
}
int ref_nn_generation(void* p) { return ((Handle*)p)->nn->get_generation(); }

int ref_nn_num_tensors(void* p) { return (int)((Handle*)p)->tensors.size(); }
// Returns numel; writes name and up to 4 dims (rank returned through *rank).
long ref_nn_tensor_info(void* p, int i, char* name, int cap, long* dims, int* rank, int* is_int64) {
    auto& t = ((Handle*)p)->tensors[i];
    strncpy(name, t.first.c_str(), cap - 1);
    name[cap - 1] = 0;
    *rank = (int)t.second.dim();
    for (int d = 0; d < t.second.dim() && d < 4; ++d) dims[d] = t.second.size(d);
    *is_int64 = t.second.scalar_type() == torch::kInt64;
    return t.second.numel();
}
void ref_nn_tensor_get(void* p, int i, float* out) {
    torch::Tensor t = ((Handle*)p)->tensors[i].second.detach().cpu().to(torch::kFloat32).contiguous();
    memcpy(out, t.data_ptr<float>(), sizeof(float) * t.numel());
}
void ref_nn_tensor_set(void* p, int i, const float* in) {
    torch::NoGradGuard guard;
    torch::Tensor& t = ((Handle*)p)->tensors[i].second;
    torch::Tensor src = torch::from_blob((void*)in, t.sizes(), torch::kFloat32).clone();
    t.copy_(src.to(t.device(), t.scalar_type()));
}

}  // extern "C"
