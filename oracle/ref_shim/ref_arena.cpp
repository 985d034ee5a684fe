// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// The UNMODIFIED reference arena (kami/evaluate.cpp, compiled from /root/reference where it lies, with the reference's
// own Env / MCTS / options) driven with INJECTED network outputs: this file supplies the member functions of kami::NN
// that evaluate.cpp calls (the class declaration is the reference's kami/nn/nn.h; the reference's nn.cpp, i.e. the
// LibTorch network, is simply not linked).  NN::infer returns a deterministic pseudo-network -- an integer hash of the
// observation -- that tests/test_gpu_arena.py restates in numpy, so the reference's arena control flow (batch routing,
// colours, scoring, early stop) can be compared game for game with the device-resident arena fed the same outputs.
#include <cstdint>
#include <cstdlib>
#include <malloc.h>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

#include <atomic>
#include <cstring>
#include <shared_mutex>
#include <vector>

#include <torch/torch.h>

#define private public  // the shim defines NN's members and sets `generation`
#include "kami/nn/nn.h"
#undef private
#include "kami/evaluate.h"
#include "kami/env.h"
#include "kami/options.h"

namespace {
std::map<kami::NN*, uint32_t> g_salt;
}

namespace kami {
NN::NN(int width, int height, int features, int psize, bool) : width(width), height(height), features(features), psize(psize) { generation = 0; }
NN::NN(NN* other) : width(other->width), height(other->height), features(other->features), psize(other->psize) { generation = other->generation; }

// pseudo-network: s = sum_j (j + 1) * (int) obs[j] (mod 2^32);
//   policy[a] = ((s * 2654435761 + a * 40503 + salt) >> 8) % 997 + 1        (positive, unnormalised: expand renormalises)
//   value     = (((s * 40503 + salt * 7919) >> 4) % 2001 - 1000) / 1000
void NN::infer(float* input, int batch, float* policy, float* value) {
    const uint32_t salt = g_salt[this];
    const int osz = width * height * features;
    for (int b = 0; b < batch; ++b) {
        uint32_t s = 0;
        for (int j = 0; j < osz; ++j) s += (uint32_t)(j + 1) * (uint32_t)(int)input[(size_t)b * osz + j];
        for (int a = 0; a < psize; ++a)
            policy[(size_t)b * psize + a] = (float)(((s * 2654435761u + (uint32_t)a * 40503u + salt) >> 8) % 997u + 1u);
        value[b] = (float)((int)(((s * 40503u + salt * 7919u) >> 4) % 2001u) - 1000) / 1000.0f;
    }
}
}  // namespace kami

extern "C" {

// Runs kami::eval (evaluate.cpp:10-160) after srand(seed) with two pseudo-networks (salts) and returns its stdout log
// followed by "VERDICT 0|1"; options are set through ref_opt_set_* of the core shim's options object (same TU set here).
// the pseudo-network on its own (tests pin their numpy restatement of it against this)
void ref_arena_pseudo_infer(uint32_t salt, float* obs, int batch, float* policy, float* value) {
    kami::NN nn(8, 8, kami::NFEATURES, kami::PSIZE);
    g_salt[&nn] = salt;
    nn.infer(obs, batch, policy, value);
    g_salt.erase(&nn);
}
void ref_arena_opt_int(const char* key, int v) { kami::options::setInt(key, v); }
void ref_arena_opt_float(const char* key, float v) { kami::options::setFloat(key, v); }
int ref_arena_run(unsigned seed, uint32_t salt_current, uint32_t salt_candidate, char* out, int cap) {
    kami::NN current(8, 8, kami::NFEATURES, kami::PSIZE), candidate(&current);
    g_salt[&current] = salt_current;
    g_salt[&candidate] = salt_candidate;
    candidate.mut.lock();
    candidate.generation = 1;  // eval() bails out unless the candidate is newer (evaluate.cpp:54-60)
    candidate.mut.unlock();
    std::stringstream log;
    std::streambuf* old = std::cout.rdbuf(log.rdbuf());
    srand(seed);
    // evaluate.cpp:25-28 allocates its two input buffers with new[] and never clears them, and (see kami/evaluate.h in this
    // repo) a network can be asked about a slot nothing was written to: the reference then evaluates heap garbage.  glibc's
    // M_PERTURB = 0xFF makes malloc hand out zero-filled memory, which pins that case to "the buffers start as zeros".
    mallopt(M_PERTURB, 0xFF);
    bool verdict = false;
    std::string err;
    try {
        verdict = kami::eval(&current, &candidate, 0);
    } catch (std::exception& e) {
        err = e.what();
    }
    mallopt(M_PERTURB, 0);
    std::cout.rdbuf(old);
    g_salt.erase(&current);
    g_salt.erase(&candidate);
    std::string text = log.str() + (err.empty() ? "" : "EXCEPTION " + err + "\n") + "VERDICT " + (verdict ? "1" : "0") + "\n";
    if ((int)text.size() + 1 > cap) return -1;
    memcpy(out, text.c_str(), text.size() + 1);
    return (int)text.size();
}

}  // extern "C"
