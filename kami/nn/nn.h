// kami::NN over the B200 C ABI -- the surface of the reference's kami/nn/nn.h:40-73 without
// LibTorch.  Forward runs as the tcgen05 kernels of libkami_b200; weights are a flat fp32 blob
// in the reference's parameter naming (oracle/nn_oracle.py:param_order).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <random>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kami_b200.h"
#include "../options.h"

namespace kami {
class NN {
   private:
    kb_net* net = nullptr;
    int width, height, features, psize;
    int filters, residuals;
    std::shared_mutex mut;
    int generation = 0;
    std::vector<float> blob;

    static void check(int rc) {
        if (rc != KB_OK) throw std::runtime_error(std::string("kami_b200: ") + kb_last_error());
    }
    // LibTorch's default initialisation (nn.cpp:45-56 register_module defaults): kaiming-uniform
    // a=sqrt(5) == U(+-1/sqrt(fan_in)) for weights and biases; BatchNorm gamma 1, beta 0,
    // running mean 0, running var 1.
    void random_init(uint64_t seed) {
        std::mt19937_64 rng(seed);
        const int F = filters, R = residuals;
        blob.clear();
        auto uni = [&](size_t n, int fan_in) {
            std::uniform_real_distribution<float> d(-1.0f / std::sqrt((float)fan_in), 1.0f / std::sqrt((float)fan_in));
            for (size_t i = 0; i < n; ++i) blob.push_back(d(rng));
        };
        auto conv = [&](int o, int c, int k) {
            uni((size_t)o * c * k * k, c * k * k);
            uni(o, c * k * k);
        };
        auto bn = [&](int c) {
            blob.insert(blob.end(), c, 1.0f);
            blob.insert(blob.end(), c, 0.0f);
            blob.insert(blob.end(), c, 0.0f);
            blob.insert(blob.end(), c, 1.0f);
        };
        conv(F, features, 3);
        bn(F);
        for (int i = 0; i < R; ++i) {
            conv(F, F, 3);
            bn(F);
            conv(F, F, 3);
            bn(F);
        }
        conv(128, F, 1);
        bn(128);
        conv(73, 128, 1);
        conv(1, F, 1);
        bn(1);
        uni((size_t)256 * 64, 64);
        uni(256, 64);
    }

   public:
    NN(int width, int height, int features, int psize, bool force_cpu = false)
        : width(width), height(height), features(features), psize(psize) {
        if (force_cpu) throw std::runtime_error("kami_b200 has no CPU path (force_cpu requested)");
        filters = options::getInt("filters", 256);     // nn.cpp:42
        residuals = options::getInt("residuals", 2);   // nn.cpp:43
        check(kb_net_create(&net, filters, residuals));
        random_init(std::random_device{}());
        check(kb_net_load_blob(net, blob.data(), blob.size()));
    }
    NN(NN* other) : width(other->width), height(other->height), features(other->features), psize(other->psize) {
        std::shared_lock<std::shared_mutex> g(other->mut);
        filters = other->filters;
        residuals = other->residuals;
        generation = other->generation;
        blob = other->blob;
        check(kb_net_create(&net, filters, residuals));
        check(kb_net_load_blob(net, blob.data(), blob.size()));
    }
    ~NN() { kb_net_destroy(net); }
    NN(const NN&) = delete;
    NN& operator=(const NN&) = delete;

    int get_generation() {
        std::shared_lock<std::shared_mutex> g(mut);
        return generation;
    }
    int get_device() { return 0; }
    bool isCUDA() { return true; }
    int obsize() const { return width * height * features; }
    int polsize() const { return psize; }
    kb_net* handle() { return net; }

    // nn.cpp:155-187: host buffers in and out, NaN -> runtime_error, value[i] = vh.flat[i]
    void infer(float* input, int batch, float* policy, float* value) {
        std::shared_lock<std::shared_mutex> g(mut);
        int rc = kb_net_infer(net, input, batch, policy, value);
        if (rc == KB_ERR_NAN) throw std::runtime_error("inference policy output contains NaN");
        check(rc);
    }
    void train(int, float*, float*, float*, bool = false) {
        throw std::runtime_error("NN::train is not built yet (SURVEY.md 8(f) #1: train step + NCCL all-reduce)");
    }
    // Checkpoint = "KB20" + filters + residuals + generation + fp32 blob.  (The reference writes a
    // torch archive, nn.cpp:189-222; archive interop is SURVEY.md 8(f) #3.)
    void write(std::string path) {
        std::shared_lock<std::shared_mutex> g(mut);
        std::ofstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("couldn't open " + path + " for writing");
        int32_t hdr[4] = {0x3032424B, filters, residuals, generation};
        f.write((const char*)hdr, sizeof(hdr));
        f.write((const char*)blob.data(), blob.size() * sizeof(float));
        std::cout << "Saved model to " << path << std::endl;
    }
    void read(std::string path) {
        std::unique_lock<std::shared_mutex> g(mut);
        std::ifstream f(path, std::ios::binary);
        int32_t hdr[4];
        if (!f || !f.read((char*)hdr, sizeof(hdr)) || hdr[0] != 0x3032424B || hdr[1] != filters || hdr[2] != residuals)
            throw std::runtime_error("couldn't read model " + path);
        std::vector<float> b(kb_net_blob_floats(filters, residuals));
        if (!f.read((char*)b.data(), b.size() * sizeof(float))) throw std::runtime_error("truncated model " + path);
        check(kb_net_load_blob(net, b.data(), b.size()));
        blob.swap(b);
        generation = hdr[3];
    }
    // explicit weight loading from a flat blob (the oracle's exchange format)
    void load_blob(const float* data, size_t n) {
        std::unique_lock<std::shared_mutex> g(mut);
        check(kb_net_load_blob(net, data, n));
        blob.assign(data, data + n);
    }
};
}  // namespace kami
