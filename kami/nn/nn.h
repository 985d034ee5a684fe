// kami::NN over the B200 C ABI -- the surface of the reference's kami/nn/nn.h:40-73 without
// LibTorch.  Forward runs as the tcgen05 kernels of libkami_b200; weights are a flat fp32 blob
// in the reference's parameter naming (oracle/nn_oracle.py:param_order).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <condition_variable>
#include <iostream>
#include <mutex>
#include <random>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kami_b200.h"
#include "../options.h"

namespace kami {
// NN::mut of the reference is a std::shared_mutex (nn.h:47).  Here the inference threads hold it shared for a whole
// device-resident step at a time, i.e. almost always, and glibc's rwlock prefers readers: read() / train() could wait
// forever.  Same interface (usable with std::shared_lock / std::unique_lock), but a waiting writer stops new readers.
class WriterFirstSharedMutex {
    std::mutex m;
    std::condition_variable cv;
    int readers = 0, writers_waiting = 0;
    bool writing = false;

   public:
    void lock_shared() {
        std::unique_lock<std::mutex> g(m);
        cv.wait(g, [&] { return !writing && writers_waiting == 0; });
        ++readers;
    }
    void unlock_shared() {
        std::unique_lock<std::mutex> g(m);
        if (--readers == 0) cv.notify_all();
    }
    void lock() {
        std::unique_lock<std::mutex> g(m);
        ++writers_waiting;
        cv.wait(g, [&] { return !writing && readers == 0; });
        --writers_waiting;
        writing = true;
    }
    void unlock() {
        std::unique_lock<std::mutex> g(m);
        writing = false;
        cv.notify_all();
    }
};

class NN {
   private:
    kb_net* net = nullptr;                    // the primary replica (device `device0`)
    int device0 = 0;
    std::vector<std::pair<int, kb_net*>> replicas;  // weight replicas on other GPUs of the box (one per GPU)
    int width, height, features, psize;
    int filters, residuals;
    WriterFirstSharedMutex mut;
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
        device0 = kb_current_device();
        random_init(std::random_device{}());
        check(kb_net_load_blob(net, blob.data(), blob.size()));
    }
    NN(NN* other) : width(other->width), height(other->height), features(other->features), psize(other->psize) {
        std::shared_lock<WriterFirstSharedMutex> g(other->mut);
        filters = other->filters;
        residuals = other->residuals;
        generation = other->generation;
        blob = other->blob;
        check(kb_net_create(&net, filters, residuals));
        device0 = kb_current_device();
        check(kb_net_load_blob(net, blob.data(), blob.size()));
    }
    ~NN() {
        kb_net_destroy(net);
        for (auto& r : replicas) kb_net_destroy(r.second);
    }
    NN(const NN&) = delete;
    NN& operator=(const NN&) = delete;

    int get_generation() {
        std::shared_lock<WriterFirstSharedMutex> g(mut);
        return generation;
    }
    int get_device() { return 0; }
    bool isCUDA() { return true; }
    int obsize() const { return width * height * features; }
    int polsize() const { return psize; }
    kb_net* handle() { return net; }
    // Self-play shards by game over the GPUs of the box with no collective: every GPU holds a full weight
    // replica.  handle(d) creates the replica on device d on first use (the calling thread stays bound to d: that is
    // what an inference thread serving GPU d wants); read() / train() / load_blob() refresh every replica.
    kb_net* handle(int device) {
        {
            std::shared_lock<WriterFirstSharedMutex> g(mut);
            if (device < 0 || device == device0) return net;
            for (auto& r : replicas)
                if (r.first == device) return r.second;
        }
        std::unique_lock<WriterFirstSharedMutex> g(mut);
        for (auto& r : replicas)
            if (r.first == device) return r.second;
        check(kb_init(device));
        kb_net* rep = nullptr;
        check(kb_net_create(&rep, filters, residuals));
        check(kb_net_load_blob(rep, blob.data(), blob.size()));
        replicas.emplace_back(device, rep);
        return rep;
    }

   private:
    void refresh_replicas(const std::vector<float>& b) {  // caller holds `mut` exclusively
        for (auto& r : replicas) check(kb_net_load_blob(r.second, b.data(), b.size()));
    }

   public:

    // nn.cpp:155-187: host buffers in and out, NaN -> runtime_error, value[i] = vh.flat[i]
    void infer(float* input, int batch, float* policy, float* value) {
        std::shared_lock<WriterFirstSharedMutex> g(mut);
        int rc = kb_net_infer(net, input, batch, policy, value);
        if (rc == KB_ERR_NAN) throw std::runtime_error("inference policy output contains NaN");
        check(rc);
    }
    // The device-resident form of the inference loop's NN::infer call (selfplay.cpp:196): `iters` rounds of
    // select -> tower -> expand on a pool, under the same shared lock NN::infer takes (nn.cpp:166), so read() / train()
    // never swap the weights under a running step.
    void pool_step(kb_pool* pool, int iters, int device = -1) {
        kb_net* h = handle(device);
        std::shared_lock<WriterFirstSharedMutex> g(mut);
        int rc = kb_pool_step(pool, h, iters);
        if (rc == KB_ERR_NAN) throw std::runtime_error("inference policy output contains NaN");
        check(rc);
    }
    // One round of the device-resident arena (kb_arena_round) with this network as the candidate and `current` as the
    // incumbent, both under their shared locks like the two NN::infer calls of evaluate.cpp:136-151.
    int arena_round(kb_arena* arena, NN* current, kb_arena_game* finished, int cap, int* n) {
        std::shared_lock<WriterFirstSharedMutex> g1(mut);
        if (current == this) return kb_arena_round(arena, net, net, finished, cap, n);
        std::shared_lock<WriterFirstSharedMutex> g2(current->mut);
        return kb_arena_round(arena, current->net, net, finished, cap, n);
    }
    // nn.cpp:224-377: `epochs` passes of shuffled mini-batches of `training_batchsize`, plain SGD with
    // lr = training_mlr / 1000, BatchNorm in training mode, ++generation, back to eval mode.  Each mini-batch
    // is one kb_trainer_forward_backward + kb_trainer_apply_sgd (tcgen05 forward / dgrad / wgrad).  Like the
    // reference the batch arrays persist across mini-batches, so the last partial batch keeps the previous
    // batch's rows (nn.cpp:278-298); the batches themselves are copied to the device as on the reference's CUDA
    // path (its CPU path aliases one stack buffer for every stored batch, nn.cpp:300-316).
    void train(int trajectories, float* inputs, float* obs_p, float* obs_v, bool detect_anomaly = false) {
        std::unique_lock<WriterFirstSharedMutex> g(mut);
        const float lr = (float)options::getInt("training_mlr", 5) / 1000.0f;
        const int epochs = options::getInt("training_epochs", 8);
        const int tbatch = options::getInt("training_batchsize", 8);
        const size_t osz = (size_t)width * height * features;
        // b200_train_devices = N > 1: the mini-batch is split over the first N GPUs of the box (kb_dp_*: a trainer replica per
        // GPU, one NCCL all-reduce of the gradient bucket per mini-batch over NVLink, the same SGD step everywhere).  The
        // reference trains on one device; with N = 1 (default) so does this, row for row.
        int ndev = options::getInt("b200_train_devices", 1);
        if (ndev < 1 || tbatch % ndev != 0) ndev = 1;
        kb_trainer* tr = nullptr;
        kb_dp* dp = nullptr;
        if (ndev > 1) {
            std::vector<int> devs(ndev);
            for (int i = 0; i < ndev; ++i) devs[i] = i;
            check(kb_dp_create(&dp, devs.data(), ndev, filters, residuals, tbatch / ndev));
        } else {
            check(kb_trainer_create(&tr, filters, residuals, tbatch));
        }
        struct Guard {
            kb_trainer* t;
            kb_dp* d;
            ~Guard() {
                kb_trainer_destroy(t);
                kb_dp_destroy(d);
            }
        } guard{tr, dp};
        check(dp ? kb_dp_load_blob(dp, blob.data(), blob.size()) : kb_trainer_load_blob(tr, blob.data(), blob.size()));
        std::vector<int> picker(trajectories);
        for (int i = 0; i < trajectories; ++i) picker[i] = i;
        auto rng = std::default_random_engine{};
        std::vector<float> next_input(tbatch * osz, 0.0f), next_policy((size_t)tbatch * psize, 0.0f), next_value(tbatch, 0.0f);
        float firstloss = 0.0f, lastloss = 0.0f;
        for (int epoch = 0; epoch < epochs; ++epoch) {
            std::shuffle(picker.begin(), picker.end(), rng);
            float avgloss = 0.0f, epfirst = 0.0f, eplast = 0.0f;
            int nbatches = 0;
            for (size_t base = 0; base < picker.size();) {
                int i = 0;
                for (; i < tbatch && base + i < picker.size(); ++i) {
                    const size_t src = (size_t)picker[base + i];
                    std::copy(inputs + src * osz, inputs + (src + 1) * osz, next_input.begin() + i * osz);
                    std::copy(obs_p + src * psize, obs_p + (src + 1) * psize, next_policy.begin() + (size_t)i * psize);
                    next_value[i] = obs_v[src];
                }
                base += i;
                if (detect_anomaly)
                    for (float v : next_input)
                        if (v != v) throw std::runtime_error("training input contains NaN");
                float loss = 0.0f;
                if (dp) {
                    check(kb_dp_step(dp, next_input.data(), next_policy.data(), next_value.data(), tbatch / ndev, lr, 1.0f, &loss));
                } else {
                    check(kb_trainer_forward_backward(tr, next_input.data(), next_policy.data(), next_value.data(), tbatch, &loss));
                }
                if (detect_anomaly && loss != loss) throw std::runtime_error("forward output contains NaN");
                if (!dp) check(kb_trainer_apply_sgd(tr, lr, 1.0f));
                avgloss += loss;
                if (!nbatches) epfirst = loss;
                eplast = loss;
                ++nbatches;
            }
            avgloss /= (float)(nbatches ? nbatches : 1);
            std::cout << "Epoch " << epoch + 1 << "/" << epochs << ": loss " << epfirst << " => " << eplast << ", " << nbatches << " batches" << std::endl;
            if (!epoch) firstloss = avgloss;
            lastloss = avgloss;
        }
        ++generation;
        std::cout << "Generated model " << generation << ", average loss " << firstloss << " to " << lastloss << " over " << epochs << " epochs\n";
        std::vector<float> b(blob.size());
        check(dp ? kb_dp_export_blob(dp, 0, b.data(), b.size()) : kb_trainer_export_blob(tr, b.data(), b.size()));
        check(kb_net_load_blob(net, b.data(), b.size()));
        refresh_replicas(b);
        blob.swap(b);
    }
    // Checkpoint = "KB20" + filters + residuals + generation + fp32 blob.  (The reference writes a
    // torch archive, nn.cpp:189-222; archive interop is SURVEY.md 8(f) #3.)
    void write(std::string path) {
        std::shared_lock<WriterFirstSharedMutex> g(mut);
        std::ofstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("couldn't open " + path + " for writing");
        int32_t hdr[4] = {0x3032424B, filters, residuals, generation};
        f.write((const char*)hdr, sizeof(hdr));
        f.write((const char*)blob.data(), blob.size() * sizeof(float));
        std::cout << "Saved model to " << path << std::endl;
    }
    void read(std::string path) {
        std::unique_lock<WriterFirstSharedMutex> g(mut);
        std::ifstream f(path, std::ios::binary);
        int32_t hdr[4];
        if (!f || !f.read((char*)hdr, sizeof(hdr)) || hdr[0] != 0x3032424B || hdr[1] != filters || hdr[2] != residuals)
            throw std::runtime_error("couldn't read model " + path);
        std::vector<float> b(kb_net_blob_floats(filters, residuals));
        if (!f.read((char*)b.data(), b.size() * sizeof(float))) throw std::runtime_error("truncated model " + path);
        check(kb_net_load_blob(net, b.data(), b.size()));
        refresh_replicas(b);
        blob.swap(b);
        generation = hdr[3];
    }
    // explicit weight loading from a flat blob (the oracle's exchange format)
    void load_blob(const float* data, size_t n) {
        std::unique_lock<WriterFirstSharedMutex> g(mut);
        check(kb_net_load_blob(net, data, n));
        blob.assign(data, data + n);
        refresh_replicas(blob);
    }
};
}  // namespace kami
