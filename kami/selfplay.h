// kami::Selfplay over the B200 C ABI -- the surface of the reference's kami/selfplay.h:24-102.
// Each inference thread owns a device-resident pool of `selfplay_batch` trees and runs the whole
// loop of Selfplay::inference_main (selfplay.cpp:113-200) on the GPU with kb_pool_step; finished
// games are drained into the host ReplayBuffer.  Training threads (train + arena) are
// SURVEY.md 8(f) items and are not started.
#pragma once
#include <atomic>
#include <chrono>
#include <iostream>
#include <list>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "env.h"
#include "nn/nn.h"
#include "options.h"
#include "replaybuffer.h"

namespace kami {
class Selfplay {
   public:
    Selfplay(NN* model)
        : model(model), replay_buffer(OBSIZE, PSIZE, options::getInt("replaybuffer_size", 512)), ibatch(options::getInt("selfplay_batch", 16)),
          nodes(options::getInt("selfplay_nodes", 512)), wants_pgn(false) {}

    void start() {
        status.code(RUNNING);
        int n_inference = options::getInt("inference_threads", 1);
        for (int i = 0; i < n_inference; ++i) {
            partial_trajectories.emplace_back(0);
            inference.push_back(std::thread(&Selfplay::inference_main, this, i));
        }
        if (options::getInt("training_threads", 1) > 0)
            std::cout << "TRAIN: training threads are not built in this round (SURVEY.md 8(f) #1/#2)" << std::endl;
    }
    void stop() {
        if (status.code() != RUNNING) throw std::runtime_error("stop() called when not running");
        status.code(WAITING);
        for (auto& t : inference) t.join();
        inference.clear();
        status.code(STOPPED);
    }

    enum StatusCode { STOPPED, RUNNING, WAITING };
    struct Status {
        StatusCode _code = STOPPED;
        std::mutex _lock;
        std::string _message;
        std::string message(std::string text = "") {
            std::lock_guard<std::mutex> lock(_lock);
            if (!text.size()) return text;
            return _message = text;
        }
        StatusCode code(int newcode = -1) {
            std::lock_guard<std::mutex> lock(_lock);
            if (newcode < 0) return _code;
            return _code = StatusCode(newcode);
        }
    };
    Status status;
    ReplayBuffer& get_rbuf() { return replay_buffer; }
    std::string get_next_pgn() { throw std::runtime_error("PGN export is not built (thc SAN printing is out of scope)"); }

   private:
    std::vector<std::thread> inference;
    NN* model;
    ReplayBuffer replay_buffer;
    int ibatch, nodes;
    std::atomic<bool> wants_pgn;
    std::list<std::atomic<int>> partial_trajectories;

    void inference_main(int id) {
        std::cout << "Starting inference thread: " << id << std::endl;
        kb_tree_cfg cfg;
        kb_check(kb_tree_default_cfg(&cfg));
        cfg.cpuct = options::getFloat("cpuct", 1.0f);
        cfg.force_expand_unvisited = options::getInt("force_expand_unvisited", 0);
        cfg.unvisited_node_value_pct = options::getInt("unvisited_node_value_pct", 100);
        cfg.bootstrap_weight = options::getInt("bootstrap_weight", 0);
        cfg.bootstrap_window = options::getInt("bootstrap_window", 1600);
        cfg.bootstrap_amp_pct = options::getInt("bootstrap_amp_pct", 75);
        cfg.scale_cpuct_by_actions = options::getInt("scale_cpuct_by_actions", 0);
        cfg.noise_weight = options::getFloat("mcts_noise_weight", 0.05f);
        cfg.seed = (uint64_t)time(NULL) * 1315423911u + (uint64_t)id;
        cfg.selfplay_nodes = nodes;
        cfg.alpha_initial = options::getFloat("selfplay_alpha_initial", 1.0f);
        cfg.alpha_decay = options::getFloat("selfplay_alpha_decay", 1.0f);
        cfg.alpha_final = options::getFloat("selfplay_alpha_final", 1.0f);
        cfg.alpha_cutoff = (int)options::getFloat("selfplay_alpha_cutoff", 1.0f);
        cfg.draw_value_pct = options::getInt("draw_value_pct", 50);
        kb_pool* pool = nullptr;
        kb_check(kb_pool_create(&pool, ibatch, options::getInt("b200_node_capacity", 1 << 18), &cfg));
        std::vector<float> obs((size_t)64 * OBSIZE), pi((size_t)64 * PSIZE), z(64);
        while (status.code() == RUNNING) {
            kb_check(kb_pool_step(pool, model->handle(), 64));
            int m = 0;
            do {
                kb_check(kb_pool_drain_samples(pool, 64, obs.data(), pi.data(), z.data(), &m));
                for (int i = 0; i < m; ++i) replay_buffer.add(&obs[(size_t)i * OBSIZE], &pi[(size_t)i * PSIZE], z[i]);
            } while (m == 64);
        }
        kb_pool_destroy(pool);
        std::cout << "Terminating inference thread: " << id << std::endl;
    }
};
}  // namespace kami
