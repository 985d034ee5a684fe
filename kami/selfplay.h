// kami::Selfplay over the B200 C ABI -- the surface of the reference's kami/selfplay.h:24-102.
// Each inference thread owns a device-resident pool of `selfplay_batch` trees and runs the whole
// loop of Selfplay::inference_main (selfplay.cpp:113-200) on the GPU with kb_pool_step; finished
// games are drained into the host ReplayBuffer.  Training threads follow Selfplay::training_main
// (selfplay.cpp:215-304): wait for the replay target, clone the model, ReplayBuffer::select_batch, NN::train
// (the CUDA training step), kami::eval (the arena), and on acceptance write + read the model file.
#pragma once
#include <algorithm>
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
#include "evaluate.h"
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
        int n_training = options::getInt("training_threads", 1);
        for (int i = 0; i < n_training; ++i) training.push_back(std::thread(&Selfplay::training_main, this, i));
    }
    void stop() {
        if (status.code() != RUNNING) throw std::runtime_error("stop() called when not running");
        status.code(WAITING);
        for (auto& t : inference) t.join();
        for (auto& t : training) t.join();
        inference.clear();
        training.clear();
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
    // extension (not in the reference): totals over the inference threads, for throughput reporting
    unsigned long long evals() const { return total_evals.load(); }
    unsigned long long moves() const { return total_moves.load(); }
    unsigned long long games() const { return total_games.load(); }
    // selfplay.h:73-80: the next game an inference thread finishes is written out as PGN movetext
    std::string get_next_pgn() {
        wants_pgn = true;
        while (wants_pgn && status.code() == RUNNING) std::this_thread::sleep_for(std::chrono::milliseconds(100));
        std::lock_guard<std::mutex> lock(pgn_lock);
        return ret_pgn;
    }

   private:
    std::vector<std::thread> inference, training;
    NN* model;
    ReplayBuffer replay_buffer;
    int ibatch, nodes;
    std::atomic<bool> wants_pgn;
    std::string ret_pgn;
    std::mutex pgn_lock;
    std::list<std::atomic<int>> partial_trajectories;
    std::atomic<unsigned long long> total_evals{0}, total_moves{0}, total_games{0};

    void inference_main(int id) {
        std::cout << "Starting inference thread: " << id << std::endl;
        // Games never interact (selfplay.cpp:97): with b200_devices = N the inference threads are spread over the first N
        // GPUs of the box (thread id -> GPU id mod N), each with its own weight replica, node pools and streams; no
        // collective on this path.  0 = every visible GPU.
        int ndev = options::getInt("b200_devices", 1);
        if (ndev <= 0) ndev = kb_device_count();
        const int dev = ndev > 1 ? id % ndev : -1;
        if (dev >= 0) kb_check(kb_init(dev));
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
        // a tree that absorbed this many terminal visits in one step sits the step out instead of holding up the whole
        // batch (include/kami_b200.h kb_pool_set_terminal_cap); 0 = the reference's batch, always one leaf per tree
        kb_check(kb_pool_set_terminal_cap(pool, options::getInt("b200_terminal_cap", 1)));
        const int DRAIN = 256;  // replay rows fetched per call (one device -> host copy per call)
        std::vector<float> obs((size_t)DRAIN * OBSIZE), pi((size_t)DRAIN * PSIZE), z(DRAIN);
        const bool flush_old_trees = options::getInt("flush_old_trees", 1) != 0;  // selfplay.cpp:61
        int source_generation = model->get_generation();                          // selfplay.cpp:103
        long long flushed = 0;
        kb_pool_stats seen = {};
        const int iters_per_call = options::getInt("b200_iters_per_call", 256);
        std::vector<int32_t> game_actions(4096);
        bool game_requested = false;
        auto partial = partial_trajectories.begin();
        std::advance(partial, id);
        while (status.code() == RUNNING) {
            // selfplay.cpp:119-131: trees searched with an older generation are replaced and their partial trajectories
            // dropped (all of a thread's trees share one source generation here: they are stepped together)
            const int gen = model->get_generation();
            if (flush_old_trees && source_generation < gen) {
                kb_pool_stats before;
                kb_check(kb_pool_get_stats(pool, &before));
                flushed = (long long)(before.moves - before.samples);  // every partial trajectory is dropped (:126-127)
                kb_check(kb_pool_flush_trees(pool));
                source_generation = gen;
            }
            model->pool_step(pool, iters_per_call, dev);
            kb_pool_stats st;
            kb_check(kb_pool_get_stats(pool, &st));
            total_evals += st.evals - seen.evals;
            total_moves += st.moves - seen.moves;
            total_games += st.games - seen.games;
            seen = st;
            *partial = (int)((long long)(st.moves - st.samples) - flushed);  // positions recorded in games still being played (selfplay.cpp:150-151)
            int m = 0;
            do {
                kb_check(kb_pool_drain_samples(pool, DRAIN, obs.data(), pi.data(), z.data(), &m));
                for (int i = 0; i < m; ++i) replay_buffer.add(&obs[(size_t)i * OBSIZE], &pi[(size_t)i * PSIZE], z[i]);
            } while (m == DRAIN);
            // selfplay.cpp:167-171: hand the next finished game to get_next_pgn().  The device keeps the moves of the
            // game; its movetext is written by replaying them on an Env.
            if (!game_requested && wants_pgn.load()) {
                kb_check(kb_pool_request_game(pool));
                game_requested = true;
            }
            if (game_requested) {
                int n = 0;
                kb_check(kb_pool_take_game(pool, game_actions.data(), (int)game_actions.size(), &n));
                if (n > 0) {
                    game_requested = false;
                    if (wants_pgn.load()) {
                        Env env;
                        for (int i = 0; i < n; ++i) env.push(game_actions[i]);
                        std::string text = env.pgn();
                        {
                            std::lock_guard<std::mutex> lock(pgn_lock);
                            ret_pgn = text;
                        }
                        wants_pgn = false;
                    }
                }
            }
        }
        kb_pool_destroy(pool);
        std::cout << "Terminating inference thread: " << id << std::endl;
    }

    // selfplay.cpp:215-304
    void training_main(int id) {
        std::cout << "TRAIN " << id << ": starting thread " << id << std::endl;
        const std::string modelpath = options::getStr("model_path", "/tmp/model.pt");
        const long cap = replay_buffer.size();
        long target = cap, from = 0;  // train when count() reaches `target`; progress is shown relative to `from`
        const int step = (int)(cap * options::getInt("rpb_train_pct", 40) / 100);
        const int trajectories = (int)(cap * options::getInt("training_sample_pct", 60) / 100);
        const bool detect_anomaly = options::getInt("training_detect_anomaly", 0) != 0;
        if (detect_anomaly && !id) std::cout << "Anomaly detection enabled" << std::endl;
        std::vector<float> inputs((size_t)trajectories * OBSIZE), visits((size_t)trajectories * PSIZE), results(trajectories);
        while (status.code() == RUNNING) {
            const long have = replay_buffer.count();
            if (have < target) {
                if (!id) {
                    std::cout << "Gen " << model->get_generation() << " RPB " << 100 * (have - from) / (target - from) << "% [" << have - from << " / "
                              << target - from << "] | Partials: ";
                    int k = 0;
                    for (auto& p : partial_trajectories) std::cout << " Inf " << k++ << ": " << p;
                    std::cout << std::endl;
                }
                std::this_thread::sleep_for(std::chrono::milliseconds(1000));
                continue;
            }
            std::cout << "TRAIN " << id << ": training generation " << model->get_generation() << " with " << trajectories
                      << " trajectories sampled from last " << cap << std::endl;
            NN candidate(model);
            replay_buffer.select_batch(inputs.data(), visits.data(), results.data(), trajectories);
            candidate.train(trajectories, inputs.data(), visits.data(), results.data(), detect_anomaly);
            bool accepted = false;
            try {
                accepted = eval(model, &candidate, id);
            } catch (std::exception& e) {
                std::cerr << "TRAIN " << id << ": evaluation failed: " << e.what() << std::endl;
            }
            if (accepted) {
                candidate.write(modelpath);
                model->read(modelpath);
                std::cout << "TRAIN " << id << ": candidate accepted: using new generation " << model->get_generation() << std::endl;
                if (options::getInt("flush_old_rpb", 1)) replay_buffer.clear();
                target = std::max(cap, replay_buffer.count() + (long)step);
                from = replay_buffer.count();
                continue;
            }
            std::cout << "TRAIN " << id << ": candidate rejected: generation remains " << model->get_generation() << std::endl;
            from = replay_buffer.count();
            target += step;
        }
        std::cout << "TRAIN " << id << ": stopping thread" << std::endl;
    }
};
}  // namespace kami
