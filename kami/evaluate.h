// kami::eval (kami/evaluate.h:6, kami/evaluate.cpp:10-160) -- arena gating of a candidate network, two drivers:
// eval() on the device-resident arena (kb_arena_*, all trees per kernel) and eval_hostlevel() through MCTS / NN objects.
// The second driver of the hot path: `evaluate_games` trees searched `evaluate_nodes` deep, each leaf
// evaluated by the network whose turn it is (two batched NN::infer calls per round, bootstrap off,
// evaluate.cpp:136-151), early pass / fail on the score bounds (evaluate.cpp:109-126).
//
// Behaviour kept from the reference: options evaluate_batch / evaluate_games / evaluate_nodes /
// evaluate_target_pct, rand()-drawn colours, pick(alpha = 0) moves, score = (1 + result * colour) / 2 per
// game, the colour a recycled tree gets (+1 when its last leaf went to the candidate's batch, else -1), the
// integer target `(egames * etarget) / 100`, the progress lines, and the "model was updated" bail-out.
// One deliberate difference: the reference sizes the colour table by evaluate_batch but indexes it by tree
// (evaluate.cpp:18-22, 68-80 -- out of bounds with the shipped 10 games > 8 batch); here every tree has a slot.
#pragma once
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <stdexcept>
#include <vector>

#include "env.h"
#include "mcts.h"
#include "nn/nn.h"
#include "options.h"

namespace kami {
// One finished arena game as both drivers report it (tests compare the two logs): tree, result (White's point of
// view), the colour the candidate played.
struct ArenaGameLog {
    int tree;
    float result;
    int colour;
};

// The colour table: the reference draws evaluate_batch values with rand() (evaluate.cpp:18-22) and then indexes the
// table by TREE (out of bounds for trees >= evaluate_batch: 10 games > 8 batch in options.def.yml).  The same draws
// for the first evaluate_batch trees (so the process-wide rand() stream stays aligned with the reference); trees beyond
// them, which the reference reads from whatever follows on its stack, alternate -1 / +1.
inline std::vector<int> arena_colours(int ebatch, int ntrees) {
    std::vector<int> colour((size_t)(ntrees > ebatch ? ntrees : ebatch));
    for (int i = 0; i < ebatch; ++i) colour[(size_t)i] = (rand() % 2) * 2 - 1;
    for (size_t i = (size_t)ebatch; i < colour.size(); ++i) colour[i] = (i % 2) ? 1 : -1;
    return colour;
}

// The arena through the reference-shaped public objects (one MCTS per game, NN::infer on host buffers): the restatement
// of evaluate.cpp:10-160 call for call.  kami::eval below runs the same algorithm on the device-resident arena
// (kb_arena_*); this one stays as its check (kami/tests/arena_parity.cpp: identical game logs and trees).
inline bool eval_hostlevel(NN* current_model, NN* candidate_model, int trainer, std::vector<ArenaGameLog>* log = nullptr) {
    const int ebatch = options::getInt("evaluate_batch");
    const int egames = options::getInt("evaluate_games");
    const int enodes = options::getInt("evaluate_nodes");
    const int etarget = options::getInt("evaluate_target_pct");
    const int ntrees = egames;

    std::vector<int> colour = arena_colours(ebatch, ntrees);  // side the candidate plays in each tree
    std::vector<MCTS> trees(ntrees);

    struct Side {  // one network's batch of pending leaves
        NN* net;
        std::vector<float> obs;
        std::vector<int> tree;
    };
    Side cur{current_model, std::vector<float>((size_t)ebatch * OBSIZE), {}};
    Side cd{candidate_model, std::vector<float>((size_t)ebatch * OBSIZE), {}};
    std::vector<float> policy((size_t)ebatch * PSIZE), value(ebatch), leaf(OBSIZE);

    float score = 0.0f;
    int games = 0;
    const float target_score = (float)((egames * etarget) / 100);
    std::cout << "EVAL " << trainer << ": evaluating model generation " << candidate_model->get_generation() << " over " << egames << " games"
              << std::endl;

    while (games < egames) {
        if (current_model->get_generation() >= candidate_model->get_generation()) {
            std::cout << "EVAL " << trainer << ": model was updated during evaluation, skipping!" << std::endl;
            return false;
        }
        cur.tree.clear();
        cd.tree.clear();
        for (int i = 0; i < ntrees; ++i) {
            if ((int)cur.tree.size() >= ebatch && (int)cd.tree.size() >= ebatch) break;
            MCTS& t = trees[i];
            // The reference picks the destination buffer AND its slot from the side to move BEFORE the descent
            // (evaluate.cpp:68-70), but files the leaf under the side to move AT the leaf (:80-90).  For a leaf at
            // odd depth the observation therefore lands in the other network's buffer and the network that is
            // asked evaluates whatever its slot held before.  Reproduced as is (arena results depend on it); only
            // the out-of-bounds write when that buffer is already full is dropped.
            const bool root_is_candidates = t.get_env().turn() == colour[i];
            Side* dst = root_is_candidates ? &cd : &cur;
            const size_t dst_slot = dst->tree.size();
            while (t.n() < enodes && !t.select(leaf.data())) {
            }
            if (t.n() < enodes) {  // a leaf waits for a network: whose turn is it there?
                const int turn = t.get_env().turn();
                Side* s = turn == colour[i] ? &cd : turn == -colour[i] ? &cur : nullptr;
                if (s && (int)s->tree.size() < ebatch) {
                    s->tree.push_back(i);
                    if ((int)dst_slot < ebatch) std::copy(leaf.begin(), leaf.end(), dst->obs.begin() + dst_slot * OBSIZE);
                }
                continue;
            }
            t.push(t.pick());
            float result;
            if (t.get_env().terminal(&result)) {
                if (log) log->push_back(ArenaGameLog{i, result, colour[i]});
                score += result * (float)colour[i] / 2.0f + 0.5f;
                ++games;
                std::cout << "EVAL " << trainer << ": game " << games << " of " << egames << " [" << result * colour[i] << "]: score "
                          << (int)(score * 100 / games) << "%" << std::endl;
                t.reset();
                colour[i] = root_is_candidates ? 1 : -1;
                if (score + (float)(egames - games) < target_score) {
                    std::cout << "EVAL " << trainer << ": aborting evaluation, score is too low" << std::endl;
                    return false;
                }
                if (score >= target_score && games < egames) {
                    std::cout << "EVAL " << trainer << ": finished evaluating early: score >=" << (int)(score * 100 / games) << "%, target " << etarget
                              << std::endl;
                    return true;
                }
            }
            --i;  // same tree again: next move, or the fresh game
        }
        for (Side* s : {&cur, &cd}) {
            if (s->tree.empty()) continue;
            s->net->infer(s->obs.data(), (int)s->tree.size(), policy.data(), value.data());
            for (size_t k = 0; k < s->tree.size(); ++k) trees[s->tree[k]].expand(policy.data() + k * PSIZE, value[k], true);
        }
    }
    std::cout << "EVAL " << trainer << ": finished evaluating: score " << (int)(score * 100 / games) << "%, target " << etarget << std::endl;
    return score * 100 / games >= (float)etarget;
}

// kami::eval on the device-resident arena (libkami_b200 kb_arena_*: all trees advance in one kernel per round, the two
// networks evaluate their batches on the device, leaves never leave HBM).  The host keeps what the reference keeps in
// its loop: the generation check (evaluate.cpp:54-60), the score, the progress lines and the early pass / fail rule
// (evaluate.cpp:100-126), applied to the finished games in the reference's order.
inline bool eval(NN* current_model, NN* candidate_model, int trainer, std::vector<ArenaGameLog>* log = nullptr, kb_arena** keep = nullptr) {
    const int ebatch = options::getInt("evaluate_batch");
    const int egames = options::getInt("evaluate_games");
    const int enodes = options::getInt("evaluate_nodes");
    const int etarget = options::getInt("evaluate_target_pct");
    std::vector<int> colour = arena_colours(ebatch, egames);

    kb_tree_cfg cfg;
    kb_check(kb_tree_default_cfg(&cfg));
    cfg.cpuct = options::getFloat("cpuct", 1.0f);
    cfg.force_expand_unvisited = options::getInt("force_expand_unvisited", 0);
    cfg.unvisited_node_value_pct = options::getInt("unvisited_node_value_pct", 100);
    cfg.bootstrap_weight = options::getInt("bootstrap_weight", 0);  // (unused: the arena expands with disable_bootstrap)
    cfg.bootstrap_window = options::getInt("bootstrap_window", 1600);
    cfg.bootstrap_amp_pct = options::getInt("bootstrap_amp_pct", 75);
    cfg.scale_cpuct_by_actions = options::getInt("scale_cpuct_by_actions", 0);
    cfg.noise_weight = options::getFloat("mcts_noise_weight", 0.05f);
    cfg.seed = (uint64_t)time(NULL);  // mcts.h:99
    kb_arena* arena = nullptr;
    kb_check(kb_arena_create(&arena, egames, ebatch, enodes, &cfg, colour.data(), (int)colour.size()));
    struct Guard {
        kb_arena* a;
        kb_arena** keep;
        ~Guard() {
            if (keep) *keep = a;  // (tests look at the trees afterwards and destroy the arena themselves)
            else kb_arena_destroy(a);
        }
    } guard{arena, keep};

    float score = 0.0f;
    int games = 0;
    const float target_score = (float)((egames * etarget) / 100);
    std::vector<kb_arena_game> finished((size_t)egames);
    std::cout << "EVAL " << trainer << ": evaluating model generation " << candidate_model->get_generation() << " over " << egames << " games"
              << std::endl;
    while (games < egames) {
        if (current_model->get_generation() >= candidate_model->get_generation()) {
            std::cout << "EVAL " << trainer << ": model was updated during evaluation, skipping!" << std::endl;
            return false;
        }
        int n = 0;
        {
            // both networks stay put for the round (NN::infer's shared lock, nn.cpp:166)
            int rc = candidate_model->arena_round(arena, current_model, finished.data(), (int)finished.size(), &n);
            if (rc == KB_ERR_NAN) throw std::runtime_error("inference policy output contains NaN");
            if (rc == KB_ERR_NO_CHILD) throw std::runtime_error("no child for action");
            kb_check(rc);
        }
        for (int k = 0; k < n; ++k) {
            const kb_arena_game& g = finished[(size_t)k];
            if (log) log->push_back(ArenaGameLog{g.tree, g.result, g.colour});
            score += g.result * (float)g.colour / 2.0f + 0.5f;
            ++games;
            std::cout << "EVAL " << trainer << ": game " << games << " of " << egames << " [" << g.result * g.colour << "]: score "
                      << (int)(score * 100 / games) << "%" << std::endl;
            if (score + (float)(egames - games) < target_score) {
                std::cout << "EVAL " << trainer << ": aborting evaluation, score is too low" << std::endl;
                return false;
            }
            if (score >= target_score && games < egames) {
                std::cout << "EVAL " << trainer << ": finished evaluating early: score >=" << (int)(score * 100 / games) << "%, target " << etarget
                          << std::endl;
                return true;
            }
            if (games >= egames) break;  // (games the reference's loop would not have reached this round)
        }
    }
    std::cout << "EVAL " << trainer << ": finished evaluating: score " << (int)(score * 100 / games) << "%, target " << etarget << std::endl;
    return score * 100 / games >= (float)etarget;
}
}  // namespace kami
