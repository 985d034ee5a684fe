// kami::eval (kami/evaluate.h:6, kami/evaluate.cpp:10-160) -- arena gating of a candidate network.
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
#include <iostream>
#include <vector>

#include "env.h"
#include "mcts.h"
#include "nn/nn.h"
#include "options.h"

namespace kami {
inline bool eval(NN* current_model, NN* candidate_model, int trainer) {
    const int ebatch = options::getInt("evaluate_batch");
    const int egames = options::getInt("evaluate_games");
    const int enodes = options::getInt("evaluate_nodes");
    const int etarget = options::getInt("evaluate_target_pct");
    const int ntrees = egames;

    std::vector<int> colour(ntrees > ebatch ? ntrees : ebatch);  // side the candidate plays in each tree
    for (auto& c : colour) c = (rand() % 2) * 2 - 1;
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
}  // namespace kami
