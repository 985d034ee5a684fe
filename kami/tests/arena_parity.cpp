// kami::eval (device-resident arena, kb_arena_*) against kami::eval_hostlevel (the reference-shaped driver: one MCTS
// per game, NN::infer on host buffers) with real networks: identical game logs and verdicts after the same srand(),
// noise off.  Shapes include more games than batch rows (leaves wait a round) and games >= 2 x batch (both batches
// full: the rest of the trees are not touched that round).  Built by `make -C kami dropin`.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include "../evaluate.h"

using namespace kami;

static bool run_case(int batch, int games, int nodes, int target, unsigned seed, NN* cur, NN* cd) {
    options::setInt("evaluate_batch", batch);
    options::setInt("evaluate_games", games);
    options::setInt("evaluate_nodes", nodes);
    options::setInt("evaluate_target_pct", target);
    std::vector<ArenaGameLog> a, b;
    std::stringstream sink;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    srand(seed);
    const bool va = eval_hostlevel(cur, cd, 0, &a);
    const int ra = rand();
    srand(seed);
    const bool vb = eval(cur, cd, 0, &b);
    const int rb = rand();
    std::cout.rdbuf(old);
    bool same = va == vb && a.size() == b.size() && ra == rb;  // (same verdict, same games, same number of rand() draws)
    for (size_t i = 0; same && i < a.size(); ++i) same = a[i].tree == b[i].tree && a[i].result == b[i].result && a[i].colour == b[i].colour;
    printf("%s batch %d games %d nodes %d target %d: %zu / %zu games, verdict %d / %d\n", same ? "PARITY OK" : "PARITY FAIL", batch, games, nodes, target,
           a.size(), b.size(), (int)va, (int)vb);
    if (!same) {
        for (size_t i = 0; i < a.size() || i < b.size(); ++i) {
            if (i < a.size()) printf("  host   %zu: tree %d result %g colour %d\n", i, a[i].tree, a[i].result, a[i].colour);
            if (i < b.size()) printf("  device %zu: tree %d result %g colour %d\n", i, b[i].tree, b[i].result, b[i].colour);
        }
    }
    return same;
}

int main() {
    options::setInt("filters", 64);
    options::setInt("residuals", 1);
    options::setFloat("cpuct", 1.5f);
    options::setInt("unvisited_node_value_pct", 50);
    options::setInt("bootstrap_weight", 20);
    options::setFloat("mcts_noise_weight", 0.0f);
    NN current(8, 8, NFEATURES, PSIZE);
    NN candidate(8, 8, NFEATURES, PSIZE);  // different random weights
    {   // the candidate must be a newer generation (evaluate.cpp:54-60)
        float in[8 * 8 * NFEATURES] = {0}, pi[PSIZE] = {0}, z[8] = {0};
        options::setInt("training_epochs", 1);
        options::setInt("training_batchsize", 1);
        pi[0] = 1.0f;
        std::stringstream sink;
        std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
        candidate.train(1, in, pi, z);
        std::cout.rdbuf(old);
    }
    bool ok = true;
    ok &= run_case(8, 6, 12, 60, 3, &current, &candidate);    // fewer games than rows
    ok &= run_case(8, 10, 10, 60, 4, &current, &candidate);   // options.def.yml shape: 10 games, 8 rows
    ok &= run_case(2, 7, 8, 70, 5, &current, &candidate);     // games >= 2 x batch: both batches fill up
    ok &= run_case(3, 3, 24, 40, 6, &current, &candidate);
    return ok ? 0 : 1;
}
