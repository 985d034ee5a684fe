// Device-resident MCTS: node pools, warp-per-tree PUCT select / expand / backup, the
// warp-cooperative legal-action pipeline and the plane encoders.
//
// Reference: kami/mcts.h:15-349 (Node, MCTS), kami/env.h:398-423 (Env::actions),
// kami/selfplay.cpp:113-200 (the loop the batched kernels reproduce).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

#include "chess.cuh"

namespace kb {

constexpr int HIST_GAME = 64;   // game keys kept per tree; a repetition can only lie within hmc < 50 plies (Q5)
constexpr int MAX_DEPTH = 447;  // selection path capacity (nodes below the root)
constexpr int HIST_CAP = HIST_GAME + MAX_DEPTH + 1;
constexpr int WARPS_PER_BLOCK = 2;
constexpr int TRAJ_MAX_CHILD = 128;

struct __align__(16) Node {
    int n;       // visits            (mcts.h:16)
    float w;     // accumulated value (mcts.h:17)
    float p;     // prior             (mcts.h:18)
    u32 child0;  // index of the first child inside this tree's current space (children are contiguous)
};
// meta word per node: action in the low 16 bits, number of children in the high 16
__device__ __forceinline__ u32 meta_pack(int action, int nchild) { return (u32)action | ((u32)nchild << 16); }

struct Cfg {
    float cpuct;
    int force_expand_unvisited;
    float fpu;  // unvisited_node_value
    float bootstrap_weight, bootstrap_window, bootstrap_amp;
    int scale_cpuct_by_actions;
    float noise_weight;
    u64 seed;
    int selfplay_nodes;
    float alpha_initial, alpha_decay, alpha_final;
    int alpha_cutoff;
    float draw_value;
    int value_index_mode;
};

// One finished-or-running game's training sample in compact form (selfplay.cpp:141-160):
// the root position, the sparse visit distribution and the point of view.
struct TrajSample {
    Pos pos;
    float pov;
    int root_n;
    int nchild;
    int action;  // the move played from this position (selfplay.cpp:157-161)
    u32 entry[TRAJ_MAX_CHILD];  // action << 16 | visits
};
struct ReplaySample {
    TrajSample s;
    float z;
    int pad[3];
};

struct TreeCtl {
    Pos root_pos;
    Pos leaf_pos;
    u32 root;    // node index of the root in the current space
    u32 alloc;   // nodes used in the current space
    u32 space;   // active semi-space
    int state;   // 0 idle, 1 leaf waiting for expand(), 3 leaf chosen but its move list is still to be generated (split select), 2 sat this step out (kb_pool_set_terminal_cap): the next expand is a no-op
    int depth;   // nodes below the root on the current path
    int n_hist;  // game keys in hist[]; path keys follow
    int leaf_nact;
    int traj_len;
    u32 want_compact;  // the arena is nearly full: k_pool_compact moves the live subtree at the next safe point
    u32 pad_;
    u64 rng;
    u64 games, moves;
    u16 leaf_act[MAX_MOVES];
    u32 path[MAX_DEPTH + 1];
    u64 hist[HIST_CAP];
};

struct Stats {
    unsigned long long evals, moves, games, terminal_visits, children_scanned, path_nodes, children_created, samples, skipped;
};

struct PoolDev {
    int n_trees;
    int tree0, tree_hi;  // the range of trees a batched launch serves (whole pool: 0, n_trees); kb_pool_step_hostio runs
                         // groups of trees as independent pipelines on their own streams
    u32 cap;  // nodes per semi-space per tree
    Node* nodes;
    u32* meta;
    TreeCtl* ctl;
    int* error;
    Stats* stats;
    TrajSample* traj;
    int traj_cap;
    ReplaySample* replay;
    int replay_cap;
    unsigned long long* replay_head;
    int* game;  // [0] state: 0 idle, 1 requested, 2 being written, 3 ready; [1] length; [2..] actions of one finished
                // game (Selfplay::get_next_pgn, selfplay.h:73-80 / selfplay.cpp:167-171)
    Cfg cfg;
    long long* dbg;     // optional [n_trees][8] cycle counters of the last k_pool_select (kb_pool_debug_select_profile)
    int defer_compact;  // batched loops: push only flags a full arena, k_pool_compact (a block per tree) copies
    // split select (kb_pool_step, cap mode): a warp publishes its leaf's planes BEFORE it generates the leaf's moves, so the
    // tower (launched with programmatic dependent launch, already resident) starts on the planes while the move
    // generators are still running.  sync[0] counts trees whose planes are out, sync[1] trees whose move list is out;
    // the tower waits for sync_target on [0] before it loads planes and on [1] before it gathers legal-move logits.
    unsigned* sync;
    unsigned sync_target;
    int terminal_cap;   // batched select: a tree that absorbed this many terminal visits in one step sits the step out (0 = no cap)
};

struct WarpScratch {
    // score[] (512 B) + sorted_mv[] (256 B) + sorted_ok[] (128 B) + fbuf head are also viewed as
    // 128 doubles by pick_once, so they stay first, contiguous and 8-byte aligned
    alignas(8) int score[MAX_MOVES];
    u16 sorted_mv[MAX_MOVES];
    u8 sorted_ok[MAX_MOVES];
    float fbuf[MAX_MOVES];
    u16 tmp_out[MAX_MOVES];
    MoveList ml;
};
static_assert(offsetof(WarpScratch, fbuf) + sizeof(float) * MAX_MOVES >= sizeof(double) * MAX_MOVES, "pick_once scratch view");

}  // namespace kb
