/* TEST INFRASTRUCTURE ONLY -- the oracle is the checker, never the product.
 *
 * Plain-C restatement of the CPU algorithm of kami's self-play hot path
 * (codeandkey/kami: kami/env.h, kami/mcts.h, kami/chess/neocortex/{position,board,attacks,
 * types,zobrist}.c, eval.h).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this library.
 *
 * PARITY PINNED: tests/test_oracle_vs_reference.py checks every function here against the
 * unmodified reference compiled from /root/reference (oracle/_ref/libkami_ref_core.so), and
 * tests/test_oracle_golden.py checks it against the reference's one deterministic golden
 * (test/encoding.cpp: 258-move game, sha256 26545df8...c0e409) and against committed
 * fixtures in tests/golden/ generated from the reference by tests/golden/make_golden.py.
 */
#ifndef KAMI_ORACLE_H
#define KAMI_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OK_NFEATURES 30
#define OK_PSIZE 4672
#define OK_OBSIZE 1920
#define OK_MAX_MOVES 256
#define OK_MAX_PLY 4096

/* One position snapshot (copy-make; the reference uses make/unmake on one board). */
typedef struct {
    uint64_t pieces[6]; /* by type: P N B R Q K */
    uint64_t colors[2]; /* 0 white, 1 black */
    int8_t sq[64];      /* piece code = type*2 + color, or -1 */
    uint64_t board_key; /* xor of piece-square keys */
    uint64_t key;       /* full position key (board ^ ep ^ castle ^ btm) */
    int ctm;            /* 0 white, 1 black */
    int castle;         /* WK=1 WQ=2 BK=4 BQ=8 */
    int ep;             /* en-passant square or -1 */
    int hmc;            /* halfmove clock */
    int check;          /* side to move in check */
} ok_pos;

typedef struct {
    ok_pos stack[OK_MAX_PLY];
    int n; /* number of snapshots; current = stack[n-1]; Env::ply() == n-1 */
    int actions[OK_MAX_MOVES];
    int n_actions;
    int actions_valid;
} ok_env;

void ok_init(void); /* tables + zobrist keys (glibc rand() seed-1 stream, zobrist.c:25-54) */

/* Env (kami/env.h) */
ok_env* ok_env_new(void);
void ok_env_free(ok_env* e);
void ok_env_reset(ok_env* e);
int ok_env_ply(const ok_env* e);
float ok_env_turn(const ok_env* e);
int ok_env_encode(const ok_env* e, int move);
int ok_env_decode(const ok_env* e, int action);
void ok_env_observe(const ok_env* e, float* dst);
void ok_env_push(ok_env* e, int action);
void ok_env_pop(ok_env* e);
int ok_env_actions(ok_env* e, int* out, int cap);
int ok_env_terminal(ok_env* e, float* value, int* reason);
float ok_env_bootstrap(const ok_env* e, float window);
int ok_env_eval(const ok_env* e);
uint64_t ok_env_key(const ok_env* e);
int ok_env_hmc(const ok_env* e);
int ok_env_check(const ok_env* e);
int ok_env_repcount(const ok_env* e);
int ok_env_castle(const ok_env* e);
int ok_env_ep(const ok_env* e);
int ok_env_piece_at(const ok_env* e, int sq);
/* compact 80-byte wire form shared with the CUDA path (include/kami_b200.h: kb_position) */
void ok_env_export(const ok_env* e, void* out80);

uint64_t ok_zobrist_piece(int sq, int p);
uint64_t ok_zobrist_castle(int r);
uint64_t ok_zobrist_ep(int f);
uint64_t ok_zobrist_btm(void);

/* MCTS (kami/mcts.h) */
typedef struct {
    float cpuct;              /* options cpuct, default 1.0 (mcts.h:89) */
    int force_expand_unvisited;
    int unvisited_node_value_pct; /* default 100 */
    int bootstrap_weight;     /* percent, default 0 */
    int bootstrap_window;     /* default 1600 */
    int bootstrap_amp_pct;    /* default 75 */
    int scale_cpuct_by_actions;
    float noise_weight;       /* bit-parity requires 0 (reference noise is time-seeded) */
    uint64_t seed;
} ok_mcts_cfg;

typedef struct ok_mcts ok_mcts;
void ok_mcts_default_cfg(ok_mcts_cfg* c);
ok_mcts* ok_mcts_new(const ok_mcts_cfg* c);
void ok_mcts_free(ok_mcts* t);
int ok_mcts_n(const ok_mcts* t);
int ok_mcts_select(ok_mcts* t, float* obs);
void ok_mcts_expand(ok_mcts* t, const float* policy, float value, int disable_bootstrap);
int ok_mcts_pick(ok_mcts* t, float alpha, double u01); /* u01 replaces rand()/RAND_MAX */
int ok_mcts_push(ok_mcts* t, int action);
void ok_mcts_reset(ok_mcts* t);
void ok_mcts_snapshot(const ok_mcts* t, float* pspace);
ok_env* ok_mcts_env(ok_mcts* t);
int ok_mcts_root_children(const ok_mcts* t, int* action, int* n, float* w, float* p, int cap);
float ok_mcts_root_w(const ok_mcts* t);
uint64_t ok_mcts_digest(const ok_mcts* t, long* count);

#ifdef __cplusplus
}
#endif
#endif
