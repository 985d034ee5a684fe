/* kami_b200 -- C ABI of the B200-native self-play hot path of codeandkey/kami.
 *
 * Everything below runs as hand-written sm_100a CUDA kernels; there is no CPU fallback.
 * Entry points take plain pointers and sizes.  Unless a parameter is named *_dev, pointers
 * are HOST pointers and the call stages data itself (that is the drop-in, reference-shaped
 * surface); *_dev pointers are device pointers for callers that keep data resident in HBM.
 * All functions return 0 on success or a negative kb_status; kb_last_error() describes the
 * last failure on the calling thread.
 *
 * Each group cites the reference interface (file:line in codeandkey/kami) it replaces.
 * INTEGRATION.md shows the reference-side binding (kami/*.h shims in this repo).
 */
#ifndef KAMI_B200_H
#define KAMI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB_NFEATURES 30   /* kami/env.h:19 */
#define KB_PSIZE 4672     /* kami/env.h:20 */
#define KB_OBSIZE 1920    /* kami/env.h:23 */
#define KB_MAX_ACTIONS 128 /* neocortex/position.h:19 NC_MAX_PL_MOVES */
#define KB_VALUE_WIDTH 256 /* kami/nn/nn.cpp:52 valuefc = Linear(64, 256) */

typedef enum {
    KB_OK = 0,
    KB_ERR_CUDA = -1,       /* a CUDA call failed or no usable device */
    KB_ERR_ARG = -2,        /* bad argument */
    KB_ERR_STATE = -3,      /* call not valid in the current state (e.g. expand without select) */
    KB_ERR_CAPACITY = -4,   /* node pool / path / history capacity exceeded */
    KB_ERR_NAN = -5,        /* network output contains NaN (nn.cpp:176-180) */
    KB_ERR_NO_CHILD = -6,   /* "no child for action" (mcts.h:129) / "no children to pick from" (:139) */
    KB_ERR_UNSUPPORTED = -7
} kb_status;

/* Compact position, 80 bytes, the unit the encoder and tree kernels read.
 * Replaces neocortex ncPosition (position.h:24-47, 98 656 B) on the hot path. */
typedef struct kb_position {
    uint64_t pieces[6];  /* occupancy by type: P N B R Q K (types.h:38-44) */
    uint64_t white;      /* white occupancy; black = (OR pieces) ^ white */
    uint64_t board_key;  /* xor of piece-square zobrist keys (board.c:87,106) */
    uint64_t key;        /* position key: board ^ ep ^ castle ^ btm (position.c:301-311) */
    uint8_t ctm;         /* 0 white, 1 black */
    uint8_t castle;      /* WK=1 WQ=2 BK=4 BQ=8 (position.h:14-17) */
    uint8_t ep;          /* en-passant square, 0xFF = none */
    uint8_t hmc;         /* halfmove clock */
    uint16_t ply;        /* Env::ply() (env.h:58) */
    uint8_t check;       /* side to move is in check */
    uint8_t pad;
} kb_position;

/* ---- library --------------------------------------------------------------------------
 * Threading (reference: selfplay.cpp:25-31 inference / training threads over one NN, nn.cpp:164-168): every host
 * thread has its own CUDA stream per device inside the library; kb_pool / kb_env / kb_trainer are owned by one
 * thread at a time like the reference's MCTS / Env; a kb_net may be used from any number of threads at once
 * (kb_net_infer, kb_pool_step, ... hold its weights under a shared lock, kb_net_load_blob takes it exclusively).
 * Devices: kb_init(d) binds the CALLING THREAD to device d; objects remember the device they were created on and
 * their entry points run there, so one process can drive every GPU of a box with one host thread per GPU. */
int kb_init(int device);          /* bind the calling thread to `device` (first call per device uploads tables); < 0: default */
int kb_current_device(void);      /* device the calling thread is bound to, -1 before kb_init */
const char* kb_last_error(void);
int kb_device_count(void);
int kb_device_name(char* out, int cap);
int kb_sm_count(void);

/* ---- Env: kami/env.h:41-485 ---------------------------------------------------------------
 * A kb_env is a device-resident game: a stack of kb_position (push/pop) whose keys are the
 * repetition history.  One warp serves one env. */
typedef struct kb_env kb_env;
int kb_env_create(kb_env** out);                       /* Env() env.h:52-56 */
int kb_env_destroy(kb_env* e);
int kb_env_reset(kb_env* e);
int kb_env_ply(kb_env* e, int* ply);                   /* env.h:58 */
int kb_env_push(kb_env* e, int action);                /* env.h:264-271 */
int kb_env_pop(kb_env* e);                             /* env.h:273-279 */
int kb_env_actions(kb_env* e, int32_t* out, int cap, int* n); /* env.h:398-423 */
int kb_env_observe(kb_env* e, float* obs);             /* env.h:202-262, [64][30] fp32 */
int kb_env_terminal(kb_env* e, int* terminal, float* value, int* reason); /* env.h:288-391 */
int kb_env_encode(kb_env* e, int move, int* action);   /* env.h:60-143 */
int kb_env_decode(kb_env* e, int action, int* move);   /* env.h:145-200 */
int kb_env_bootstrap(kb_env* e, float window, float* out); /* env.h:476-484 */
int kb_env_position(kb_env* e, kb_position* out);      /* current compact position */

/* ---- batched position kernels (the B200 replacements of Env on many positions) ----------
 * hist_dev/hist_len may be NULL (no repetition history). */
int kb_encode_planes(const kb_position* pos, int n, float* obs);  /* Env::observe x n, fp32 */
int kb_legal_actions(const kb_position* pos, int n, int32_t* actions /*[n][128]*/, int32_t* counts);
int kb_apply_actions(kb_position* pos /*in,out*/, int n, const int32_t* actions);
int kb_static_eval(const kb_position* pos, int n, int32_t* eval); /* ncPositionEvaluate position.c:1082 */
/* device-resident variants used by bench.py (positions and outputs already in HBM) */
int kb_encode_planes_dev(const kb_position* pos_dev, int n, float* obs_dev);
int kb_encode_planes_bf16_dev(const kb_position* pos_dev, int n, void* planes_dev);
int kb_legal_actions_dev(const kb_position* pos_dev, int n, int32_t* actions_dev, int32_t* counts_dev);

/* ---- NN: kami/nn/nn.h:40-73, kami/nn/nn.cpp:26-34,59-91,155-187 -------------------------- */
typedef struct kb_net kb_net;
int kb_net_create(kb_net** out, int filters, int residuals);   /* NN(8,8,30,4672) nn.cpp:107 */
int kb_net_destroy(kb_net* net);
/* fp32 tensors in oracle/nn_oracle.py:param_order (reference register_module names) */
size_t kb_net_blob_floats(int filters, int residuals);
int kb_net_load_blob(kb_net* net, const float* blob, size_t n_floats);
/* NN::infer (nn.cpp:155-187): value[i] = vh.flat[i] (the reference's memcpy of a [B,256] tensor) */
int kb_net_infer(kb_net* net, const float* obs, int batch, float* policy, float* value);
/* all 256 value-head outputs per position, for tolerance tests */
int kb_net_forward_full(kb_net* net, const float* obs, int batch, float* policy, float* value256);
/* device-resident: planes_dev is the bf16 layout kb_encode_planes_bf16_dev writes */
int kb_net_forward_dev(kb_net* net, const void* planes_dev, int batch, float* policy_dev, float* value256_dev);
size_t kb_net_planes_bytes(int batch);
/* per-layer timing of the last forward (CUDA events), ms; names via kb_net_layer_name */
int kb_net_flops(kb_net* net, double* tower_flops, double* heads_flops); /* per position */
/* test hook: one board of an internal activation tensor as fp32 [channels][64] */
int kb_net_debug_activation(kb_net* net, int which, int board, float* out, int* channels);
/* profiling hook: clock64 stamps of the fused tower kernel's phases (CTA 0) */
int kb_net_debug_timestamps(kb_net* net, int enable, long long* out, int cap, int* count);
/* profiling hook: globaltimer (ns) at entry and exit of every CTA of the last fused tower launch, out[cta][2] */
int kb_net_debug_cta_spans(kb_net* net, long long* out, int cap_ctas, int* count);

/* ---- MCTS: kami/mcts.h:15-349; Selfplay::inference_main kami/selfplay.cpp:58-213 ---------- */
typedef struct kb_tree_cfg {
    float cpuct;                 /* mcts.h:89 */
    int force_expand_unvisited;  /* mcts.h:90 */
    int unvisited_node_value_pct; /* mcts.h:91 */
    int bootstrap_weight;        /* percent, mcts.h:92 */
    int bootstrap_window;        /* mcts.h:93 */
    int bootstrap_amp_pct;       /* mcts.h:94 */
    int scale_cpuct_by_actions;  /* mcts.h:95 */
    float noise_weight;          /* mcts.h:97 */
    uint64_t seed;               /* replaces rng.seed(time(NULL)) mcts.h:99 and rand() :173 */
    /* self-play schedule (selfplay.cpp:16-17, 61-75) */
    int selfplay_nodes;
    float alpha_initial, alpha_decay, alpha_final;
    int alpha_cutoff;
    int draw_value_pct;
    int value_index_mode;        /* 0 = reference (value[i] = vh.flat[i]), 1 = vh[i][0] */
} kb_tree_cfg;
int kb_tree_default_cfg(kb_tree_cfg* cfg);   /* the reference's in-code defaults */

typedef struct kb_pool kb_pool;
int kb_pool_create(kb_pool** out, int n_trees, int node_capacity, const kb_tree_cfg* cfg);
int kb_pool_destroy(kb_pool* p);
int kb_pool_size(kb_pool* p);

/* single-tree API, same call protocol as kami::MCTS */
int kb_tree_n(kb_pool* p, int tree, int* n);                       /* mcts.h:111 */
int kb_tree_select(kb_pool* p, int tree, float* obs, int* need_eval); /* mcts.h:186-255 */
int kb_tree_expand(kb_pool* p, int tree, const float* policy, float value, int disable_bootstrap); /* :257-327 */
int kb_tree_pick(kb_pool* p, int tree, float alpha, double u01, int* action); /* :137-184; u01 = rand()/RAND_MAX */
int kb_tree_push(kb_pool* p, int tree, int action);                /* :113-135 */
int kb_tree_reset(kb_pool* p, int tree);                           /* :331-339 */
int kb_tree_snapshot(kb_pool* p, int tree, float* pspace);         /* :341-348 */
int kb_tree_root_children(kb_pool* p, int tree, int32_t* action, int32_t* n, float* w, float* prior, int cap, int* count);
int kb_tree_root_w(kb_pool* p, int tree, float* w);
/* actions from the root to the pending leaf: the reference's Env sits AT the leaf after select() (mcts.h:252-254) */
int kb_tree_leaf_path(kb_pool* p, int tree, int32_t* actions, int cap, int* depth);
int kb_tree_digest(kb_pool* p, int tree, uint64_t* digest, int64_t* count);
int kb_tree_env(kb_pool* p, int tree, kb_position* out);           /* MCTS::get_env() root position */

/* batched phases over all trees of the pool (device-resident) */
int kb_pool_select(kb_pool* p);                 /* every tree: one leaf (terminals absorbed, moves made at budget) */
int kb_pool_leaf_positions(kb_pool* p, kb_position* out /*[n_trees] host*/);
/* (a pinned policy array -- kb_host_alloc_pinned / kb_host_register -- is read in place: only the legal moves' entries
 *  cross the link, mcts.h:273; a pageable one is copied in full) */
int kb_pool_expand(kb_pool* p, const float* policy /*[n][4672] host*/, const float* value /*[n] host*/, int disable_bootstrap);
int kb_pool_expand_dev(kb_pool* p, const float* policy_dev, const float* value_dev, int disable_bootstrap);
/* compact forms of the same exchange: every pending leaf's Env::actions() list (row pitch 128, -1 padded) out;
 * prior[t][i] = (unnormalised) policy mass of tree t's i-th legal action + value[t] in (mcts.h:257-327) */
int kb_pool_leaf_actions(kb_pool* p, int32_t* actions /*[n][128] host*/, int32_t* counts /*[n] host*/);
int kb_pool_expand_compact(kb_pool* p, const float* prior /*[n][128] host*/, const float* value /*[n] host*/, int disable_bootstrap);
/* the whole loop of selfplay.cpp:113-200: select -> encode -> tower+heads -> expand/backup, `iters` times */
int kb_pool_step(kb_pool* p, kb_net* net, int iters);
/* same, but every iteration's data goes through the caller's host arrays like the reference's batch / inf_policy /
 * inf_value (selfplay.cpp:107-109): leaf observations out and back in (NN::infer's input), dense policy rows + values out,
 * values back in.  MCTS::expand reads policy[action] for the legal moves only (mcts.h:273): with a pinned policy array
 * (kb_host_alloc_pinned / kb_host_register) the expand kernels read those entries straight from it; a pageable one is
 * copied back to the device in full.  Used for the end-to-end bench figure. */
int kb_pool_step_hostio(kb_pool* p, kb_net* net, int iters, float* obs_host, float* policy_host, float* value_host);
/* kb_pool_step_hostio serves the trees as `groups` independent pipelines, the analogue of the reference's
 * inference_threads (selfplay.cpp:25-31): each group has its own NN::infer batch (so Q1's value indexing is per
 * group) and its own streams, so one group's transfers overlap another's compute.  0 = default (4), max 8. */
int kb_pool_set_hostio_groups(kb_pool* p, int groups);
/* kb_pool_step_hostio with the compact forms on the wire: 80-byte leaf positions (instead of [1920] fp32 observation
 * rows) out and back in, [128] legal-move priors (instead of [4672] policy rows) + value out and back in */
int kb_pool_step_hostio_compact(kb_pool* p, kb_net* net, int iters, kb_position* leaf_host /*[n]*/, float* prior_host /*[n][128]*/,
                                float* value_host /*[n]*/);
/* kb_pool_step serves the trees as `groups` independent pipelines (the analogue of inference_threads, selfplay.cpp:25-31:
 * own trees, own NN batch -- Q1's value indexing is per group -- own stream); 0 = default (KB_STEP_GROUPS or 1), max 8 */
int kb_pool_set_step_groups(kb_pool* p, int groups);
/* phase timing of kb_pool_step (kb_pool_last_phase_ms) is opt-in: when on, up to 32 iterations of a call are bracketed
 * by CUDA events and not fused with their neighbours; off (default) the call records nothing and always fuses */
int kb_pool_set_profiling(kb_pool* p, int on);
/* Terminal leaves are absorbed inside `while (n < nodes && !select())` (selfplay.cpp:133); a tree deep in the 50-ply
 * rule can absorb a hundred of them in one step while every other tree of the batch waits.  With a cap > 0 such a tree
 * sits the step out after `cap` absorbed visits (its own visit sequence is unchanged; the batch goes out one leaf short;
 * kb_pool_stats.evals counts real evaluations only; with NN::infer's value indexing, value_index_mode 0, a tree that sits
 * out leaves its board's previous planes in the batch).  0 (default) = the reference's batch: always one leaf per tree. */
int kb_pool_set_terminal_cap(kb_pool* p, int max_terminal_visits);
/* Experimental ("split select"): kb_pool_step with a terminal cap publishes every leaf's input planes BEFORE it generates
 * the leaf's moves, and the tower, already launched, starts on the planes while the move generators still run.  Correct
 * (tests) but not faster on B200: -1 (default) = off unless KB_SPLIT_SELECT=1; 0 off; 1 on.  A leaf that
 * turns out to be mate / stalemate then sits its step out. */
int kb_pool_set_split_select(kb_pool* p, int mode);
/* change the node budget per move of a live pool (bench.py ages its synthetic games quickly with a small budget) */
int kb_pool_set_selfplay_nodes(kb_pool* p, int nodes);
/* flush_old_trees (selfplay.cpp:61,119-131): MCTS::reset on every tree, partial trajectories dropped */
int kb_pool_flush_trees(kb_pool* p);

typedef struct kb_pool_stats {
    uint64_t evals;          /* NN evaluations (leaves expanded) */
    uint64_t moves;          /* game moves played (positions) */
    uint64_t games;          /* games finished */
    uint64_t terminal_visits;/* simulations that ended in a terminal leaf */
    uint64_t children_scanned; /* PUCT child records read (roofline bytes) */
    uint64_t path_nodes;     /* nodes updated by backup */
    uint64_t children_created;
    uint64_t samples;        /* replay samples emitted */
    uint64_t nodes_in_use;   /* sum over trees */
    uint64_t kernel_launches;/* kernels this pool launched */
    uint64_t skipped_leaves; /* tree-steps that went out without a leaf (kb_pool_set_terminal_cap) */
} kb_pool_stats;
int kb_pool_get_stats(kb_pool* p, kb_pool_stats* out);
int kb_pool_reset_stats(kb_pool* p);
/* kb_pool_step policy path: 0 (default) = softmax over each leaf's legal moves only, fused behind the policy head and
 * equal to nn.cpp:80 followed by MCTS::expand's renormalisation (mcts.h:273-276,296); 1 = dense [n][4672] softmax */
int kb_pool_set_policy_mode(kb_pool* p, int dense);
/* replay sink: finished-game samples (selfplay.cpp:141-188) kept on the device as sparse rows */
int kb_pool_drain_samples(kb_pool* p, int max_samples, float* obs /*[m][1920]*/, float* pi /*[m][4672]*/, float* z /*[m]*/, int* count);

/* Selfplay::get_next_pgn (selfplay.h:73-80, selfplay.cpp:167-171): request the action list of the next self-play game
 * that finishes inside kb_pool_step, then poll for it (count = 0 until one has finished).  Replaying the actions on a
 * kami::Env gives Env::pgn(). */
int kb_pool_request_game(kb_pool* p);
int kb_pool_take_game(kb_pool* p, int32_t* actions, int cap, int* count);

/* ---- arena: kami::eval, kami/evaluate.h:6, kami/evaluate.cpp:10-160 --------------------------------------------
 * `games` device-resident trees searched `nodes` deep; every leaf is evaluated by the network whose turn it is at the
 * leaf (two batches of at most `batch` rows per round, bootstrap off), greedy moves.  One kb_arena_round is one pass of
 * the reference's while-loop body (evaluate.cpp:45-151); the caller keeps the score and the early-stop rule
 * (evaluate.cpp:100-126) by reading the finished games, which arrive in the reference's order.  colours[i] = the side
 * the candidate plays in tree i (evaluate.cpp:18-22: (rand() % 2) * 2 - 1). */
typedef struct kb_arena kb_arena;
typedef struct kb_arena_game {
    int32_t tree;    /* tree whose game ended */
    float result;    /* terminal value, White's point of view (env.h:288-385) */
    int32_t colour;  /* side the candidate played in that game: score += result * colour / 2 + 0.5 (evaluate.cpp:101) */
} kb_arena_game;
int kb_arena_create(kb_arena** out, int games, int batch, int nodes, const kb_tree_cfg* cfg, const int32_t* colours, int n_colours);
int kb_arena_destroy(kb_arena* a);
int kb_arena_round(kb_arena* a, kb_net* current, kb_net* candidate, kb_arena_game* finished, int cap, int* n_finished);
/* the same round in two halves, for hosts that evaluate the leaves themselves: begin() hands out the two networks' input
 * batches exactly as the reference passes them to NN::infer (host fp32 [n][1920]), end() takes policy [n][4672] / value [n] */
int kb_arena_begin(kb_arena* a, kb_arena_game* finished, int cap, int* n_finished, float* cur_obs, int* cur_n, float* cd_obs, int* cd_n);
int kb_arena_end(kb_arena* a, const float* cur_policy, const float* cur_value, const float* cd_policy, const float* cd_value);
kb_pool* kb_arena_pool(kb_arena* a);                      /* the trees, for kb_tree_digest / kb_tree_n / ... */
int kb_arena_colours(kb_arena* a, int32_t* out, int cap); /* current colour table */

/* timing helper: elapsed ms of the last kb_pool_step / kb_net_forward_dev per phase */
typedef struct kb_phase_ms {
    float select, encode, tower, heads, expand, total;
} kb_phase_ms;
int kb_pool_last_phase_ms(kb_pool* p, kb_phase_ms* out);
/* debug: per-tree cycle counters of the last batched select, 8 int64 per tree (see tree.cu) */
int kb_pool_debug_select_profile(kb_pool* p, int enable, long long* out, int cap_trees);

/* raw device memory helpers so that hosts without a CUDA runtime binding (ctypes) can keep
 * buffers resident */
int kb_dev_alloc(void** out, size_t bytes);
int kb_dev_free(void* ptr);
int kb_dev_upload(void* dst_dev, const void* src_host, size_t bytes);
int kb_dev_download(void* dst_host, const void* src_dev, size_t bytes);
int kb_dev_sync(void);
int kb_host_alloc_pinned(void** out, size_t bytes);
int kb_host_free_pinned(void* ptr);
/* pin a caller-owned host buffer (cudaHostRegister) so the host-pointer entry points DMA straight from / into it */
int kb_host_register(void* ptr, size_t bytes);
int kb_host_unregister(void* ptr);
/* CUDA-event timer and L2 flush on the library's stream (bench.py) */
int kb_timer_start(void);
int kb_timer_stop(float* ms);
int kb_flush_l2(size_t bytes);
/* capture window for profilers started with "profile from start off" (cudaProfilerStart / cudaProfilerStop) */
int kb_profiler_start(void);
int kb_profiler_stop(void);

/* ---------------------------------------------------------------------------------------------
 * Training step: NN::train's mini-batch (kami/nn/nn.cpp:224-377) = NNModule::forward in
 * training mode (BatchNorm batch statistics), NNModule::loss (nn.cpp:93-105), backward, plain SGD
 * (nn.cpp:239-241).  fp32 master weights / gradients / running statistics in the blob order of
 * kb_net_load_blob; bf16 tcgen05 convolutions (forward, dgrad, wgrad).  Data-parallel hosts all-reduce
 * the buffer kb_trainer_grad_buffer returns (NCCL) between forward_backward and apply_sgd.
 * ------------------------------------------------------------------------------------------- */
typedef struct kb_trainer kb_trainer;
int kb_trainer_create(kb_trainer** out, int filters, int residuals, int max_batch);
int kb_trainer_destroy(kb_trainer* t);
int kb_trainer_load_blob(kb_trainer* t, const float* blob, size_t n_floats);
int kb_trainer_export_blob(kb_trainer* t, float* blob, size_t n_floats);
int kb_trainer_export_grads(kb_trainer* t, float* out, size_t n_floats);
int kb_trainer_grad_buffer(kb_trainer* t, void** dev_ptr, size_t* n_floats);
/* inputs [batch][1920], obs_p [batch][4672], obs_v [batch]: the arrays of NN::train (nn.h:67); host pointers */
int kb_trainer_forward_backward(kb_trainer* t, const float* obs, const float* obs_p, const float* obs_v, int batch, float* loss);
int kb_trainer_forward_backward_dev(kb_trainer* t, const float* obs_dev, const float* obs_p_dev, const float* obs_v_dev, int batch, float* loss);
int kb_trainer_apply_sgd(kb_trainer* t, float lr, float grad_scale);
/* the flat parameter vector on the device (blob order) and the (offset, count) ranges of it that hold BatchNorm running
 * statistics: data-parallel hosts average those over the replicas, the rest moves by the all-reduced gradient */
int kb_trainer_param_buffer(kb_trainer* t, void** dev_ptr, size_t* n_floats);
int kb_trainer_stat_ranges(kb_trainer* t, size_t* offsets, size_t* counts, int cap, int* n);
/* test hook: one board of a saved training activation, fp32 [channels][64] (which: 0 conv output, 1 layer output) */
int kb_trainer_debug_activation(kb_trainer* t, int layer, int which, int board, float* out, int* channels);

/* ---- data-parallel training from one host process (BASELINE config 5) ------------------------------------
 * A kb_trainer replica per GPU, one NCCL all-reduce(sum) of the flat fp32 gradient bucket per mini-batch over NVLink, the
 * same SGD step on every replica.  NCCL is bound at run time (libnccl.so.2); without it kb_dp_create reports
 * KB_ERR_UNSUPPORTED and everything else in the library still works. */
typedef struct kb_dp kb_dp;
int kb_dp_create(kb_dp** out, const int* devices, int n, int filters, int residuals, int max_batch_per_device);
int kb_dp_destroy(kb_dp* d);
int kb_dp_size(kb_dp* d);
kb_trainer* kb_dp_replica(kb_dp* d, int rank);
int kb_dp_load_blob(kb_dp* d, const float* blob, size_t n_floats);
int kb_dp_export_blob(kb_dp* d, int rank, float* blob, size_t n_floats);
/* rank r takes rows [r * batch_per_device, (r + 1) * batch_per_device) of the host arrays (NN::train's arrays, nn.h:67) */
int kb_dp_step(kb_dp* d, const float* obs, const float* obs_p, const float* obs_v, int batch_per_device, float lr, float grad_scale, float* loss);
/* every replica's batch already resident on its own GPU (arrays of n device pointers) */
int kb_dp_step_dev(kb_dp* d, const float* const* obs_dev, const float* const* obs_p_dev, const float* const* obs_v_dev, int batch_per_device, float lr,
                   float grad_scale);
int kb_dp_allreduce_apply(kb_dp* d, float lr, float grad_scale);

#ifdef __cplusplus
}
#endif
#endif /* KAMI_B200_H */
