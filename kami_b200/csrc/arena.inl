// Device-resident arena: kami::eval (kami/evaluate.h:6, kami/evaluate.cpp:10-160) on the pool kernels.
// Included at the end of tree.cu (same translation unit: it uses select_once / pick_once / push_once / expand_once).
//
// The reference's arena is the second driver of the hot path: `evaluate_games` trees searched `evaluate_nodes` deep,
// each leaf evaluated by the network whose turn it is AT THE LEAF (two NN::infer calls per round of at most
// `evaluate_batch` rows each, bootstrap off, evaluate.cpp:136-151), greedy moves, one game result per finished game.
// One pass of its while-loop body is one kb_arena_round:
//
//   k_arena_advance   every tree (a warp each): moves at the node budget (pick(0) + push, game end -> result + reset),
//                     terminal leaves absorbed, until a leaf waits for a network          evaluate.cpp:72-96, 128-131
//   host              the reference's sequential bookkeeping over the trees IN ORDER: which buffer and slot the leaf's
//                     observation lands in, which network's batch it is filed under, batch caps, colours  evaluate.cpp:60-93
//   k_arena_route     leaf planes (bf16 tall layout for the tower, fp32 rows for the split API) into the persistent
//                     per-network input buffers -- slots that are not written keep what they held (see below)
//   tower x 2         current / candidate network on its own batch (dense softmax, NN::infer's value indexing, Q1)
//   k_arena_expand    MCTS::expand(policy row, value, disable_bootstrap = true) for every filed tree  evaluate.cpp:141,150
//
// Reference behaviour kept on purpose: the destination buffer and slot of an observation are chosen from the side to
// move at the ROOT before the descent (evaluate.cpp:68-70) while the leaf is filed under the side to move at the LEAF
// (:80-90) -- for leaves at odd depth the observation lands in the other network's buffer and the network that is asked
// evaluates whatever its slot held before; a recycled tree's colour is +1 when its last root belonged to the
// candidate's buffer, else -1 (:107); trees whose network batch is full keep their leaf for the next round; when both
// batches are full the remaining trees are not touched this round (:62-63) -- the advance kernel therefore runs over
// chunks of trees no longer than the free slots.  Dropped: the out-of-bounds write when the destination buffer is full.

namespace kb {

struct ArenaOut {   // per tree, per advance
    int finished;   // a game ended while this tree was advanced (at most one per round: a fresh tree needs a network first)
    float result;   // its terminal value, White's point of view (env.h:288-385)
    int mover_ctm;  // side to move at the root before the game's last move (0 white, 1 black)
    int has_leaf;   // a leaf waits for a network
    int root_ctm;   // side to move at the root when that leaf was selected
    int leaf_ctm;   // side to move at the leaf
    int moves;      // moves made during this advance
    int pad;
};
struct ArenaJob {
    int tree, which, slot, pad;  // which: 0 current network, 1 candidate
};

__device__ void arena_advance_tree(const PoolDev& P, int t, WarpScratch& s, ArenaOut* out) {
    TreeCtl& c = P.ctl[t];
    ArenaOut o = {0, 0.0f, 0, 0, 0, 0, 0, 0};
    for (int guard = 0; guard < (1 << 20); ++guard) {
        if (*P.error) return;
        const int rn = tree_nodes(P, t, c.space)[c.root].n;
        if (rn >= P.cfg.selfplay_nodes) {  // evaluate.cpp:95-96: trees[i].push(trees[i].pick())
            const int mover = c.root_pos.ctm;
            const int action = pick_once(P, t, 0.0f, 0.0, s);
            if (action < 0) {  // (Q15) pick() returns -1 when no child was visited: push(-1) throws "no child for action"
                raise(P, KB_ERR_NO_CHILD);
                return;
            }
            if (!push_once(P, t, action)) return;
            o.moves += 1;
            if (lane_id() == 0) {
                c.moves += 1;
                atomicAdd(&P.stats->moves, 1ULL);
            }
            float value;
            if (root_terminal(P, t, s, &value)) {  // evaluate.cpp:98-131
                o.finished = 1;
                o.result = value;
                o.mover_ctm = mover;
                if (lane_id() == 0) {
                    c.games += 1;
                    atomicAdd(&P.stats->games, 1ULL);
                }
                __syncwarp();
                tree_reset(P, t);
            }
            continue;
        }
        if (select_once(P, t, s)) {
            o.has_leaf = 1;
            break;
        }
    }
    o.root_ctm = c.root_pos.ctm;
    o.leaf_ctm = c.leaf_pos.ctm;
    if (lane_id() == 0) out[t] = o;
}
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_arena_advance(PoolDev P, ArenaOut* out) {
    __shared__ WarpScratch scratch[WARPS_PER_BLOCK];
    const int w = threadIdx.x >> 5;
    const int t = P.tree0 + blockIdx.x * WARPS_PER_BLOCK + w;
    if (t >= P.tree_hi) return;
    arena_advance_tree(P, t, scratch[w], out);
}
// leaf of jobs[j].tree -> input slot jobs[j].slot of network jobs[j].which: fp32 row [64][30] and bf16 tall planes
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_arena_route(PoolDev P, const ArenaJob* jobs, int njobs, float* obs0, float* obs1, uint4* planes0,
                                                                      uint4* planes1) {
    __shared__ WarpScratch scratch[WARPS_PER_BLOCK];
    const int w = threadIdx.x >> 5;
    const int j = blockIdx.x * WARPS_PER_BLOCK + w;
    if (j >= njobs) return;
    const ArenaJob job = jobs[j];
    const Pos pos = P.ctl[job.tree].leaf_pos;
    WarpScratch& s = scratch[w];
    warp_encode_f32(pos, (job.which ? obs1 : obs0) + (size_t)job.slot * KB_OBSIZE, s.fbuf, reinterpret_cast<signed char*>(s.sorted_ok));
    warp_encode_tall(pos, job.slot, job.which ? planes1 : planes0);
}
// MCTS::expand(policy + row * PSIZE, value[row], true) (evaluate.cpp:141, 150); value rows are NN::infer's (Q1)
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_arena_expand(PoolDev P, const ArenaJob* jobs, int njobs, const float* pol0, const float* pol1,
                                                                       const float* val0, const float* val1, int value_stride) {
    __shared__ WarpScratch scratch[WARPS_PER_BLOCK];
    const int w = threadIdx.x >> 5;
    const int j = blockIdx.x * WARPS_PER_BLOCK + w;
    if (j >= njobs) return;
    if (*P.error) return;
    const ArenaJob job = jobs[j];
    const float* pol = (job.which ? pol1 : pol0) + (size_t)job.slot * PSIZE;
    const float v = (job.which ? val1 : val0)[(size_t)job.slot * value_stride];
    expand_once(P, job.tree, pol, v, true, scratch[w]);
}

}  // namespace kb

struct kb_arena {
    kb_pool* pool = nullptr;
    int games = 0, batch = 0, nodes = 0;
    int value_index_mode = 0;
    std::vector<int> colour;         // side the candidate plays in each tree (evaluate.cpp:18-22, 107)
    NetWs ws[2];                     // persistent input planes of the current / candidate network (+ activations)
    float* obs_dev[2] = {nullptr, nullptr};   // the same inputs as fp32 rows [batch][1920] (split API)
    float* pol_dev[2] = {nullptr, nullptr};   // [batch][4672]
    float* val_dev[2] = {nullptr, nullptr};   // [batch][256]
    ArenaOut* out_dev = nullptr;
    ArenaJob* jobs_dev = nullptr;
    std::vector<ArenaOut> out_host;
    std::vector<int> filed[2];       // trees filed under each network this round, in order (cur_targets / cd_targets)
    std::vector<char> unfiled;       // tree kept its leaf from an earlier round (its network's batch was full)
    bool pending = false;            // begin() done, end() due
    unsigned long long rounds = 0;
};

namespace {

// advance + the reference's sequential bookkeeping; fills a->filed and the route jobs, runs k_arena_route
int arena_begin_impl(kb_arena* a, kb_arena_game* finished, int cap, int* n_finished) {
    kb_pool* p = a->pool;
    cudaStream_t st = main_stream();
    *n_finished = 0;
    a->filed[0].clear();
    a->filed[1].clear();
    std::vector<ArenaJob> route;
    int nfin = 0;
    for (int i0 = 0; i0 < a->games;) {
        const int free_slots = (a->batch - (int)a->filed[0].size()) + (a->batch - (int)a->filed[1].size());
        if (free_slots <= 0) break;  // evaluate.cpp:62-63: both batches full, the remaining trees wait for the next round
        const int hi = i0 + free_slots < a->games ? i0 + free_slots : a->games;
        PoolDev d = p->d;
        d.tree0 = i0;
        d.tree_hi = hi;
        d.defer_compact = 0;  // (warp-level collector inside push, as in the single-tree protocol)
        const int m = hi - i0;
        k_arena_advance<<<(m + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, 32 * WARPS_PER_BLOCK, 0, st>>>(d, a->out_dev);
        KB_CUDA(cudaGetLastError());
        p->launches++;
        KB_CUDA(cudaMemcpyAsync(a->out_host.data() + i0, a->out_dev + i0, sizeof(ArenaOut) * (size_t)m, cudaMemcpyDeviceToHost, st));
        int r = pool_check(p, true);
        if (r) return r;
        for (int i = i0; i < hi; ++i) {
            const ArenaOut& o = a->out_host[(size_t)i];
            if (o.finished) {
                if (nfin < cap) finished[nfin] = kb_arena_game{i, o.result, a->colour[(size_t)i]};
                ++nfin;
                // evaluate.cpp:107: the recycled tree's colour follows the buffer its last root belonged to
                const int mover_turn = o.mover_ctm == 0 ? 1 : -1;
                a->colour[(size_t)i] = mover_turn == a->colour[(size_t)i] ? 1 : -1;
            }
            if (!o.has_leaf) continue;
            const int leaf_turn = o.leaf_ctm == 0 ? 1 : -1;
            // get_env().turn() before the descent: the root's side to move -- except for a tree that still holds last
            // round's leaf, whose Env sits AT that leaf (mcts.h:252-254), so the reference reads the leaf's side there
            const int root_turn = a->unfiled[(size_t)i] ? leaf_turn : (o.root_ctm == 0 ? 1 : -1);
            const int dst = root_turn == a->colour[(size_t)i] ? 1 : 0;  // :68 the buffer, chosen before the descent
            const int dst_slot = (int)a->filed[dst].size();              // :70 and its slot
            const int s = leaf_turn == a->colour[(size_t)i] ? 1 : 0;     // :80-90 the network that is asked
            if ((int)a->filed[s].size() < a->batch) {
                a->filed[s].push_back(i);
                a->unfiled[(size_t)i] = 0;
                if (dst_slot < a->batch) {
                    // the reference copies sequentially: a later tree overwrites an earlier one that chose the same slot
                    // (possible when the earlier leaf was filed under the other network) -- keep one job per slot, the last
                    bool replaced = false;
                    for (ArenaJob& j : route)
                        if (j.which == dst && j.slot == dst_slot) {
                            j.tree = i;
                            replaced = true;
                        }
                    if (!replaced) route.push_back(ArenaJob{i, dst, dst_slot, 0});
                }
            } else {
                a->unfiled[(size_t)i] = 1;
            }
        }
        i0 = hi;
    }
    *n_finished = nfin;
    if (nfin > cap) {
        set_error("%d games finished in one round, caller's buffer holds %d", nfin, cap);
        return KB_ERR_CAPACITY;
    }
    if (!route.empty()) {
        KB_CUDA(cudaMemcpyAsync(a->jobs_dev, route.data(), sizeof(ArenaJob) * route.size(), cudaMemcpyHostToDevice, st));
        const int nj = (int)route.size();
        k_arena_route<<<(nj + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, 32 * WARPS_PER_BLOCK, 0, st>>>(p->d, a->jobs_dev, nj, a->obs_dev[0], a->obs_dev[1],
                                                                                                   ws_planes(a->ws[0]), ws_planes(a->ws[1]));
        KB_CUDA(cudaGetLastError());
        KB_CUDA(cudaStreamSynchronize(st));  // `route` is a pageable host vector
        p->launches++;
    }
    a->pending = true;
    a->rounds++;
    return KB_OK;
}

int arena_expand_impl(kb_arena* a, int value_stride) {
    kb_pool* p = a->pool;
    cudaStream_t st = main_stream();
    std::vector<ArenaJob> jobs;
    for (int w = 0; w < 2; ++w)
        for (size_t k = 0; k < a->filed[w].size(); ++k) jobs.push_back(ArenaJob{a->filed[w][k], w, (int)k, 0});
    a->pending = false;
    if (jobs.empty()) return KB_OK;
    KB_CUDA(cudaMemcpyAsync(a->jobs_dev, jobs.data(), sizeof(ArenaJob) * jobs.size(), cudaMemcpyHostToDevice, st));
    const int nj = (int)jobs.size();
    k_arena_expand<<<(nj + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, 32 * WARPS_PER_BLOCK, 0, st>>>(p->d, a->jobs_dev, nj, a->pol_dev[0], a->pol_dev[1], a->val_dev[0],
                                                                                                a->val_dev[1], value_stride);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    return pool_check(p, true);
}

}  // namespace

extern "C" {

int kb_arena_create(kb_arena** out, int games, int batch, int nodes, const kb_tree_cfg* cfg, const int32_t* colours, int n_colours) {
    KB_REQUIRE_INIT();
    KB_ARG(out && cfg && colours, "out/cfg/colours");
    KB_ARG(games > 0 && batch > 0 && nodes >= 2 && n_colours >= games, "games > 0, batch > 0, nodes >= 2, one colour per tree");
    kb_arena* a = new (std::nothrow) kb_arena();
    if (!a) return KB_ERR_ARG;
    a->games = games;
    a->batch = batch;
    a->nodes = nodes;
    a->value_index_mode = cfg->value_index_mode;
    kb_tree_cfg c = *cfg;
    c.selfplay_nodes = nodes;  // the node budget of the advance kernel (evaluate_nodes)
    int r = kb_pool_create(&a->pool, games, 1 << 16, &c);
    if (r) {
        delete a;
        return r;
    }
    a->colour.assign(colours, colours + games);
    a->out_host.resize((size_t)games);
    a->unfiled.assign((size_t)games, 0);
    for (int w = 0; w < 2; ++w) {
        KB_CUDA(cudaMalloc(&a->obs_dev[w], sizeof(float) * KB_OBSIZE * (size_t)batch));
        KB_CUDA(cudaMemsetAsync(a->obs_dev[w], 0, sizeof(float) * KB_OBSIZE * (size_t)batch, main_stream()));
        KB_CUDA(cudaMalloc(&a->pol_dev[w], sizeof(float) * KB_PSIZE * (size_t)batch));
        KB_CUDA(cudaMalloc(&a->val_dev[w], sizeof(float) * KB_VALUE_WIDTH * (size_t)batch));
    }
    KB_CUDA(cudaMalloc(&a->out_dev, sizeof(ArenaOut) * (size_t)games));
    KB_CUDA(cudaMalloc(&a->jobs_dev, sizeof(ArenaJob) * (size_t)(2 * batch + games)));
    *out = a;
    return KB_OK;
}
int kb_arena_destroy(kb_arena* a) {
    if (!a) return KB_OK;
    kb_pool* p = a->pool;
    KB_BIND(p);
    cudaStreamSynchronize(main_stream());
    for (int w = 0; w < 2; ++w) {
        ws_free(a->ws[w]);
        cudaFree(a->obs_dev[w]); cudaFree(a->pol_dev[w]); cudaFree(a->val_dev[w]);
    }
    cudaFree(a->out_dev); cudaFree(a->jobs_dev);
    kb_pool_destroy(a->pool);
    delete a;
    return KB_OK;
}
kb_pool* kb_arena_pool(kb_arena* a) { return a ? a->pool : nullptr; }
int kb_arena_colours(kb_arena* a, int32_t* out, int cap) {
    KB_ARG(a && out && cap >= a->games, "arena/out/cap");
    for (int i = 0; i < a->games; ++i) out[i] = a->colour[(size_t)i];
    return KB_OK;
}

// One pass of the reference's while-loop body with both networks on the device.
int kb_arena_round(kb_arena* a, kb_net* current, kb_net* candidate, kb_arena_game* finished, int cap, int* n_finished) {
    KB_ARG(a && current && candidate && finished && n_finished && cap > 0, "arena/nets/finished");
    kb_pool* p = a->pool;
    KB_BIND(p);
    KB_ARG(net_device(current) == p->device && net_device(candidate) == p->device, "arena and nets live on different devices");
    KB_ARG(!a->pending, "kb_arena_begin without kb_arena_end");
    kb_net* nets[2] = {current, candidate};
    NetReadGuard l0(current);
    // (the same net on both sides is legal -- one shared lock is enough then; a writer-first lock must not be taken twice)
    struct Second {
        kb_net* n;
        ~Second() { if (n) net_unlock_shared(n); }
    } l1{candidate != current ? candidate : nullptr};
    if (l1.n) net_lock_shared(l1.n);
    cudaStream_t st = main_stream();
    int r;
    for (int w = 0; w < 2; ++w)
        if ((r = ws_reserve(a->ws[w], nets[w], a->batch, st))) return r;
    if ((r = arena_begin_impl(a, finished, cap, n_finished))) return r;
    for (int w = 0; w < 2; ++w) {
        const int n = (int)a->filed[w].size();
        if (!n) continue;
        if ((r = net_forward_async(nets[w], a->ws[w], ws_planes(a->ws[w]), n, a->pol_dev[w], a->val_dev[w], st))) return r;
        p->launches += (unsigned long long)net_launches_per_forward(nets[w]);
    }
    // NN::infer hands out the first n floats of its [n][256] value tensor (Q1); value_index_mode 1 = column 0 of each row
    return arena_expand_impl(a, a->value_index_mode == 1 ? KB_VALUE_WIDTH : 1);
}

// The same round in two halves for hosts that evaluate the leaves themselves: begin() returns the two networks' input
// batches exactly as the reference would hand them to NN::infer (stale slots included), end() takes the outputs.
int kb_arena_begin(kb_arena* a, kb_arena_game* finished, int cap, int* n_finished, float* cur_obs, int* cur_n, float* cd_obs, int* cd_n) {
    KB_ARG(a && finished && n_finished && cap > 0 && cur_obs && cur_n && cd_obs && cd_n, "arena/finished/obs");
    kb_pool* p = a->pool;
    KB_BIND(p);
    KB_ARG(!a->pending, "kb_arena_begin without kb_arena_end");
    cudaStream_t st = main_stream();
    for (int w = 0; w < 2; ++w) {  // plane buffers exist even when no device network is involved (the route kernel writes both forms)
        if (a->ws[w].cap_boards < a->batch) {
            cudaFree(a->ws[w].P);
            a->ws[w].P = nullptr;
            KB_CUDA(cudaMalloc(&a->ws[w].P, act_bytes(a->batch, IN_SLABS)));
            KB_CUDA(cudaMemsetAsync(a->ws[w].P, 0, act_bytes(a->batch, IN_SLABS), st));
            a->ws[w].cap_boards = a->batch;
        }
    }
    int r = arena_begin_impl(a, finished, cap, n_finished);
    if (r) return r;
    *cur_n = (int)a->filed[0].size();
    *cd_n = (int)a->filed[1].size();
    if (*cur_n) KB_CUDA(cudaMemcpyAsync(cur_obs, a->obs_dev[0], sizeof(float) * KB_OBSIZE * (size_t)*cur_n, cudaMemcpyDeviceToHost, st));
    if (*cd_n) KB_CUDA(cudaMemcpyAsync(cd_obs, a->obs_dev[1], sizeof(float) * KB_OBSIZE * (size_t)*cd_n, cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    return KB_OK;
}
int kb_arena_end(kb_arena* a, const float* cur_policy, const float* cur_value, const float* cd_policy, const float* cd_value) {
    KB_ARG(a, "arena");
    kb_pool* p = a->pool;
    KB_BIND(p);
    KB_ARG(a->pending, "kb_arena_end without kb_arena_begin");
    cudaStream_t st = main_stream();
    const float* pol[2] = {cur_policy, cd_policy};
    const float* val[2] = {cur_value, cd_value};
    for (int w = 0; w < 2; ++w) {
        const size_t n = a->filed[w].size();
        if (!n) continue;
        KB_ARG(pol[w] && val[w], "policy / value rows of a network with filed leaves");
        KB_CUDA(cudaMemcpyAsync(a->pol_dev[w], pol[w], sizeof(float) * KB_PSIZE * n, cudaMemcpyHostToDevice, st));
        KB_CUDA(cudaMemcpyAsync(a->val_dev[w], val[w], sizeof(float) * n, cudaMemcpyHostToDevice, st));
    }
    return arena_expand_impl(a, 1);
}

}  // extern "C"
