// Device-resident game trees: PUCT select, expand, backup, move selection and re-rooting,
// one warp per tree; plus the warp-per-position encoders and legal-action kernels.
//
// Reference behaviour: kami/mcts.h (Node::backprop :35-42, select :186-255, expand :257-327,
// pick :137-184, push :113-135, snapshot :341-348, reset :331-339), kami/env.h (actions
// :398-423, terminal :288-385, observe :202-262) and the batch loop of
// kami/selfplay.cpp:113-200.
#include "tree.cuh"

#include <math.h>
#include <string.h>

#include <chrono>
#include <new>
#include <vector>

#include "common.cuh"
#include "layout.cuh"
#include "net.cuh"
#include "zobrist.h"

namespace kb {

int upload_tables() {
    u64 h[ZK_COUNT];
    make_zobrist(h, ZK_COUNT);
    KB_CUDA(cudaMemcpyToSymbol(d_zobrist, h, sizeof(h)));
    return KB_OK;
}

// ------------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------------
constexpr int COMPACT_PERIOD = 8;  // batched loops run the block-cooperative collector every COMPACT_PERIOD selects
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void raise(const PoolDev& P, int code) { atomicCAS(P.error, 0, code); }

__device__ __forceinline__ u64 splitmix(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double u01_from(u64 z) { return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

// Legal actions of p in the reference order (env.h:398-423): pseudo-legal generation by lane 0
// (emission order is sequential by nature), then SEE scoring, the make-move legality test, the
// stable descending sort and the action encoding spread over the 32 lanes.
__device__ int warp_legal_actions(const Pos& p, WarpScratch& s, u16* out) {
    const int lane = lane_id();
    if (lane == 0) gen_pseudo_legal(p, s.ml);
    __syncwarp();
    const int n = s.ml.n < MAX_MOVES ? s.ml.n : MAX_MOVES;
    u32 okbits = 0;
    for (int j = 0, i = lane; i < n; ++j, i += 32) {
        const u16 mv = s.ml.mv[i];
        s.score[i] = order_score(p, mv);
        Pos tmp;
        if (make_move<false>(p, mv, tmp)) okbits |= 1u << j;
    }
    __syncwarp();
    for (int j = 0, i = lane; i < n; ++j, i += 32) {
        const int sc = s.score[i];
        int rank = 0;
        for (int k = 0; k < n; ++k) {
            const int o = s.score[k];
            rank += (o > sc) || (o == sc && k < i);
        }
        s.sorted_mv[rank] = s.ml.mv[i];
        s.sorted_ok[rank] = (okbits >> j) & 1;
    }
    __syncwarp();
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const bool ok = i < n && s.sorted_ok[i];
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) out[base + __popc(m & ((1u << lane) - 1))] = (u16)encode_action(p, s.sorted_mv[i]);
        base += __popc(m);
    }
    __syncwarp();
    return base;
}

// ---- plane encoders (north_star kernel 1; Env::observe, env.h:202-262) ----------------------------------------------
// A position's 30 features per square are an 18-feature header that is the same on all 64 squares (ply bits 0-7, Q13;
// halfmove-clock bits 0-5; the four castle slots holding RAW mask values 1/2/4/8, Q2) and a one-hot over 12 piece
// planes at the POV-rotated square.  Both encoders build the header once per position and look up each square's
// hot feature once, then only assemble 16-byte vectors.

// hot feature (18..29) of POV square q, or -1 for an empty square
__device__ __forceinline__ int hot_feature(const Pos& p, int q) {
    const int sq = p.ctm == BLACK ? 63 - q : q;
    const int t = type_at(p, sq);
    return t < 0 ? -1 : 18 + (color_at(p, sq) != p.ctm ? 6 : 0) + t;
}
// value of castle slot f (14..17) as the raw mask the reference stores (env.h:219-235)
__device__ __forceinline__ int castle_slot(const Pos& p, int f) {
    const int wm = 1 << (f - 14);
    const int m = p.ctm == WHITE ? wm : (wm < 4 ? wm << 2 : wm >> 2);
    return p.castle & m;
}
// bf16 bit pattern of v in {0, 1, 2, 4, 8}
__device__ __forceinline__ u32 bf16_pow2(int v) { return v ? 0x3F80u + ((u32)(31 - __clz(v)) << 7) : 0u; }

// fp32 [64][30], the API-compat layout (7 680 B per position): the warp writes the position's 480 float4 in order,
// 15 per lane, fully coalesced.  hdr / hot: 18 floats and 64 bytes of per-warp shared scratch.
__device__ void warp_encode_f32(const Pos& p, float* dst, float* hdr, signed char* hot) {
    const int lane = lane_id();
    if (lane < 18) {
        const int f = lane;
        hdr[f] = f < 8 ? (float)((p.ply >> f) & 1) : f < 14 ? (float)((p.hmc >> (f - 8)) & 1) : (float)castle_slot(p, f);
    }
    hot[lane] = (signed char)hot_feature(p, lane);
    hot[lane + 32] = (signed char)hot_feature(p, lane + 32);
    __syncwarp();
    float4* out = reinterpret_cast<float4*>(dst);
#pragma unroll 5
    for (int j = 0; j < 15; ++j) {
        const int v = lane + 32 * j;
        int q = (4 * v) / NFEATURES, f = 4 * v - q * NFEATURES;
        float e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            e[i] = f < 18 ? hdr[f] : (hot[q] == f ? 1.0f : 0.0f);
            if (++f == NFEATURES) {
                f = 0;
                ++q;
            }
        }
        out[v] = make_float4(e[0], e[1], e[2], e[3]);
    }
    __syncwarp();
}
// bf16 planes in the swizzled tall-image layout the conv tower consumes (layout.cuh): the 30 features fill channels
// 0..29 of the input slab = chunks 0..3 (64 B) of each pixel line.  Chunks 0-1 (features 0..15) and the first word of
// chunk 2 (features 16, 17) are header-only; the rest is the one-hot.  Four lanes serve one pixel, lane & 3 = chunk:
// the swizzle (slot = chunk ^ (pixel & 7)) keeps chunks {0,1} and {2,3} in one aligned 32-byte sector each, so a warp
// store writes 16 whole sectors (8 pixels x 64 B) instead of 32 half sectors.
__device__ void warp_encode_tall(const Pos& p, int board, uint4* planes) {
    const int lane = lane_id();
    const int item = board / NB, slot = board - item * NB;
    uint4* base = planes + (size_t)item * IN_SLABS * SLAB_U4;
    const int c = lane & 3;
    // hot features of squares lane and lane + 32; pixel q's value is fetched from lane q & 31
    const int hot_lo = hot_feature(p, lane), hot_hi = hot_feature(p, lane + 32);
    uint4 hdr = make_uint4(0u, 0u, 0u, 0u);  // this lane's header words: chunk 0 / 1 entirely, word 0 of chunk 2
    if (c < 2) {
        const u32 bits = ((u32)(p.ply & 0xFF) | ((u32)(p.hmc & 0x3F) << 8)) >> (8 * c);
        u32 w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = ((bits >> (2 * k)) & 1u) * 0x3F80u | ((bits >> (2 * k + 1)) & 1u) * 0x3F800000u;
        if (c == 1) w[3] = bf16_pow2(castle_slot(p, 14)) | (bf16_pow2(castle_slot(p, 15)) << 16);
        hdr = make_uint4(w[0], w[1], w[2], w[3]);
    } else if (c == 2) {
        hdr.x = bf16_pow2(castle_slot(p, 16)) | (bf16_pow2(castle_slot(p, 17)) << 16);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int q = (lane >> 2) + 8 * j;
        const int lo = __shfl_sync(0xffffffffu, hot_lo, q & 31), hi = __shfl_sync(0xffffffffu, hot_hi, q & 31);
        const int hot = j < 4 ? lo : hi;
        const int px = tall_pixel(slot, q);
        uint4 v = hdr;
        if (c >= 2) {
            // one-hot over features 18..29 as bf16 pairs: pair k holds features 18+2k, 19+2k; chunk 2 = [hdr, pair 0, 1, 2],
            // chunk 3 = [pair 3, 4, 5, 0]
            const int r = hot - 18 - (c == 2 ? -2 : 6);  // position of the hot feature relative to this chunk's word 0
            v.x |= r == 0 ? 0x3F80u : r == 1 ? 0x3F800000u : 0u;
            v.y = r == 2 ? 0x3F80u : r == 3 ? 0x3F800000u : 0u;
            v.z = r == 4 ? 0x3F80u : r == 5 ? 0x3F800000u : 0u;
            v.w = r == 6 ? 0x3F80u : r == 7 ? 0x3F800000u : 0u;
        }
        base[chunk_u4(px, c)] = v;
    }
}

// ------------------------------------------------------------------------------------------
// tree primitives
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ Node* tree_nodes(const PoolDev& P, int t, u32 space) { return P.nodes + ((size_t)t * 2 + space) * P.cap; }
__device__ __forceinline__ u32* tree_meta(const PoolDev& P, int t, u32 space) { return P.meta + ((size_t)t * 2 + space) * P.cap; }
__device__ __forceinline__ float root_turn_of(const Pos& root) { return root.ctm == WHITE ? -1.0f : 1.0f; }  // mcts.h:88

// Node::backprop (mcts.h:35-42) over the recorded path; one lane per node.
__device__ void warp_backprop(const PoolDev& P, TreeCtl& c, Node* nodes, int depth, float value) {
    const int lane = lane_id();
    const float rt = root_turn_of(c.root_pos);
    for (int d = lane; d <= depth; d += 32) {
        Node* nd = nodes + c.path[d];
        const float turn = (d & 1) ? -rt : rt;
        nd->n += 1;
        nd->w = __fadd_rn(nd->w, __fadd_rn(0.5f, __fdiv_rn(__fmul_rn(value, turn), 2.0f)));
    }
    if (lane == 0) atomicAdd(&P.stats->path_nodes, (unsigned long long)(depth + 1));
    __syncwarp();
}

__device__ void tree_reset(const PoolDev& P, int t) {  // mcts.h:331-339
    TreeCtl& c = P.ctl[t];
    if (lane_id() == 0) {
        Node* nodes = tree_nodes(P, t, c.space);
        u32* meta = tree_meta(P, t, c.space);
        nodes[0] = Node{0, 0.0f, 0.0f, 0u};
        meta[0] = meta_pack(0xFFFF, 0);
        c.root = 0;
        c.alloc = 1;
        c.state = 0;
        c.depth = 0;
        c.n_hist = 0;
        c.leaf_nact = 0;
        c.traj_len = 0;
        c.want_compact = 0;
        start_position(c.root_pos);
    }
    __syncwarp();
}

// One MCTS::select call (mcts.h:186-255).  Returns true when a leaf is waiting for the network.
__device__ bool select_once(const PoolDev& P, int t, WarpScratch& s) {
    const int lane = lane_id();
    TreeCtl& c = P.ctl[t];
    if (c.state == 1) return true;  // target already set: the reference re-observes the same leaf
    Node* nodes = tree_nodes(P, t, c.space);
    u32* meta = tree_meta(P, t, c.space);
    Pos pos = c.root_pos;
    u32 cur = c.root;
    int depth = 0, nh = c.n_hist;
    float turn = root_turn_of(pos);
    if (lane == 0) c.path[0] = cur;
    unsigned long long scanned = 0;
    // One dependent memory round trip per level: a level's scan loads the children's node AND meta
    // words (two independent coalesced loads); the winner's copies are then shuffled out of the lane
    // that scored it, so the next level starts from registers instead of re-reading nodes[cur] /
    // meta[cur] / meta[child].
    const bool prof = P.dbg != nullptr;
    const long long pc0 = prof ? clock64() : 0;
    Node tn = nodes[cur];
    u32 m = meta[cur];
    for (;;) {
        const int k = (int)(m >> 16);
        if (k == 0) break;
        float cpuct = P.cfg.cpuct;
        if (P.cfg.scale_cpuct_by_actions) cpuct = __fdiv_rn(cpuct, (float)k);
        const double sq = __dsqrt_rn((double)tn.n);
        const float child_turn = -turn;
        const float fpu = __fmul_rn(P.cfg.fpu, child_turn);  // (Q6) sign follows the child's turn
        double best = -1000.0;
        int bi = 0x7fffffff;
        Node bch = Node{0, 0.0f, 0.0f, 0u};
        u32 bcm = 0;
        unsigned unvisited = 0;
        for (int i0 = 0; i0 < k; i0 += 32) {
            const int i = i0 + lane;
            bool nov = false;
            if (i < k) {
                const Node ch = nodes[tn.child0 + i];
                const u32 cm = meta[tn.child0 + i];
                nov = ch.n == 0;
                const float q = ch.n > 0 ? __fdiv_rn(ch.w, (float)ch.n) : fpu;
                const float pc = __fmul_rn(ch.p, cpuct);
                const double u = __ddiv_rn(__dmul_rn((double)pc, sq), (double)(ch.n + 1));
                const double uct = __dadd_rn((double)q, u);  // mcts.h:233
                if (uct > best) {
                    best = uct;
                    bi = i;
                    bch = ch;
                    bcm = cm;
                }
            }
            if (P.cfg.force_expand_unvisited && !unvisited) {
                const unsigned b = __ballot_sync(0xffffffffu, nov);
                if (b) unvisited = (unsigned)(i0 + __ffs(b));  // 1 + index of the first unvisited child
            }
        }
        scanned += k;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {  // first maximum wins (strict > in list order)
            const double ob = shfl_xor_d(best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ob > best || (ob == best && oi < bi)) {
                best = ob;
                bi = oi;
            }
        }
        if (bi == 0x7fffffff || depth >= MAX_DEPTH) {
            raise(P, bi == 0x7fffffff ? KB_ERR_STATE : KB_ERR_CAPACITY);
            return false;
        }
        u32 child;
        if (unvisited) {  // mcts.h:226-231: the forced child is nobody's arg-max, read it
            bi = (int)unvisited - 1;
            child = tn.child0 + (u32)bi;
            tn = nodes[child];
            m = meta[child];
        } else {          // lane (bi & 31) scored child bi and still holds its words
            const int src = bi & 31;
            child = tn.child0 + (u32)bi;
            tn.n = __shfl_sync(0xffffffffu, bch.n, src);
            tn.w = __shfl_sync(0xffffffffu, bch.w, src);
            tn.p = __shfl_sync(0xffffffffu, bch.p, src);
            tn.child0 = __shfl_sync(0xffffffffu, bch.child0, src);
            m = __shfl_sync(0xffffffffu, bcm, src);
        }
        const int action = (int)(m & 0xFFFF);
        // Children were created from legal actions, so the king-safety test is skipped; the check flag is only
        // needed at the leaf (movegen, Env::terminal) and is computed once after the descent.  The full key is
        // needed at every level (repetition history).
        Pos nx;
        make_move<false, true>(pos, decode_action(pos, action), nx);
        set_full_key(nx);
        if (lane == 0) {
            c.hist[nh] = pos.key;
            c.path[depth + 1] = child;
        }
        ++nh;
        ++depth;
        pos = nx;
        cur = child;
        turn = child_turn;
    }
    if (depth > 0) pos.check = in_check(pos);
    __syncwarp();
    const long long pc1 = prof ? clock64() : 0;
    if (lane == 0 && scanned) atomicAdd(&P.stats->children_scanned, scanned);
    // A node is one move sequence from the game start, so its terminal status never changes: the
    // first visit caches it in the (otherwise unused) child0 word of the childless node and the
    // repeat visits the reference absorbs inside its while loop (selfplay.cpp:133) skip movegen.
    const u32 tcache = tn.child0;
    int reason = 0;
    float value = 0.0f;
    int n = 0;
    if (tcache & 0x80000000u) {
        reason = 1;
        value = (tcache & 3u) == 1u ? 1.0f : (tcache & 3u) == 2u ? -1.0f : 0.0f;
    } else {
        reason = terminal_before_movegen(pos, c.hist, nh);
        if (!reason) {
            n = warp_legal_actions(pos, s, c.leaf_act);
            if (n == 0) value = no_moves_value(pos, &reason);
        }
        if (reason && lane == 0) nodes[cur].child0 = 0x80000000u | (value > 0.0f ? 1u : value < 0.0f ? 2u : 0u);
    }
    if (reason) {  // terminal leaf: back up the absolute value and report "no observation"
        warp_backprop(P, c, nodes, depth, value);
        if (lane == 0) atomicAdd(&P.stats->terminal_visits, 1ULL);
        return false;
    }
    if (lane == 0) {
        c.leaf_pos = pos;
        c.leaf_nact = n;
        c.depth = depth;
        c.state = 1;
        if (prof) {  // slots 3 / 4 of the select profile: descent and leaf (terminal test + legal actions) cycles
            P.dbg[(size_t)t * 8 + 3] = pc1 - pc0;
            P.dbg[(size_t)t * 8 + 4] = clock64() - pc1;
        }
    }
    __syncwarp();
    return true;
}

// MCTS::expand (mcts.h:257-327)
// compact: policy[i] is already the (unnormalised) prior of the i-th legal action (net_forward_legal_async)
__device__ void expand_once(const PoolDev& P, int t, const float* policy, float value, bool disable_bootstrap, WarpScratch& s, bool compact = false) {
    const int lane = lane_id();
    TreeCtl& c = P.ctl[t];
    if (c.state != 1) {
        raise(P, KB_ERR_STATE);
        return;
    }
    Node* nodes = tree_nodes(P, t, c.space);
    u32* meta = tree_meta(P, t, c.space);
    const int n = c.leaf_nact, depth = c.depth;
    const u32 leaf = c.path[depth];
    const u32 base = c.alloc;
    if (base + (u32)n > P.cap) {
        raise(P, KB_ERR_CAPACITY);
        return;
    }
    const float nw = P.cfg.noise_weight;
    for (int i = lane; i < n; i += 32) {
        s.fbuf[i] = compact ? policy[i] : policy[c.leaf_act[i]];
        // (Q9) Exp(1) noise at every expansion; the reference's is time-seeded so only the
        // distribution can be matched.  With noise_weight == 0 the term is exactly +0.
        float nz = 1.0f;
        if (nw != 0.0f) nz = (float)(-log(u01_from(splitmix(c.rng + 0x1000ULL * (u64)(i + 1)))));
        reinterpret_cast<float*>(s.score)[i] = nz;
    }
    __syncwarp();
    float ptotal = 0.0f, ntotal = 0.0f;
    if (lane == 0) {  // fp32 sums in list order (mcts.h:273-276, 282-287)
        for (int i = 0; i < n; ++i) {
            ptotal = __fadd_rn(ptotal, s.fbuf[i]);
            ntotal = __fadd_rn(ntotal, reinterpret_cast<float*>(s.score)[i]);
        }
    }
    ptotal = __shfl_sync(0xffffffffu, ptotal, 0);
    ntotal = __shfl_sync(0xffffffffu, ntotal, 0);
    const float keep = __fsub_rn(1.0f, nw);
    for (int i = lane; i < n; i += 32) {
        const float a = __fdiv_rn(__fmul_rn(keep, s.fbuf[i]), ptotal);
        const float b = __fmul_rn(nw, __fdiv_rn(reinterpret_cast<float*>(s.score)[i], ntotal));
        nodes[base + i] = Node{0, 0.0f, __fadd_rn(a, b), 0u};  // mcts.h:296
        meta[base + i] = meta_pack(c.leaf_act[i], 0);
    }
    const float rt = root_turn_of(c.root_pos);
    const float leaf_turn = (depth & 1) ? -rt : rt;
    value = __fmul_rn(value, leaf_turn);  // (Q8) mcts.h:313
    if (!disable_bootstrap && P.cfg.bootstrap_weight > 0.0f) {
        // Env::bootstrap_value (env.h:476-484): the eval's 20 guard terms are spread over the lanes and summed
        // (integer sums: same result in any order), the rest is cheap and uniform
        int gm = 0, ge = 0;
        if (lane < EVAL_GUARD_TERMS) eval_guard_term(c.leaf_pos, lane, gm, ge);
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            gm += __shfl_xor_sync(0xffffffffu, gm, off);
            ge += __shfl_xor_sync(0xffffffffu, ge, off);
        }
        float bs = __fdiv_rn((float)static_eval(c.leaf_pos, &gm, &ge), P.cfg.bootstrap_window);
        bs = bs < 1.0f ? bs : 1.0f;
        bs = bs > -1.0f ? bs : -1.0f;
        const float a = __fmul_rn(__fsub_rn(1.0f, P.cfg.bootstrap_weight), value);
        const float b = __fmul_rn(__fmul_rn(P.cfg.bootstrap_weight, bs), P.cfg.bootstrap_amp);
        value = __fadd_rn(a, b);  // mcts.h:315-316
    }
    __syncwarp();
    if (lane == 0) {
        nodes[leaf].child0 = base;
        meta[leaf] = (meta[leaf] & 0xFFFFu) | ((u32)n << 16);
        c.alloc = base + (u32)n;
        c.rng += 1;
        atomicAdd(&P.stats->children_created, (unsigned long long)n);
        atomicAdd(&P.stats->evals, 1ULL);
    }
    __syncwarp();
    warp_backprop(P, c, nodes, depth, value);
    if (lane == 0) c.state = 0;
    __syncwarp();
}

// Copying collector: moves the subtree under `keep` into the other semi-space in breadth-first
// order (children stay contiguous and ordered) and makes it the root.
__device__ void compact_into_other_space(const PoolDev& P, int t, u32 keep) {
    const int lane = lane_id();
    TreeCtl& c = P.ctl[t];
    const Node* sn = tree_nodes(P, t, c.space);
    const u32* sm = tree_meta(P, t, c.space);
    Node* dn = tree_nodes(P, t, c.space ^ 1);
    u32* dm = tree_meta(P, t, c.space ^ 1);
    if (lane == 0) {
        dn[0] = sn[keep];
        dm[0] = sm[keep];
    }
    __syncwarp();
    u32 tail = 1, head = 0;
    while (head < tail) {  // [head, tail) = copied nodes whose children still live in the old space
        const u32 wave = tail - head < 32u ? tail - head : 32u;
        const u32 i = head + lane;
        u32 k = 0, oc = 0;
        if ((u32)lane < wave) {
            k = dm[i] >> 16;
            oc = dn[i].child0;
        }
        u32 incl = k;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u32 v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        const u32 total = __shfl_sync(0xffffffffu, incl, 31);
        const u32 nc0 = tail + incl - k;
        if (k) dn[i].child0 = nc0;
        unsigned todo = __ballot_sync(0xffffffffu, k > 0);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const u32 kk = __shfl_sync(0xffffffffu, k, l);
            const u32 so = __shfl_sync(0xffffffffu, oc, l);
            const u32 d0 = __shfl_sync(0xffffffffu, nc0, l);
            for (u32 j = lane; j < kk; j += 32) {
                dn[d0 + j] = sn[so + j];
                dm[d0 + j] = sm[so + j];
            }
        }
        tail += total;
        head += wave;
        __syncwarp();
    }
    if (lane == 0) {
        c.space ^= 1;
        c.root = 0;
        c.alloc = tail;
    }
    __syncwarp();
}

// MCTS::push (mcts.h:113-135): re-root on the child that carries `action`, keeping its subtree.
__device__ bool push_once(const PoolDev& P, int t, int action) {
    const int lane = lane_id();
    TreeCtl& c = P.ctl[t];
    Node* nodes = tree_nodes(P, t, c.space);
    u32* meta = tree_meta(P, t, c.space);
    const u32 root = c.root;
    const int k = (int)(meta[root] >> 16);
    const u32 c0 = nodes[root].child0;
    int found = -1;
    for (int i0 = 0; i0 < k && found < 0; i0 += 32) {
        const int i = i0 + lane;
        const bool hit = i < k && (int)(meta[c0 + i] & 0xFFFF) == action;
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (b) found = i0 + __ffs(b) - 1;
    }
    if (found < 0) {
        raise(P, KB_ERR_NO_CHILD);
        return false;
    }
    const u32 keep = c0 + (u32)found;
    Pos nx;
    make_move<true>(c.root_pos, decode_action(c.root_pos, action), nx);
    const u64 oldkey = c.root_pos.key;
    __syncwarp();
    int nh = c.n_hist;
    if (nh == HIST_GAME) {  // slide the key window; only the last hmc < 50 keys can ever match
        u64 a = c.hist[lane + 1], b = lane + 33 < HIST_GAME ? c.hist[lane + 33] : 0;
        __syncwarp();
        c.hist[lane] = a;
        if (lane + 32 < HIST_GAME - 1) c.hist[lane + 32] = b;
        nh = HIST_GAME - 1;
        __syncwarp();
    }
    if (lane == 0) {
        c.hist[nh] = oldkey;
        c.n_hist = nh + 1;
        c.root_pos = nx;
        c.state = 0;
        c.depth = 0;
    }
    __syncwarp();
    // Re-rooting is free (the root index moves); the copying collector only runs when the arena
    // could not hold another move's worth of expansions.
    // The reserve must also cover what the batched loops allocate between the flag and the deferred collector's next run:
    // COMPACT_PERIOD selects with up to MAX_MOVES children each (tiny node budgets used to overflow here: 6 x 64 = 384
    // nodes were less than eight expansions of ~50 children).
    // (single-tree protocol, selfplay_nodes == 0: the caller's visit budget is unknown, half the arena is kept free)
    u32 reserve = P.cfg.selfplay_nodes > 0 ? (u32)P.cfg.selfplay_nodes * 64u : P.cap / 2;
    if (reserve < (u32)(COMPACT_PERIOD + 2) * MAX_MOVES) reserve = (u32)(COMPACT_PERIOD + 2) * MAX_MOVES;
    if (reserve > P.cap / 2) reserve = P.cap / 2;
    if (c.alloc + reserve > P.cap && !P.defer_compact) compact_into_other_space(P, t, keep);
    else if (lane == 0) {
        c.root = keep;
        if (c.alloc + reserve > P.cap) c.want_compact = 1;
    }
    __syncwarp();
    return true;
}

// MCTS::pick (mcts.h:137-184); u01 stands in for rand()/RAND_MAX.  The pow() calls run one per
// lane; the double-precision sums stay sequential in list order like the reference's loops.
__device__ int pick_once(const PoolDev& P, int t, float alpha, double u01, WarpScratch& s) {
    TreeCtl& c = P.ctl[t];
    const Node* nodes = tree_nodes(P, t, c.space);
    const u32* meta = tree_meta(P, t, c.space);
    const int k = (int)(meta[c.root] >> 16);
    const u32 c0 = nodes[c.root].child0;
    const int lane = lane_id();
    if (k == 0) return -2;
    int result = -1;
    if (alpha < 0.1f) {
        int bn = 0, bi = 0x7fffffff;
        for (int i = lane; i < k; i += 32) {
            const int n = nodes[c0 + i].n;
            if (n > bn) {  // strict >: the first maximum wins
                bn = n;
                bi = i;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int on = __shfl_xor_sync(0xffffffffu, bn, off), oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (on > bn || (on == bn && oi < bi)) {
                bn = on;
                bi = oi;
            }
        }
        if (bn > 0) result = (int)(meta[c0 + bi] & 0xFFFF);  // (Q15) -1 when every child has n == 0
        return result;
    }
    double* dist = reinterpret_cast<double*>(s.score);  // 128 doubles over score[] + sorted_mv[] (1 KB)
    const double e = (double)__fdiv_rn(1.0f, alpha);
    for (int i = lane; i < k && i < MAX_MOVES; i += 32) dist[i] = pow((double)nodes[c0 + i].n, e);
    __syncwarp();
    if (lane == 0) {
        double len = 0.0;
        for (int i = 0; i < k; ++i) len += dist[i];
        double ind = u01;
        int pi = k - 1;
        for (int i = 0; i < k; ++i) {
            ind -= dist[i] / len;
            if (ind <= 0.0) {
                pi = i;
                break;
            }
        }
        result = (int)(meta[c0 + pi] & 0xFFFF);
    }
    __syncwarp();
    return __shfl_sync(0xffffffffu, result, 0);
}

// Full Env::terminal (env.h:288-391) on the root position of a tree.
__device__ bool root_terminal(const PoolDev& P, int t, WarpScratch& s, float* value) {
    TreeCtl& c = P.ctl[t];
    const Pos pos = c.root_pos;
    int reason = terminal_before_movegen(pos, c.hist, c.n_hist);
    *value = 0.0f;
    if (!reason) {
        const int n = warp_legal_actions(pos, s, s.tmp_out);  // output discarded
        if (n == 0) *value = no_moves_value(pos, &reason);
    }
    return reason != 0;
}

// The "budget reached" branch of Selfplay::inference_main (selfplay.cpp:136-192): record the
// sample, choose a move with the temperature schedule, re-root, and recycle finished games.
__device__ void play_move(const PoolDev& P, int t, WarpScratch& s) {
    const int lane = lane_id();
    TreeCtl& c = P.ctl[t];
    const Node* nodes = tree_nodes(P, t, c.space);
    const u32* meta = tree_meta(P, t, c.space);
    const u32 root = c.root;
    const int k = (int)(meta[root] >> 16);
    const u32 c0 = nodes[root].child0;
    const float turn = c.root_pos.ctm == WHITE ? 1.0f : -1.0f;
    if (c.traj_len < P.traj_cap) {
        TrajSample* ts = P.traj + (size_t)t * P.traj_cap + c.traj_len;
        if (lane == 0) {
            ts->pos = c.root_pos;
            ts->pov = -turn;  // selfplay.cpp:148
            ts->root_n = nodes[root].n;
            ts->nchild = k;
        }
        for (int i = lane; i < k && i < TRAJ_MAX_CHILD; i += 32)
            ts->entry[i] = ((meta[c0 + i] & 0xFFFFu) << 16) | ((u32)nodes[c0 + i].n & 0xFFFFu);
    } else {
        raise(P, KB_ERR_CAPACITY);
    }
    __syncwarp();
    const int ply = c.root_pos.ply;
    float alpha = P.cfg.alpha_final;  // selfplay.cpp:153-156
    if (ply < P.cfg.alpha_cutoff) alpha = (float)(pow((double)P.cfg.alpha_decay, (double)ply) * (double)P.cfg.alpha_initial);
    const double u = u01_from(splitmix(c.rng ^ 0xA5A5A5A5DEADBEEFULL));
    const int action = pick_once(P, t, alpha, u, s);
    if (action < 0) {
        raise(P, KB_ERR_NO_CHILD);
        return;
    }
    if (lane == 0) {
        if (c.traj_len < P.traj_cap) P.traj[(size_t)t * P.traj_cap + c.traj_len].action = action;
        c.traj_len += 1;
        c.rng += 1;
        c.moves += 1;
        atomicAdd(&P.stats->moves, 1ULL);
    }
    __syncwarp();
    if (!push_once(P, t, action)) return;
    float value;
    if (root_terminal(P, t, s, &value)) {  // selfplay.cpp:165-188
        const int len = c.traj_len;
        unsigned long long slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(P.replay_head, (unsigned long long)len);
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        {   // one flat copy of len samples (38 x 16 B each) with eight loads in flight per lane
            constexpr int CH = (int)(sizeof(TrajSample) / 16);
            const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(P.traj + (size_t)t * P.traj_cap);
            const int total = len * CH;
            for (int base = 0; base < total; base += 32 * 8) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + u * 32 + lane;
                    if (e < total) v[u] = __ldg(s4 + e);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + u * 32 + lane;
                    if (e < total) {
                        const int i = e / CH, j = e - i * CH;
                        ReplaySample* dst = P.replay + (slot0 + i) % (unsigned long long)P.replay_cap;
                        reinterpret_cast<uint4*>(&dst->s)[j] = v[u];
                    }
                }
            }
            for (int i = lane; i < len; i += 32) {
                const TrajSample* src = P.traj + (size_t)t * P.traj_cap + i;
                ReplaySample* dst = P.replay + (slot0 + i) % (unsigned long long)P.replay_cap;
                dst->z = value == 0.0f ? P.cfg.draw_value : __fmul_rn(src->pov, value);
            }
        }
        // a host asked for the next finished game's moves: the first warp to finish one hands them over
        int take = 0;
        if (lane == 0 && *(volatile int*)P.game == 1) take = atomicCAS(P.game, 1, 2) == 1;
        if (__shfl_sync(0xffffffffu, take, 0)) {
            for (int i = lane; i < len; i += 32) P.game[2 + i] = P.traj[(size_t)t * P.traj_cap + i].action;
            __syncwarp();
            if (lane == 0) {
                P.game[1] = len;
                __threadfence();
                atomicExch(P.game, 3);
            }
        }
        if (lane == 0) {
            c.games += 1;
            atomicAdd(&P.stats->games, 1ULL);
            atomicAdd(&P.stats->samples, (unsigned long long)len);
        }
        __syncwarp();
        tree_reset(P, t);
    }
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_pool_reset(PoolDev P) {
    const int t = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (t >= P.n_trees) return;
    if (lane_id() == 0) {
        P.ctl[t].space = 0;
        P.ctl[t].rng = splitmix(P.cfg.seed + 0x632BE59BD9B4E019ULL * (u64)(t + 1));
        P.ctl[t].games = 0;
        P.ctl[t].moves = 0;
    }
    __syncwarp();
    tree_reset(P, t);
}

// Block-cooperative copying collector for the batched loops (one block per tree, launched every
// few steps at a safe point: no leaf pending).  Same breadth-first order as the warp version, so
// the resulting arena is identical; 256 parents per wave, children copied as one flat range.
__global__ void __launch_bounds__(256) k_pool_compact(PoolDev P) {
    const int t = P.tree0 + blockIdx.x;
    TreeCtl& c = P.ctl[t];
    if (!c.want_compact) return;
    __shared__ u32 s_incl[256], s_oc[256], s_warp[8];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const Node* __restrict__ sn = tree_nodes(P, t, c.space);
    const u32* __restrict__ sm = tree_meta(P, t, c.space);
    Node* dn = tree_nodes(P, t, c.space ^ 1);
    u32* dm = tree_meta(P, t, c.space ^ 1);
    const u32 keep = c.root;
    if (tid == 0) {
        dn[0] = sn[keep];
        dm[0] = sm[keep];
    }
    __syncthreads();
    u32 tail = 1, head = 0;
    while (head < tail) {
        const u32 wave = tail - head < 256u ? tail - head : 256u;
        const u32 i = head + tid;
        u32 k = 0, oc = 0;
        if ((u32)tid < wave) {
            k = dm[i] >> 16;
            oc = dn[i].child0;
        }
        u32 incl = k;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u32 v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        u32 before = 0;
        for (int j = 0; j < w; ++j) before += s_warp[j];
        incl += before;
        s_incl[tid] = incl;
        s_oc[tid] = oc;
        if (k) dn[i].child0 = tail + incl - k;
        __syncthreads();
        const u32 total = s_incl[255];
        for (u32 base = 0; base < total; base += 256 * 4) {
            u32 src[4];
            Node nv[4];
            u32 mv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const u32 j = base + u * 256 + tid;
                if (j < total) {
                    int lo = 0;  // first parent whose inclusive child count exceeds j
#pragma unroll
                    for (int step = 128; step > 0; step >>= 1)
                        if (s_incl[lo + step - 1] <= j) lo += step;
                    src[u] = s_oc[lo] + j - (lo ? s_incl[lo - 1] : 0u);
                    nv[u] = sn[src[u]];
                    mv[u] = sm[src[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const u32 j = base + u * 256 + tid;
                if (j < total) {
                    dn[tail + j] = nv[u];
                    dm[tail + j] = mv[u];
                }
            }
        }
        __syncthreads();
        tail += total;
        head += wave;
    }
    if (tid == 0) {
        c.space ^= 1;
        c.root = 0;
        c.alloc = tail;
        c.want_compact = 0;
    }
}

// Batched select: the inner loop of selfplay.cpp:116-193 for every tree at once.
// planes != nullptr: also writes the leaf's bf16 input planes (kernel 1 fused behind select).
__device__ __forceinline__ void pool_select_tree(const PoolDev& P, int t, WarpScratch& s, uint4* planes, Pos* leaf_out) {
    TreeCtl& c = P.ctl[t];
    const bool prof = P.dbg != nullptr;
    long long t_move = 0, t_sel = 0, t_enc = 0, n_sel = 0, n_move = 0, c0 = 0, t0 = prof ? clock64() : 0;
    for (int guard = 0; guard < (1 << 20); ++guard) {
        if (*P.error) return;
        if (P.cfg.selfplay_nodes > 0) {
            const int rn = tree_nodes(P, t, c.space)[c.root].n;
            if (rn >= P.cfg.selfplay_nodes) {
                if (prof) c0 = clock64();
                play_move(P, t, s);
                if (prof) { t_move += clock64() - c0; ++n_move; }
                continue;
            }
        }
        if (prof) c0 = clock64();
        const bool got = select_once(P, t, s);
        if (prof) { t_sel += clock64() - c0; ++n_sel; }
        if (got) break;
    }
    if (prof) c0 = clock64();
    if (planes) warp_encode_tall(c.leaf_pos, t, planes);
    if (leaf_out && lane_id() == 0) leaf_out[t] = c.leaf_pos;
    if (prof && lane_id() == 0) {
        t_enc = clock64() - c0;
        long long* d = P.dbg + (size_t)t * 8;
        d[0] = clock64() - t0; d[1] = t_sel; d[2] = n_sel; if (n_move) { d[3] = -t_move; d[4] = -n_move; } d[5] = t_enc; d[6] = c.depth; d[7] = c.leaf_nact;
    }
}
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_pool_select(PoolDev P, uint4* planes, Pos* leaf_out) {
    __shared__ WarpScratch scratch[WARPS_PER_BLOCK];
    const int w = threadIdx.x >> 5;
    const int t = P.tree0 + blockIdx.x * WARPS_PER_BLOCK + w;
    pdl_launch_dependents();  // the tower's CTAs may move in (and set up) as this grid's blocks drain
    pdl_wait();
    if (t >= P.tree_hi) return;
    pool_select_tree(P, t, scratch[w], planes, leaf_out);
}

// Batched expand + backup.  value_stride/value_mode implement NN::infer's value indexing (Q1).
__device__ __forceinline__ void pool_expand_tree(const PoolDev& P, int t, WarpScratch& s, const float* policy, const float* value, int value_is_256,
                                                 int disable_bootstrap, int compact) {
    float v;
    if (!value_is_256) v = value[t];
    else v = P.cfg.value_index_mode == 0 ? value[t] /* vh.flat[t], nn.cpp:186 */ : value[(size_t)t * 256];
    expand_once(P, t, policy + (size_t)t * (compact ? 128 : PSIZE), v, disable_bootstrap != 0, s, compact != 0);
}
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_pool_expand(PoolDev P, const float* policy, const float* value, int value_is_256, int disable_bootstrap,
                                                                      int compact) {
    __shared__ WarpScratch scratch[WARPS_PER_BLOCK];
    const int w = threadIdx.x >> 5;
    const int t = P.tree0 + blockIdx.x * WARPS_PER_BLOCK + w;
    pdl_launch_dependents();
    pdl_wait();
    if (t >= P.tree_hi) return;
    if (*P.error) return;
    pool_expand_tree(P, t, scratch[w], policy, value, value_is_256, disable_bootstrap, compact);
}
// expand of iteration i and select of iteration i + 1 in one launch (kb_pool_step): a tree's warp goes straight from its
// backup to its next descent, so the slowest expand no longer holds back every select (and one launch boundary goes away).
// Every value the trees read (Q1: tree t reads vh.flat[t], another board's output) was written by the tower before the
// launch; a warp only writes its own tree, its own plane rows and its own leaf list.
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_pool_expand_select(PoolDev P, const float* policy, const float* value, int value_is_256,
                                                                             int disable_bootstrap, int compact, uint4* planes, Pos* leaf_out) {
    __shared__ WarpScratch scratch[WARPS_PER_BLOCK];
    const int w = threadIdx.x >> 5;
    const int t = P.tree0 + blockIdx.x * WARPS_PER_BLOCK + w;
    pdl_launch_dependents();
    pdl_wait();
    if (t >= P.tree_hi) return;
    if (*P.error) return;
    pool_expand_tree(P, t, scratch[w], policy, value, value_is_256, disable_bootstrap, compact);
    __syncwarp();
    pool_select_tree(P, t, scratch[w], planes, leaf_out);
}

// Single-tree entry points (the kami::MCTS call protocol)
__global__ void k_tree_select(PoolDev P, int t, float* obs, int* need_eval) {
    __shared__ WarpScratch s;
    const bool r = select_once(P, t, s);
    if (r) warp_encode_f32(P.ctl[t].leaf_pos, obs, s.fbuf, reinterpret_cast<signed char*>(s.sorted_ok));
    if (lane_id() == 0) *need_eval = r ? 1 : 0;
}
__global__ void k_tree_expand(PoolDev P, int t, const float* policy, float value, int disable_bootstrap) {
    __shared__ WarpScratch s;
    expand_once(P, t, policy, value, disable_bootstrap != 0, s);
}
__global__ void k_tree_pick(PoolDev P, int t, float alpha, double u01, int* out) {
    __shared__ WarpScratch s;
    const int a = pick_once(P, t, alpha, u01, s);
    if (lane_id() == 0) *out = a;
}
__global__ void k_tree_push(PoolDev P, int t, int action) { push_once(P, t, action); }
__global__ void k_tree_reset(PoolDev P, int t) { tree_reset(P, t); }
// actions along the selected path (root -> pending leaf); out[0] = depth (0 when no leaf is pending)
__global__ void k_tree_leaf_path(PoolDev P, int t, int* out, int cap) {
    const TreeCtl& c = P.ctl[t];
    const u32* meta = tree_meta(P, t, c.space);
    const int depth = c.state == 1 ? c.depth : 0;
    if (threadIdx.x == 0) out[0] = depth;
    for (int d = threadIdx.x; d < depth && d + 1 < cap; d += 32) out[1 + d] = (int)(meta[c.path[d + 1]] & 0xFFFF);
}
__global__ void k_tree_snapshot(PoolDev P, int t, float* ps) {  // mcts.h:341-348
    const TreeCtl& c = P.ctl[t];
    const Node* nodes = tree_nodes(P, t, c.space);
    const u32* meta = tree_meta(P, t, c.space);
    for (int i = threadIdx.x; i < PSIZE; i += blockDim.x) ps[i] = 0.0f;
    __syncthreads();
    const int k = (int)(meta[c.root] >> 16);
    const u32 c0 = nodes[c.root].child0;
    const float den = (float)(nodes[c.root].n - 1);  // (Q15)
    for (int i = threadIdx.x; i < k; i += blockDim.x) ps[meta[c0 + i] & 0xFFFF] = __fdiv_rn((float)nodes[c0 + i].n, den);
}
struct RootInfo {
    int n;
    float w;
    int k;
    int pad;
};
__global__ void k_tree_root(PoolDev P, int t, RootInfo* info, int* action, int* n, float* w, float* p, int cap) {
    const TreeCtl& c = P.ctl[t];
    const Node* nodes = tree_nodes(P, t, c.space);
    const u32* meta = tree_meta(P, t, c.space);
    const int k = (int)(meta[c.root] >> 16);
    const u32 c0 = nodes[c.root].child0;
    if (threadIdx.x == 0) *info = RootInfo{nodes[c.root].n, nodes[c.root].w, k, 0};
    for (int i = threadIdx.x; i < k && i < cap; i += blockDim.x) {
        const Node ch = nodes[c0 + i];
        action[i] = (int)(meta[c0 + i] & 0xFFFF);
        n[i] = ch.n;
        w[i] = ch.w;
        p[i] = ch.p;
    }
}
// Pre-order digest of the whole tree, same mixing as oracle/kami_oracle.c:digest (test hook).
__global__ void k_tree_digest(PoolDev P, int t, u64* out) {
    const TreeCtl& c = P.ctl[t];
    const Node* nodes = tree_nodes(P, t, c.space);
    const u32* meta = tree_meta(P, t, c.space);
    u32 stack_node[MAX_DEPTH + 2];
    int stack_next[MAX_DEPTH + 2];
    int sp = 0;
    u64 h = 0, cnt = 0;
    stack_node[0] = c.root;
    stack_next[0] = -1;
    while (sp >= 0) {
        const u32 nd = stack_node[sp];
        const int k = (int)(meta[nd] >> 16);
        if (stack_next[sp] < 0) {
            const u32 a16 = meta[nd] & 0xFFFF;
            const u64 v[5] = {a16 == 0xFFFF ? 0xFFFFFFFFULL : (u64)a16, (u64)(u32)nodes[nd].n, (u64)__float_as_uint(nodes[nd].w),
                              (u64)__float_as_uint(nodes[nd].p), (u64)k};
            for (int i = 0; i < 5; ++i) h ^= v[i] + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
            ++cnt;
            stack_next[sp] = 0;
        }
        if (stack_next[sp] < k && sp < MAX_DEPTH) {
            const u32 ch = nodes[nd].child0 + (u32)stack_next[sp];
            stack_next[sp] += 1;
            ++sp;
            stack_node[sp] = ch;
            stack_next[sp] = -1;
        } else
            --sp;
    }
    out[0] = h;
    out[1] = cnt;
}

// Batched position kernels -------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_encode_f32(const Pos* pos, int n, float* obs) {
    __shared__ float hdr[4][20];
    __shared__ signed char hot[4][64];
    const int w = threadIdx.x >> 5;
    const int b = blockIdx.x * 4 + w;
    if (b >= n) return;
    const Pos p = pos[b];
    warp_encode_f32(p, obs + (size_t)b * 64 * NFEATURES, hdr[w], hot[w]);
}
__global__ void __launch_bounds__(128) k_encode_tall(const Pos* pos, int n, uint4* planes) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= n) return;
    const Pos p = pos[b];
    warp_encode_tall(p, b, planes);
}
// fp32 [n][64][30] observations (NN::infer's input) -> bf16 swizzled tall-image input slab
__global__ void __launch_bounds__(128) k_obs_to_tall(const float* obs, int n, uint4* planes) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= n) return;
    const int lane = lane_id();
    const int item = b / NB, slot = b - item * NB;
    uint4* base = planes + (size_t)item * IN_SLABS * SLAB_U4;
    const float* src = obs + (size_t)b * 64 * NFEATURES;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int q = lane + 32 * h;
        const int px = tall_pixel(slot, q);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            u32 wv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int f0 = c * 8 + 2 * k, f1 = f0 + 1;
                const float v0 = f0 < NFEATURES ? src[q * NFEATURES + f0] : 0.0f;
                const float v1 = f1 < NFEATURES ? src[q * NFEATURES + f1] : 0.0f;
                const __nv_bfloat162 bb = __floats2bfloat162_rn(v0, v1);
                wv[k] = *reinterpret_cast<const u32*>(&bb);
            }
            base[chunk_u4(px, c)] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
    }
}
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_legal_actions(const Pos* pos, int n, int* actions, int* counts) {
    __shared__ WarpScratch scratch[WARPS_PER_BLOCK];
    __shared__ u16 outbuf[WARPS_PER_BLOCK][MAX_MOVES];
    const int w = threadIdx.x >> 5;
    const int b = blockIdx.x * WARPS_PER_BLOCK + w;
    if (b >= n) return;
    const Pos p = pos[b];
    const int k = warp_legal_actions(p, scratch[w], outbuf[w]);
    for (int i = lane_id(); i < MAX_MOVES; i += 32) actions[(size_t)b * MAX_MOVES + i] = i < k ? (int)outbuf[w][i] : -1;
    if (lane_id() == 0) counts[b] = k;
}
__global__ void k_apply_actions(Pos* pos, int n, const int* actions) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const Pos p = pos[b];
    Pos nx;
    make_move<true>(p, decode_action(p, actions[b]), nx);
    pos[b] = nx;
}
__global__ void k_static_eval(const Pos* pos, int n, int* out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    out[b] = static_eval(pos[b]);
}

// Env kernels: a device-resident stack of positions (Env::push / pop / actions / terminal)
struct EnvDev {
    Pos* stack;
    u64* keys;  // keys[i] = stack[i].key, contiguous for the repetition scan
    int* count;
    int cap;
};
__global__ void k_env_reset(EnvDev e) {
    if (threadIdx.x == 0) {
        start_position(e.stack[0]);
        e.keys[0] = e.stack[0].key;
        *e.count = 1;
    }
}
__global__ void k_env_push(EnvDev e, int action, int* status) {
    if (threadIdx.x) return;
    const int n = *e.count;
    if (n >= e.cap) {
        *status = KB_ERR_CAPACITY;
        return;
    }
    const Pos p = e.stack[n - 1];
    Pos nx;
    make_move<true>(p, decode_action(p, action), nx);
    e.stack[n] = nx;
    e.keys[n] = nx.key;
    *e.count = n + 1;
    *status = 0;
}
struct EnvQuery {
    int n_actions;
    int terminal;
    int reason;
    float value;
    float bootstrap;
    int code;
    int ply;
    int pad;
};
// op bit 0: actions, 1: terminal, 2: observe, 3: bootstrap, 4: encode(arg), 5: decode(arg)
__global__ void k_env_query(EnvDev e, int ops, int arg, float farg, EnvQuery* q, int* actions, float* obs) {
    __shared__ WarpScratch s;
    __shared__ u16 outbuf[MAX_MOVES];
    const int n = *e.count;
    const Pos p = e.stack[n - 1];
    const int lane = lane_id();
    if (lane == 0) q->ply = n - 1;
    int nact = -1;
    if (ops & 3) {
        int reason = 0;
        float value = 0.0f;
        if (ops & 2) reason = terminal_before_movegen(p, e.keys, n - 1);
        if (!reason) {
            nact = warp_legal_actions(p, s, outbuf);
            if ((ops & 2) && nact == 0) value = no_moves_value(p, &reason);
        }
        if (lane == 0) {
            q->terminal = reason != 0;
            q->reason = reason;
            q->value = value;
            q->n_actions = nact;
        }
        if ((ops & 1) && nact >= 0)
            for (int i = lane; i < nact; i += 32) actions[i] = outbuf[i];
    }
    if (ops & 4) warp_encode_f32(p, obs, s.fbuf, reinterpret_cast<signed char*>(s.sorted_ok));
    if (lane == 0) {
        if (ops & 8) q->bootstrap = bootstrap_value(p, farg);
        if (ops & 16) q->code = encode_action(p, (u16)arg);
        if (ops & 32) q->code = decode_action(p, arg);
    }
}

}  // namespace kb

// ==========================================================================================
// C ABI: Env, batched position kernels, pools
// ==========================================================================================
using namespace kb;

struct kb_env {
    EnvDev d;
    EnvQuery* q_dev;
    int* act_dev;
    float* obs_dev;
    int* status_dev;
};

extern "C" {

int kb_env_create(kb_env** out) {
    KB_REQUIRE_INIT();
    KB_ARG(out, "out");
    kb_env* e = new (std::nothrow) kb_env();
    if (!e) return KB_ERR_ARG;
    e->d.cap = 2048;  // position.h:20 NC_MAX_PLY
    KB_CUDA(cudaMalloc(&e->d.stack, sizeof(Pos) * e->d.cap));
    KB_CUDA(cudaMalloc(&e->d.keys, sizeof(u64) * e->d.cap));
    KB_CUDA(cudaMalloc(&e->d.count, sizeof(int)));
    KB_CUDA(cudaMalloc(&e->q_dev, sizeof(EnvQuery)));
    KB_CUDA(cudaMalloc(&e->act_dev, sizeof(int) * MAX_MOVES));
    KB_CUDA(cudaMalloc(&e->obs_dev, sizeof(float) * KB_OBSIZE));
    KB_CUDA(cudaMalloc(&e->status_dev, sizeof(int)));
    *out = e;
    return kb_env_reset(e);
}
int kb_env_destroy(kb_env* e) {
    if (!e) return KB_OK;
    cudaFree(e->d.stack);
    cudaFree(e->d.keys);
    cudaFree(e->d.count);
    cudaFree(e->q_dev);
    cudaFree(e->act_dev);
    cudaFree(e->obs_dev);
    cudaFree(e->status_dev);
    delete e;
    return KB_OK;
}
int kb_env_reset(kb_env* e) {
    KB_ARG(e, "env");
    k_env_reset<<<1, 32, 0, main_stream()>>>(e->d);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}
static int env_query(kb_env* e, int ops, int arg, float farg, EnvQuery* q) {
    k_env_query<<<1, 32, 0, main_stream()>>>(e->d, ops, arg, farg, e->q_dev, e->act_dev, e->obs_dev);
    KB_CUDA(cudaGetLastError());
    KB_CUDA(cudaMemcpyAsync(q, e->q_dev, sizeof(EnvQuery), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_env_ply(kb_env* e, int* ply) {
    KB_ARG(e && ply, "env/ply");
    EnvQuery q;
    int r = env_query(e, 0, 0, 0.0f, &q);
    if (r) return r;
    *ply = q.ply;
    return KB_OK;
}
int kb_env_push(kb_env* e, int action) {
    KB_ARG(e, "env");
    KB_ARG(action >= 0 && action < KB_PSIZE, "action out of range");
    k_env_push<<<1, 32, 0, main_stream()>>>(e->d, action, e->status_dev);
    KB_CUDA(cudaGetLastError());
    int st = 0;
    KB_CUDA(cudaMemcpyAsync(&st, e->status_dev, sizeof(int), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    if (st) {
        set_error("env stack capacity exceeded");
        return st;
    }
    return KB_OK;
}
int kb_env_pop(kb_env* e) {
    KB_ARG(e, "env");
    int n = 0;
    KB_CUDA(cudaMemcpyAsync(&n, e->d.count, sizeof(int), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    if (n <= 1) {
        set_error("pop on an empty history");
        return KB_ERR_STATE;
    }
    --n;
    KB_CUDA(cudaMemcpyAsync(e->d.count, &n, sizeof(int), cudaMemcpyHostToDevice, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_env_actions(kb_env* e, int32_t* out, int cap, int* n) {
    KB_ARG(e && out && n, "env/out/n");
    EnvQuery q;
    int r = env_query(e, 1, 0, 0.0f, &q);
    if (r) return r;
    *n = q.n_actions;
    int k = q.n_actions < cap ? q.n_actions : cap;
    if (k > 0) KB_CUDA(cudaMemcpy(out, e->act_dev, sizeof(int) * k, cudaMemcpyDeviceToHost));
    return KB_OK;
}
int kb_env_observe(kb_env* e, float* obs) {
    KB_ARG(e && obs, "env/obs");
    EnvQuery q;
    int r = env_query(e, 4, 0, 0.0f, &q);
    if (r) return r;
    KB_CUDA(cudaMemcpy(obs, e->obs_dev, sizeof(float) * KB_OBSIZE, cudaMemcpyDeviceToHost));
    return KB_OK;
}
int kb_env_terminal(kb_env* e, int* terminal, float* value, int* reason) {
    KB_ARG(e && terminal && value, "env/terminal/value");
    EnvQuery q;
    int r = env_query(e, 2, 0, 0.0f, &q);
    if (r) return r;
    *terminal = q.terminal;
    *value = q.value;
    if (reason) *reason = q.reason;
    return KB_OK;
}
int kb_env_encode(kb_env* e, int move, int* action) {
    KB_ARG(e && action, "env/action");
    EnvQuery q;
    int r = env_query(e, 16, move, 0.0f, &q);
    if (r) return r;
    *action = q.code;
    return KB_OK;
}
int kb_env_decode(kb_env* e, int action, int* move) {
    KB_ARG(e && move, "env/move");
    EnvQuery q;
    int r = env_query(e, 32, action, 0.0f, &q);
    if (r) return r;
    *move = q.code;
    return KB_OK;
}
int kb_env_bootstrap(kb_env* e, float window, float* out) {
    KB_ARG(e && out, "env/out");
    EnvQuery q;
    int r = env_query(e, 8, 0, window, &q);
    if (r) return r;
    *out = q.bootstrap;
    return KB_OK;
}
int kb_env_position(kb_env* e, kb_position* out) {
    KB_ARG(e && out, "env/out");
    // on the library's stream: it is non-blocking, a legacy-stream copy would not wait for a pending reset / push
    int n = 0;
    KB_CUDA(cudaMemcpyAsync(&n, e->d.count, sizeof(int), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    KB_CUDA(cudaMemcpyAsync(out, e->d.stack + (n - 1), sizeof(Pos), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}

// ---- batched position kernels --------------------------------------------------------------
int kb_encode_planes_dev(const kb_position* pos_dev, int n, float* obs_dev) {
    KB_REQUIRE_INIT();
    KB_ARG(pos_dev && obs_dev && n >= 0, "pos/obs/n");
    if (n == 0) return KB_OK;
    k_encode_f32<<<(n + 3) / 4, 128, 0, main_stream()>>>(reinterpret_cast<const Pos*>(pos_dev), n, obs_dev);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}
int kb_encode_planes_bf16_dev(const kb_position* pos_dev, int n, void* planes_dev) {
    KB_REQUIRE_INIT();
    KB_ARG(pos_dev && planes_dev && n >= 0, "pos/planes/n");
    if (n == 0) return KB_OK;
    k_encode_tall<<<(n + 3) / 4, 128, 0, main_stream()>>>(reinterpret_cast<const Pos*>(pos_dev), n, reinterpret_cast<uint4*>(planes_dev));
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}
int kb_legal_actions_dev(const kb_position* pos_dev, int n, int32_t* actions_dev, int32_t* counts_dev) {
    KB_REQUIRE_INIT();
    KB_ARG(pos_dev && actions_dev && counts_dev && n >= 0, "pos/actions/counts/n");
    if (n == 0) return KB_OK;
    k_legal_actions<<<(n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, 32 * WARPS_PER_BLOCK, 0, main_stream()>>>(
        reinterpret_cast<const Pos*>(pos_dev), n, actions_dev, counts_dev);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        KB_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
        return KB_OK;
    }
};
}  // namespace

int kb_encode_planes(const kb_position* pos, int n, float* obs) {
    KB_REQUIRE_INIT();
    KB_ARG(pos && obs && n >= 0, "pos/obs/n");
    if (n == 0) return KB_OK;
    DevBuf dp, dout;
    int r;
    if ((r = dp.alloc(sizeof(Pos) * n)) || (r = dout.alloc(sizeof(float) * KB_OBSIZE * (size_t)n))) return r;
    KB_CUDA(cudaMemcpyAsync(dp.p, pos, sizeof(Pos) * n, cudaMemcpyHostToDevice, main_stream()));
    if ((r = kb_encode_planes_dev((const kb_position*)dp.p, n, (float*)dout.p))) return r;
    KB_CUDA(cudaMemcpyAsync(obs, dout.p, sizeof(float) * KB_OBSIZE * (size_t)n, cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_legal_actions(const kb_position* pos, int n, int32_t* actions, int32_t* counts) {
    KB_REQUIRE_INIT();
    KB_ARG(pos && actions && counts && n >= 0, "pos/actions/counts/n");
    if (n == 0) return KB_OK;
    DevBuf dp, da, dc;
    int r;
    if ((r = dp.alloc(sizeof(Pos) * n)) || (r = da.alloc(sizeof(int) * MAX_MOVES * (size_t)n)) || (r = dc.alloc(sizeof(int) * n))) return r;
    KB_CUDA(cudaMemcpyAsync(dp.p, pos, sizeof(Pos) * n, cudaMemcpyHostToDevice, main_stream()));
    if ((r = kb_legal_actions_dev((const kb_position*)dp.p, n, (int*)da.p, (int*)dc.p))) return r;
    KB_CUDA(cudaMemcpyAsync(actions, da.p, sizeof(int) * MAX_MOVES * (size_t)n, cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaMemcpyAsync(counts, dc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_apply_actions(kb_position* pos, int n, const int32_t* actions) {
    KB_REQUIRE_INIT();
    KB_ARG(pos && actions && n >= 0, "pos/actions/n");
    if (n == 0) return KB_OK;
    DevBuf dp, da;
    int r;
    if ((r = dp.alloc(sizeof(Pos) * n)) || (r = da.alloc(sizeof(int) * n))) return r;
    KB_CUDA(cudaMemcpyAsync(dp.p, pos, sizeof(Pos) * n, cudaMemcpyHostToDevice, main_stream()));
    KB_CUDA(cudaMemcpyAsync(da.p, actions, sizeof(int) * n, cudaMemcpyHostToDevice, main_stream()));
    k_apply_actions<<<(n + 127) / 128, 128, 0, main_stream()>>>((Pos*)dp.p, n, (const int*)da.p);
    KB_CUDA(cudaGetLastError());
    KB_CUDA(cudaMemcpyAsync(pos, dp.p, sizeof(Pos) * n, cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_static_eval(const kb_position* pos, int n, int32_t* eval) {
    KB_REQUIRE_INIT();
    KB_ARG(pos && eval && n >= 0, "pos/eval/n");
    if (n == 0) return KB_OK;
    DevBuf dp, de;
    int r;
    if ((r = dp.alloc(sizeof(Pos) * n)) || (r = de.alloc(sizeof(int) * n))) return r;
    KB_CUDA(cudaMemcpyAsync(dp.p, pos, sizeof(Pos) * n, cudaMemcpyHostToDevice, main_stream()));
    k_static_eval<<<(n + 127) / 128, 128, 0, main_stream()>>>((const Pos*)dp.p, n, (int*)de.p);
    KB_CUDA(cudaGetLastError());
    KB_CUDA(cudaMemcpyAsync(eval, de.p, sizeof(int) * n, cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}

}  // extern "C"

namespace kb {
int obs_to_tall_launch(const float* obs_dev, int n, void* planes, cudaStream_t st) {
    k_obs_to_tall<<<(n + 3) / 4, 128, 0, st>>>(obs_dev, n, reinterpret_cast<uint4*>(planes));
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}
}  // namespace kb

// ==========================================================================================
// C ABI: tree pools
// ==========================================================================================
constexpr int HOSTIO_MAX_GROUPS = 8, HOSTIO_MAX_CHUNKS = 4;
struct IoGroup {
    cudaStream_t cs, d2h, h2d;
    cudaEvent_t ev[4], evc[HOSTIO_MAX_CHUNKS];
    unsigned selects;
};
struct kb_pool {
    PoolDev d;
    kb_tree_cfg cfg;
    // scratch for the single-tree API
    float* obs_dev;
    float* pol_dev;
    int* int_dev;
    u64* u64_dev;
    RootInfo* info_dev;
    int* child_i;   // action[256], n[256]
    float* child_f; // w[256], p[256]
    // batched buffers
    Pos* leaf_dev;
    float* policy_dev;   // [n][4672]
    float* value_dev;    // [n][256]
    float* obs_batch_dev; // [n][1920] staging of the host-I/O path
    unsigned long long launches;
    unsigned selects_since_compact;
    int policy_mode;  // kb_pool_step: 0 softmax over the legal moves only (default), 1 dense [n][4672] softmax
    cudaEvent_t ev[6];
    IoGroup io[HOSTIO_MAX_GROUPS];  // kb_pool_step_hostio: per group of trees a compute stream and a copy stream per direction
    cudaEvent_t io_start;
    int hostio_groups;  // 0 = default
    bool io_ready;
    cudaEvent_t evs[32][4];  // phase boundaries of up to 32 evenly spaced iterations of a kb_pool_step call
    bool evs_ready;
    kb_phase_ms last;
    unsigned long long replay_tail;
};

static int pool_check(kb_pool* p, bool sync) {
    if (sync) KB_CUDA(cudaStreamSynchronize(main_stream()));
    int err = 0;
    KB_CUDA(cudaMemcpy(&err, p->d.error, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) {
        const char* what = err == KB_ERR_CAPACITY ? "node pool / path / trajectory capacity exceeded"
                           : err == KB_ERR_NO_CHILD ? "no child for action / no children to pick from"
                           : err == KB_ERR_STATE   ? "expand without a selected leaf (or no selectable child)"
                                                   : "device-side tree error";
        set_error("%s", what);
        int zero = 0;
        cudaMemcpy(p->d.error, &zero, sizeof(int), cudaMemcpyHostToDevice);
        return err;
    }
    return KB_OK;
}

static Cfg to_dev_cfg(const kb_tree_cfg& c) {
    Cfg d;
    d.cpuct = c.cpuct;
    d.force_expand_unvisited = c.force_expand_unvisited;
    d.fpu = (float)c.unvisited_node_value_pct / 100.0f;   // mcts.h:91
    d.bootstrap_weight = (float)c.bootstrap_weight / 100.0f;  // :92
    d.bootstrap_window = (float)c.bootstrap_window;           // :93
    d.bootstrap_amp = (float)c.bootstrap_amp_pct / 100.0f;    // :94
    d.scale_cpuct_by_actions = c.scale_cpuct_by_actions;
    d.noise_weight = c.noise_weight;
    d.seed = c.seed;
    d.selfplay_nodes = c.selfplay_nodes;
    d.alpha_initial = c.alpha_initial;
    d.alpha_decay = c.alpha_decay;
    d.alpha_final = c.alpha_final;
    d.alpha_cutoff = c.alpha_cutoff;
    d.draw_value = ((float)c.draw_value_pct / 100.0f) * 2.0f - 1.0f;  // selfplay.cpp:70
    d.value_index_mode = c.value_index_mode;
    return d;
}

extern "C" {

int kb_tree_default_cfg(kb_tree_cfg* c) {
    KB_ARG(c, "cfg");
    memset(c, 0, sizeof(*c));
    c->cpuct = 1.0f;
    c->unvisited_node_value_pct = 100;
    c->bootstrap_weight = 0;
    c->bootstrap_window = 1600;
    c->bootstrap_amp_pct = 75;
    c->noise_weight = 0.05f;
    c->selfplay_nodes = 0;  // 0 = no automatic moves (single-tree protocol); selfplay.cpp:17 default is 512
    c->alpha_initial = c->alpha_decay = c->alpha_final = 1.0f;
    c->alpha_cutoff = 1;
    c->draw_value_pct = 50;
    return KB_OK;
}

int kb_pool_create(kb_pool** out, int n_trees, int node_capacity, const kb_tree_cfg* cfg) {
    KB_REQUIRE_INIT();
    KB_ARG(out && cfg, "out/cfg");
    KB_ARG(n_trees > 0 && node_capacity >= 1024, "n_trees > 0 and node_capacity >= 1024");
    kb_pool* p = new (std::nothrow) kb_pool();
    if (!p) return KB_ERR_ARG;
    memset(p, 0, sizeof(*p));
    p->cfg = *cfg;
    PoolDev& d = p->d;
    d.n_trees = n_trees;
    d.tree0 = 0;
    d.tree_hi = n_trees;
    d.cap = (u32)node_capacity;
    d.cfg = to_dev_cfg(*cfg);
    // plies recorded per running game (selfplay.cpp:150-151 keeps every position of the game).  Measured on 15 k near-random
    // games (6-node budget): mean 290 plies, p99 466, max 579 -- the 50-ply rule (Q4) bounds a game by the irreversible moves
    // it can contain.  2048 leaves a wide margin at 1.2 MB per tree; beyond it the pool reports KB_ERR_CAPACITY.
    d.traj_cap = cfg->selfplay_nodes > 0 ? 2048 : 1;
    d.replay_cap = cfg->selfplay_nodes > 0 ? (n_trees * 64 < 16384 ? 16384 : n_trees * 64) : 1;
    const size_t nn = (size_t)n_trees * 2 * d.cap;
    KB_CUDA(cudaMalloc(&d.nodes, nn * sizeof(Node)));
    KB_CUDA(cudaMalloc(&d.meta, nn * sizeof(u32)));
    KB_CUDA(cudaMalloc(&d.ctl, (size_t)n_trees * sizeof(TreeCtl)));
    KB_CUDA(cudaMemsetAsync(d.ctl, 0, (size_t)n_trees * sizeof(TreeCtl), main_stream()));
    KB_CUDA(cudaMalloc(&d.error, sizeof(int)));
    KB_CUDA(cudaMemsetAsync(d.error, 0, sizeof(int), main_stream()));
    KB_CUDA(cudaMalloc(&d.stats, sizeof(Stats)));
    KB_CUDA(cudaMemsetAsync(d.stats, 0, sizeof(Stats), main_stream()));
    KB_CUDA(cudaMalloc(&d.traj, (size_t)n_trees * d.traj_cap * sizeof(TrajSample)));
    KB_CUDA(cudaMalloc(&d.replay, (size_t)d.replay_cap * sizeof(ReplaySample)));
    KB_CUDA(cudaMalloc(&d.replay_head, sizeof(unsigned long long)));
    KB_CUDA(cudaMemsetAsync(d.replay_head, 0, sizeof(unsigned long long), main_stream()));
    KB_CUDA(cudaMalloc(&d.game, sizeof(int) * (size_t)(2 + d.traj_cap)));
    KB_CUDA(cudaMemsetAsync(d.game, 0, sizeof(int) * (size_t)(2 + d.traj_cap), main_stream()));
    KB_CUDA(cudaMalloc(&p->obs_dev, sizeof(float) * KB_OBSIZE));
    KB_CUDA(cudaMalloc(&p->pol_dev, sizeof(float) * KB_PSIZE));
    KB_CUDA(cudaMalloc(&p->int_dev, sizeof(int) * 4));
    KB_CUDA(cudaMalloc(&p->u64_dev, sizeof(u64) * 2));
    KB_CUDA(cudaMalloc(&p->info_dev, sizeof(RootInfo)));
    KB_CUDA(cudaMalloc(&p->child_i, sizeof(int) * 512));
    KB_CUDA(cudaMalloc(&p->child_f, sizeof(float) * 512));
    KB_CUDA(cudaMalloc(&p->leaf_dev, sizeof(Pos) * (size_t)n_trees));
    KB_CUDA(cudaMalloc(&p->policy_dev, sizeof(float) * KB_PSIZE * (size_t)n_trees));
    KB_CUDA(cudaMalloc(&p->value_dev, sizeof(float) * KB_VALUE_WIDTH * (size_t)n_trees));
    for (int i = 0; i < 6; ++i) KB_CUDA(cudaEventCreate(&p->ev[i]));
    const int blocks = (n_trees + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    k_pool_reset<<<blocks, 32 * WARPS_PER_BLOCK, 0, main_stream()>>>(d);
    KB_CUDA(cudaGetLastError());
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    p->launches = 1;
    *out = p;
    return KB_OK;
}

int kb_pool_destroy(kb_pool* p) {
    if (!p) return KB_OK;
    cudaStreamSynchronize(main_stream());
    PoolDev& d = p->d;
    cudaFree(d.nodes); cudaFree(d.meta); cudaFree(d.ctl); cudaFree(d.error); cudaFree(d.stats);
    cudaFree(d.traj); cudaFree(d.replay); cudaFree(d.replay_head); cudaFree(d.game);
    cudaFree(p->obs_dev); cudaFree(p->pol_dev); cudaFree(p->int_dev); cudaFree(p->u64_dev); cudaFree(p->info_dev);
    cudaFree(p->child_i); cudaFree(p->child_f); cudaFree(p->leaf_dev); cudaFree(p->policy_dev); cudaFree(p->value_dev);
    cudaFree(p->obs_batch_dev); cudaFree(d.dbg);
    for (int i = 0; i < 6; ++i) cudaEventDestroy(p->ev[i]);
    if (p->io_ready) {
        for (int g = 0; g < HOSTIO_MAX_GROUPS; ++g) {
            cudaStreamDestroy(p->io[g].cs);
            cudaStreamDestroy(p->io[g].d2h);
            cudaStreamDestroy(p->io[g].h2d);
            for (int i = 0; i < 4; ++i) cudaEventDestroy(p->io[g].ev[i]);
            for (int i = 0; i < HOSTIO_MAX_CHUNKS; ++i) cudaEventDestroy(p->io[g].evc[i]);
        }
        cudaEventDestroy(p->io_start);
    }
    if (p->evs_ready)
        for (int i = 0; i < 32; ++i)
            for (int j = 0; j < 4; ++j) cudaEventDestroy(p->evs[i][j]);
    delete p;
    return KB_OK;
}
int kb_pool_size(kb_pool* p) { return p ? p->d.n_trees : KB_ERR_ARG; }

#define KB_TREE_ARGS() \
    KB_ARG(p, "pool");  \
    KB_ARG(tree >= 0 && tree < p->d.n_trees, "tree index out of range")

int kb_tree_n(kb_pool* p, int tree, int* n) {
    KB_TREE_ARGS();
    KB_ARG(n, "n");
    k_tree_root<<<1, 32, 0, main_stream()>>>(p->d, tree, p->info_dev, p->child_i, p->child_i + 256, p->child_f, p->child_f + 256, 0);
    KB_CUDA(cudaGetLastError());
    RootInfo info;
    KB_CUDA(cudaMemcpyAsync(&info, p->info_dev, sizeof(info), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    p->launches++;
    *n = info.n;
    return KB_OK;
}
// The moves from the root to the leaf that waits for the network (MCTS::select leaves the reference's Env AT that
// leaf, mcts.h:252-254; callers such as evaluate.cpp:80-90 read get_env().turn() there).
int kb_tree_leaf_path(kb_pool* p, int tree, int32_t* actions, int cap, int* depth) {
    KB_TREE_ARGS();
    KB_ARG(actions && depth && cap > 0, "actions/cap/depth");
    const int n = cap < 255 ? cap : 255;
    k_tree_leaf_path<<<1, 32, 0, main_stream()>>>(p->d, tree, p->child_i, n + 1);
    KB_CUDA(cudaGetLastError());
    int host[256];
    KB_CUDA(cudaMemcpyAsync(host, p->child_i, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    p->launches++;
    *depth = host[0];
    if (host[0] > n) {
        set_error("leaf path (%d plies) does not fit the caller's buffer (%d)", host[0], n);
        return KB_ERR_CAPACITY;
    }
    for (int i = 0; i < host[0]; ++i) actions[i] = host[1 + i];
    return KB_OK;
}
int kb_tree_select(kb_pool* p, int tree, float* obs, int* need_eval) {
    KB_TREE_ARGS();
    KB_ARG(obs && need_eval, "obs/need_eval");
    k_tree_select<<<1, 32, 0, main_stream()>>>(p->d, tree, p->obs_dev, p->int_dev);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    int ne = 0;
    KB_CUDA(cudaMemcpyAsync(&ne, p->int_dev, sizeof(int), cudaMemcpyDeviceToHost, main_stream()));
    int r = pool_check(p, true);
    if (r) return r;
    *need_eval = ne;
    if (ne) KB_CUDA(cudaMemcpy(obs, p->obs_dev, sizeof(float) * KB_OBSIZE, cudaMemcpyDeviceToHost));
    return KB_OK;
}
int kb_tree_expand(kb_pool* p, int tree, const float* policy, float value, int disable_bootstrap) {
    KB_TREE_ARGS();
    KB_ARG(policy, "policy");
    KB_CUDA(cudaMemcpyAsync(p->pol_dev, policy, sizeof(float) * KB_PSIZE, cudaMemcpyHostToDevice, main_stream()));
    k_tree_expand<<<1, 32, 0, main_stream()>>>(p->d, tree, p->pol_dev, value, disable_bootstrap);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    return pool_check(p, true);
}
int kb_tree_pick(kb_pool* p, int tree, float alpha, double u01, int* action) {
    KB_TREE_ARGS();
    KB_ARG(action, "action");
    k_tree_pick<<<1, 32, 0, main_stream()>>>(p->d, tree, alpha, u01, p->int_dev);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    int a = -1;
    KB_CUDA(cudaMemcpyAsync(&a, p->int_dev, sizeof(int), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    if (a == -2) {
        set_error("no children to pick from");
        return KB_ERR_NO_CHILD;
    }
    *action = a;
    return KB_OK;
}
int kb_tree_push(kb_pool* p, int tree, int action) {
    KB_TREE_ARGS();
    k_tree_push<<<1, 32, 0, main_stream()>>>(p->d, tree, action);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    return pool_check(p, true);
}
int kb_tree_reset(kb_pool* p, int tree) {
    KB_TREE_ARGS();
    k_tree_reset<<<1, 32, 0, main_stream()>>>(p->d, tree);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    return KB_OK;
}
int kb_tree_snapshot(kb_pool* p, int tree, float* pspace) {
    KB_TREE_ARGS();
    KB_ARG(pspace, "pspace");
    k_tree_snapshot<<<1, 256, 0, main_stream()>>>(p->d, tree, p->pol_dev);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    KB_CUDA(cudaMemcpyAsync(pspace, p->pol_dev, sizeof(float) * KB_PSIZE, cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_tree_root_children(kb_pool* p, int tree, int32_t* action, int32_t* n, float* w, float* prior, int cap, int* count) {
    KB_TREE_ARGS();
    KB_ARG(count, "count");
    if (cap > 256) cap = 256;
    k_tree_root<<<1, 32, 0, main_stream()>>>(p->d, tree, p->info_dev, p->child_i, p->child_i + 256, p->child_f, p->child_f + 256, cap);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    RootInfo info;
    KB_CUDA(cudaMemcpyAsync(&info, p->info_dev, sizeof(info), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    *count = info.k;
    int k = info.k < cap ? info.k : cap;
    if (k > 0) {
        if (action) KB_CUDA(cudaMemcpy(action, p->child_i, sizeof(int) * k, cudaMemcpyDeviceToHost));
        if (n) KB_CUDA(cudaMemcpy(n, p->child_i + 256, sizeof(int) * k, cudaMemcpyDeviceToHost));
        if (w) KB_CUDA(cudaMemcpy(w, p->child_f, sizeof(float) * k, cudaMemcpyDeviceToHost));
        if (prior) KB_CUDA(cudaMemcpy(prior, p->child_f + 256, sizeof(float) * k, cudaMemcpyDeviceToHost));
    }
    return KB_OK;
}
int kb_tree_root_w(kb_pool* p, int tree, float* w) {
    KB_TREE_ARGS();
    KB_ARG(w, "w");
    k_tree_root<<<1, 32, 0, main_stream()>>>(p->d, tree, p->info_dev, p->child_i, p->child_i + 256, p->child_f, p->child_f + 256, 0);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    RootInfo info;
    KB_CUDA(cudaMemcpyAsync(&info, p->info_dev, sizeof(info), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    *w = info.w;
    return KB_OK;
}
int kb_tree_digest(kb_pool* p, int tree, uint64_t* digest, int64_t* count) {
    KB_TREE_ARGS();
    KB_ARG(digest, "digest");
    k_tree_digest<<<1, 1, 0, main_stream()>>>(p->d, tree, p->u64_dev);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    u64 out[2];
    KB_CUDA(cudaMemcpyAsync(out, p->u64_dev, sizeof(out), cudaMemcpyDeviceToHost, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    *digest = out[0];
    if (count) *count = (int64_t)out[1];
    return KB_OK;
}
int kb_tree_env(kb_pool* p, int tree, kb_position* out) {
    KB_TREE_ARGS();
    KB_ARG(out, "out");
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    KB_CUDA(cudaMemcpy(out, &p->d.ctl[tree].root_pos, sizeof(Pos), cudaMemcpyDeviceToHost));
    return KB_OK;
}

// ---- batched phases -------------------------------------------------------------------------
static inline int pool_blocks(kb_pool* p) { return (p->d.n_trees + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK; }

// Select launch of the batched loops.  Arena compaction is deferred (push only flags a nearly full
// arena; the reserve covers many more steps) and done by k_pool_compact every COMPACT_PERIOD selects.
static int pool_launch_select(kb_pool* p, uint4* planes, Pos* leaf_out, cudaStream_t st) {
    PoolDev d = p->d;
    d.defer_compact = 1;
    if (p->selects_since_compact++ % COMPACT_PERIOD == 0) {
        k_pool_compact<<<d.n_trees, 256, 0, st>>>(d);
        KB_CUDA(cudaGetLastError());
        p->launches++;
    }
    KB_CUDA(launch_pdl(0, k_pool_select, dim3(pool_blocks(p)), dim3(32 * WARPS_PER_BLOCK), 0, st, d, planes, leaf_out));
    return KB_OK;
}

int kb_pool_select(kb_pool* p) {
    KB_ARG(p, "pool");
    int r = pool_launch_select(p, nullptr, p->leaf_dev, main_stream());
    if (r) return r;
    p->launches++;
    return pool_check(p, true);
}
int kb_pool_leaf_positions(kb_pool* p, kb_position* out) {
    KB_ARG(p && out, "pool/out");
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    KB_CUDA(cudaMemcpy(out, p->leaf_dev, sizeof(Pos) * (size_t)p->d.n_trees, cudaMemcpyDeviceToHost));
    return KB_OK;
}
int kb_pool_expand_dev(kb_pool* p, const float* policy_dev, const float* value_dev, int disable_bootstrap) {
    KB_ARG(p && policy_dev && value_dev, "pool/policy/value");
    k_pool_expand<<<pool_blocks(p), 32 * WARPS_PER_BLOCK, 0, main_stream()>>>(p->d, policy_dev, value_dev, 0, disable_bootstrap, 0);
    KB_CUDA(cudaGetLastError());
    p->launches++;
    return KB_OK;
}
int kb_pool_expand(kb_pool* p, const float* policy, const float* value, int disable_bootstrap) {
    KB_ARG(p && policy && value, "pool/policy/value");
    const size_t n = (size_t)p->d.n_trees;
    KB_CUDA(cudaMemcpyAsync(p->policy_dev, policy, sizeof(float) * KB_PSIZE * n, cudaMemcpyHostToDevice, main_stream()));
    KB_CUDA(cudaMemcpyAsync(p->value_dev, value, sizeof(float) * n, cudaMemcpyHostToDevice, main_stream()));
    int r = kb_pool_expand_dev(p, p->policy_dev, p->value_dev, disable_bootstrap);
    if (r) return r;
    return pool_check(p, true);
}

// selfplay.cpp:113-200, `iters` times, everything resident in HBM, one sync at the end.
int kb_pool_step(kb_pool* p, kb_net* net, int iters) {
    KB_ARG(p && net && iters > 0, "pool/net/iters");
    const int n = p->d.n_trees;
    int r = net_reserve(net, n);
    if (r) return r;
    cudaStream_t st = main_stream();
    uint4* planes = reinterpret_cast<uint4*>(net_input_planes(net));
    const bool legal = p->policy_mode == 0;
    if (!p->evs_ready) {
        for (int i = 0; i < 32; ++i)
            for (int j = 0; j < 4; ++j) KB_CUDA(cudaEventCreate(&p->evs[i][j]));
        p->evs_ready = true;
    }
    // phase times are averaged over up to 32 evenly spaced iterations: single steps vary a lot (a tree that
    // absorbs many terminal visits stretches its select)
    const int nsamp = iters < 32 ? iters : 32, stride = iters / nsamp;
    KB_CUDA(cudaEventRecord(p->ev[0], st));
    auto is_timed = [&](int it) { return it % stride == 0 && it / stride < nsamp; };
    static const bool fuse_ok = !(getenv("KB_NO_FUSED_EXPAND_SELECT") && atoi(getenv("KB_NO_FUSED_EXPAND_SELECT")));
    bool have_leaf = false;  // this iteration's select already ran, fused behind the previous iteration's expand
    for (int it = 0; it < iters; ++it) {
        const int k = it / stride;
        const bool timed = is_timed(it);
        if (!have_leaf) {
            if (timed) KB_CUDA(cudaEventRecord(p->evs[k][0], st));
            if ((r = pool_launch_select(p, planes, nullptr, st))) return r;
            if (timed) KB_CUDA(cudaEventRecord(p->evs[k][1], st));
            p->launches += 1;
        }
        if (legal)  // softmax over the legal moves only: priors straight into p->policy_dev[tree][128]
            r = net_forward_legal_async(net, planes, n, &p->d.ctl[0].leaf_act[0], &p->d.ctl[0].leaf_nact, sizeof(TreeCtl), p->policy_dev, p->value_dev, st);
        else
            r = net_forward_async(net, planes, n, p->policy_dev, p->value_dev, st);
        if (r) return r;
        if (timed) KB_CUDA(cudaEventRecord(p->evs[k][2], st));
        // expand(it) + select(it + 1) in one launch, unless a phase boundary is being timed, the deferred collector is due
        // between the two (it needs every tree without a pending leaf) or this is the call's last iteration
        const bool fuse = fuse_ok && it + 1 < iters && !timed && !is_timed(it + 1) && p->selects_since_compact % COMPACT_PERIOD != 0;
        if (fuse) {
            PoolDev d = p->d;
            d.defer_compact = 1;
            p->selects_since_compact++;
            KB_CUDA(launch_pdl(3, k_pool_expand_select, dim3(pool_blocks(p)), dim3(32 * WARPS_PER_BLOCK), 0, st, d, (const float*)p->policy_dev,
                               (const float*)p->value_dev, 1, 0, legal ? 1 : 0, planes, (Pos*)nullptr));
        } else {
            KB_CUDA(launch_pdl(2, k_pool_expand, dim3(pool_blocks(p)), dim3(32 * WARPS_PER_BLOCK), 0, st, p->d, (const float*)p->policy_dev,
                               (const float*)p->value_dev, 1, 0, legal ? 1 : 0));
        }
        have_leaf = fuse;
        if (timed) KB_CUDA(cudaEventRecord(p->evs[k][3], st));
        p->launches += 1 + (unsigned long long)net_launches_per_forward(net);
    }
    KB_CUDA(cudaEventRecord(p->ev[5], st));
    r = pool_check(p, true);
    if (r) return r;
    float acc[3] = {0.0f, 0.0f, 0.0f};
    for (int k = 0; k < nsamp; ++k)
        for (int j = 0; j < 3; ++j) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, p->evs[k][j], p->evs[k][j + 1]);
            acc[j] += ms;
        }
    p->last.select = acc[0] / nsamp;
    p->last.tower = acc[1] / nsamp;
    p->last.expand = acc[2] / nsamp;
    cudaEventElapsedTime(&p->last.total, p->ev[0], p->ev[5]);
    p->last.encode = 0.0f;  // fused into select
    p->last.heads = 0.0f;   // reported inside tower
    return KB_OK;
}

// Reference-shaped data flow: every iteration's observations go to the host and come back, and so
// do the policy / value rows (kami::NN::infer takes and returns host buffers, nn.cpp:155-187;
// MCTS::expand takes a host policy row, mcts.h:257).  Buffers should be pinned.
//
// The trees are served as HOSTIO_GROUPS independent groups, each the analogue of one of the reference's inference
// threads (selfplay.cpp:25-31: `inference_threads` loops, each with its own trees and its own NN::infer call on its
// own batch -- so NN::infer's value indexing, Q1, applies per group).  A group is one chain
//   select -> observe -> obs D2H -> obs H2D -> tower -> policy/value D2H -> policy/value H2D -> expand
// on its own compute stream and its own copy stream per direction; the chains of different groups are independent,
// so while one group computes or sends, another one receives: both directions of the link stay busy.
static int hostio_env(const char* name, int def, int lo, int hi) {
    const char* e = getenv(name);
    int v = e ? atoi(e) : def;
    return v < lo ? lo : v > hi ? hi : v;
}
int kb_pool_step_hostio(kb_pool* p, kb_net* net, int iters, float* obs_host, float* policy_host, float* value_host) {
    KB_ARG(p && net && iters > 0 && obs_host && policy_host && value_host, "pool/net/iters/buffers");
    const int n = p->d.n_trees;
    cudaStream_t st = main_stream();
    if (!p->obs_batch_dev) KB_CUDA(cudaMalloc(&p->obs_batch_dev, sizeof(float) * KB_OBSIZE * (size_t)n));
    if (!p->io_ready) {
        for (int g = 0; g < HOSTIO_MAX_GROUPS; ++g) {
            IoGroup& G = p->io[g];
            KB_CUDA(cudaStreamCreateWithFlags(&G.cs, cudaStreamNonBlocking));
            KB_CUDA(cudaStreamCreateWithFlags(&G.d2h, cudaStreamNonBlocking));
            KB_CUDA(cudaStreamCreateWithFlags(&G.h2d, cudaStreamNonBlocking));
            for (int i = 0; i < 4; ++i) KB_CUDA(cudaEventCreateWithFlags(&G.ev[i], cudaEventDisableTiming));
            for (int i = 0; i < HOSTIO_MAX_CHUNKS; ++i) KB_CUDA(cudaEventCreateWithFlags(&G.evc[i], cudaEventDisableTiming));
            G.selects = 0;
        }
        KB_CUDA(cudaEventCreateWithFlags(&p->io_start, cudaEventDisableTiming));
        p->io_ready = true;
    }
    int ngroups = p->hostio_groups > 0 ? p->hostio_groups : hostio_env("KB_HOSTIO_GROUPS", 4, 1, HOSTIO_MAX_GROUPS);
    if (ngroups > n) ngroups = n;
    const int per = (n + ngroups - 1) / ngroups;        // trees per group
    const int items_per = items_for(per);               // workspace items per group (groups start on item boundaries)
    int r = net_reserve(net, items_per * NB * ngroups);
    if (r) return r;
    KB_CUDA(cudaEventRecord(p->io_start, st));
    for (int g = 0; g < ngroups; ++g) KB_CUDA(cudaStreamWaitEvent(p->io[g].cs, p->io_start, 0));
    // dev -> host on the group's D2H stream once the compute stream got there, host -> dev behind it, compute resumes.
    // The rows of array 0 move in `chunks` pieces so that piece c returns while piece c+1 is still leaving.
    const int chunks = hostio_env("KB_HOSTIO_CHUNKS", 1, 1, HOSTIO_MAX_CHUNKS);
    auto round_trip = [&](IoGroup& G, int rows, float* dev0, float* host0, size_t row_floats, float* dev1, float* host1, size_t bytes1) -> int {
        KB_CUDA(cudaEventRecord(G.ev[0], G.cs));
        KB_CUDA(cudaStreamWaitEvent(G.d2h, G.ev[0], 0));
        if (bytes1) KB_CUDA(cudaMemcpyAsync(host1, dev1, bytes1, cudaMemcpyDeviceToHost, G.d2h));
        for (int c = 0; c < chunks; ++c) {
            const size_t r0 = (size_t)rows * c / chunks, r1 = (size_t)rows * (c + 1) / chunks;
            if (r1 == r0) continue;
            const size_t off = r0 * row_floats, bytes = (r1 - r0) * row_floats * sizeof(float);
            KB_CUDA(cudaMemcpyAsync(host0 + off, dev0 + off, bytes, cudaMemcpyDeviceToHost, G.d2h));
            KB_CUDA(cudaEventRecord(G.evc[c], G.d2h));
            KB_CUDA(cudaStreamWaitEvent(G.h2d, G.evc[c], 0));
            if (c == 0 && bytes1) KB_CUDA(cudaMemcpyAsync(dev1, host1, bytes1, cudaMemcpyHostToDevice, G.h2d));
            KB_CUDA(cudaMemcpyAsync(dev0 + off, host0 + off, bytes, cudaMemcpyHostToDevice, G.h2d));
        }
        KB_CUDA(cudaEventRecord(G.ev[2], G.h2d));
        KB_CUDA(cudaStreamWaitEvent(G.cs, G.ev[2], 0));
        return KB_OK;
    };
    const bool dbg = getenv("KB_HOSTIO_DEBUG") != nullptr;
    const auto host_t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < iters; ++it) {
        for (int g = 0; g < ngroups; ++g) {
            IoGroup& G = p->io[g];
            const int t0 = g * per, t1 = (g + 1) * per < n ? (g + 1) * per : n, m = t1 - t0;
            if (m <= 0) continue;
            PoolDev d = p->d;
            d.tree0 = t0;
            d.tree_hi = t1;
            d.defer_compact = 1;
            const int blocks = (m + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
            if (G.selects++ % COMPACT_PERIOD == 0) {
                k_pool_compact<<<m, 256, 0, G.cs>>>(d);
                p->launches++;
            }
            // select + Env::observe on the device; observations out to the caller's buffer and back in as NN::infer's input
            k_pool_select<<<blocks, 32 * WARPS_PER_BLOCK, 0, G.cs>>>(d, nullptr, p->leaf_dev);
            float* obs_dev = p->obs_batch_dev + (size_t)t0 * KB_OBSIZE;
            k_encode_f32<<<(m + 3) / 4, 128, 0, G.cs>>>(p->leaf_dev + t0, m, obs_dev);
            KB_CUDA(cudaGetLastError());
            if ((r = round_trip(G, m, obs_dev, obs_host + (size_t)t0 * KB_OBSIZE, KB_OBSIZE, nullptr, nullptr, 0))) return r;
            const int item0 = g * items_per;
            if ((r = obs_to_tall_launch(obs_dev, m, net_group_planes(net, item0), G.cs))) return r;
            float* pol_dev = p->policy_dev + (size_t)t0 * KB_PSIZE;
            float* val_dev = p->value_dev + (size_t)t0 * KB_VALUE_WIDTH;  // [m][256]; NN::infer hands out its first m floats (Q1)
            if ((r = net_forward_group_async(net, item0, m, pol_dev, val_dev, G.cs))) return r;
            // NN::infer's host policy / value out, MCTS::expand's host policy row / value in
            const bool fixed_value = p->d.cfg.value_index_mode == 1;  // opt-in: value[i] = vh[i][0] instead of vh.flat[i]
            if (fixed_value) {  // strided gather of column 0 through the caller's value array and back
                KB_CUDA(cudaEventRecord(G.ev[0], G.cs));
                KB_CUDA(cudaStreamWaitEvent(G.d2h, G.ev[0], 0));
                KB_CUDA(cudaMemcpy2DAsync(value_host + t0, sizeof(float), val_dev, sizeof(float) * KB_VALUE_WIDTH, sizeof(float), (size_t)m,
                                          cudaMemcpyDeviceToHost, G.d2h));
                KB_CUDA(cudaMemcpy2DAsync(val_dev, sizeof(float) * KB_VALUE_WIDTH, value_host + t0, sizeof(float), sizeof(float), (size_t)m,
                                          cudaMemcpyHostToDevice, G.d2h));
                KB_CUDA(cudaEventRecord(G.ev[2], G.d2h));
                KB_CUDA(cudaStreamWaitEvent(G.cs, G.ev[2], 0));
                if ((r = round_trip(G, m, pol_dev, policy_host + (size_t)t0 * KB_PSIZE, KB_PSIZE, nullptr, nullptr, 0))) return r;
                k_pool_expand<<<blocks, 32 * WARPS_PER_BLOCK, 0, G.cs>>>(d, p->policy_dev, p->value_dev, 1, 0, 0);
            } else {
                if ((r = round_trip(G, m, pol_dev, policy_host + (size_t)t0 * KB_PSIZE, KB_PSIZE, val_dev, value_host + t0, sizeof(float) * (size_t)m)))
                    return r;
                // value[t] of the kernel = the group's flat value row (nn.cpp:186): val_dev[t - t0]
                k_pool_expand<<<blocks, 32 * WARPS_PER_BLOCK, 0, G.cs>>>(d, p->policy_dev, val_dev - t0, 0, 0, 0);
            }
            KB_CUDA(cudaGetLastError());
            p->launches += 4 + (unsigned long long)net_launches_per_forward(net);
        }
    }
    if (dbg)
        fprintf(stderr, "[kb hostio] %d groups x %d chunks: host enqueue %.3f ms for %d iterations\n", ngroups, chunks,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(), iters);
    for (int g = 0; g < ngroups; ++g) {
        KB_CUDA(cudaEventRecord(p->io[g].ev[3], p->io[g].cs));
        KB_CUDA(cudaStreamWaitEvent(st, p->io[g].ev[3], 0));
    }
    return pool_check(p, true);
}

int kb_pool_set_hostio_groups(kb_pool* p, int groups) {
    KB_ARG(p && groups >= 0 && groups <= HOSTIO_MAX_GROUPS, "pool/groups");
    p->hostio_groups = groups;
    return KB_OK;
}

int kb_pool_get_stats(kb_pool* p, kb_pool_stats* out) {
    KB_ARG(p && out, "pool/out");
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    Stats s;
    KB_CUDA(cudaMemcpy(&s, p->d.stats, sizeof(s), cudaMemcpyDeviceToHost));
    out->evals = s.evals;
    out->moves = s.moves;
    out->games = s.games;
    out->terminal_visits = s.terminal_visits;
    out->children_scanned = s.children_scanned;
    out->path_nodes = s.path_nodes;
    out->children_created = s.children_created;
    out->samples = s.samples;
    out->nodes_in_use = 0;
    out->kernel_launches = p->launches;
    return KB_OK;
}
int kb_pool_reset_stats(kb_pool* p) {
    KB_ARG(p, "pool");
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    KB_CUDA(cudaMemsetAsync(p->d.stats, 0, sizeof(Stats), main_stream()));
    p->launches = 0;
    return KB_OK;
}
// Debug hook: per-tree cycle counters of the last select kernel: [total, select cycles, select calls,
// play_move cycles, moves, encode cycles, leaf depth, leaf actions].  enable allocates the buffer.
int kb_pool_debug_select_profile(kb_pool* p, int enable, long long* out, int cap_trees) {
    KB_ARG(p, "pool");
    if (enable && !p->d.dbg) {
        KB_CUDA(cudaMalloc(&p->d.dbg, sizeof(long long) * 8 * (size_t)p->d.n_trees));
        KB_CUDA(cudaMemsetAsync(p->d.dbg, 0, sizeof(long long) * 8 * (size_t)p->d.n_trees, main_stream()));
    }
    if (out && p->d.dbg) {
        KB_CUDA(cudaStreamSynchronize(main_stream()));
        const int n = cap_trees < p->d.n_trees ? cap_trees : p->d.n_trees;
        KB_CUDA(cudaMemcpy(out, p->d.dbg, sizeof(long long) * 8 * (size_t)n, cudaMemcpyDeviceToHost));
    }
    if (!enable && p->d.dbg) {
        cudaStreamSynchronize(main_stream());
        cudaFree(p->d.dbg);
        p->d.dbg = nullptr;
    }
    return KB_OK;
}

int kb_pool_set_policy_mode(kb_pool* p, int dense) {
    KB_ARG(p && (dense == 0 || dense == 1), "pool/mode");
    p->policy_mode = dense;
    return KB_OK;
}

int kb_pool_last_phase_ms(kb_pool* p, kb_phase_ms* out) {
    KB_ARG(p && out, "pool/out");
    *out = p->last;
    return KB_OK;
}

// Expands the device-side sparse samples into the reference's replay rows
// (obs[1920], mcts[4672], result) -- selfplay.cpp:176-186, replaybuffer.h:36-56.
// Selfplay::get_next_pgn (selfplay.h:73-80): ask for the moves of the next game that finishes, then poll.
int kb_pool_request_game(kb_pool* p) {
    KB_ARG(p, "pool");
    int one = 1;
    KB_CUDA(cudaMemcpyAsync(p->d.game, &one, sizeof(int), cudaMemcpyHostToDevice, main_stream()));
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_pool_take_game(kb_pool* p, int32_t* actions, int cap, int* count) {
    KB_ARG(p && actions && count && cap > 0, "pool/actions/cap/count");
    *count = 0;
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    int hdr[2] = {0, 0};
    KB_CUDA(cudaMemcpy(hdr, p->d.game, sizeof(hdr), cudaMemcpyDeviceToHost));
    if (hdr[0] != 3) return KB_OK;  // nothing yet (or nothing requested)
    if (hdr[1] > cap) {
        set_error("game of %d moves does not fit the caller's buffer of %d", hdr[1], cap);
        return KB_ERR_CAPACITY;
    }
    KB_CUDA(cudaMemcpy(actions, p->d.game + 2, sizeof(int) * (size_t)hdr[1], cudaMemcpyDeviceToHost));
    KB_CUDA(cudaMemset(p->d.game, 0, sizeof(int)));
    *count = hdr[1];
    return KB_OK;
}

int kb_pool_drain_samples(kb_pool* p, int max_samples, float* obs, float* pi, float* z, int* count) {
    KB_ARG(p && count, "pool/count");
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    unsigned long long head = 0;
    KB_CUDA(cudaMemcpy(&head, p->d.replay_head, sizeof(head), cudaMemcpyDeviceToHost));
    unsigned long long tail = p->replay_tail;
    if (head - tail > (unsigned long long)p->d.replay_cap) tail = head - p->d.replay_cap;  // ring overwrote the oldest
    int m = 0;
    std::vector<ReplaySample> host(1);
    std::vector<Pos> poss;
    while (tail < head && m < max_samples) {
        KB_CUDA(cudaMemcpy(host.data(), p->d.replay + (tail % p->d.replay_cap), sizeof(ReplaySample), cudaMemcpyDeviceToHost));
        const ReplaySample& rs = host[0];
        if (pi) {
            float* row = pi + (size_t)m * KB_PSIZE;
            memset(row, 0, sizeof(float) * KB_PSIZE);
            const float den = (float)(rs.s.root_n - 1);
            for (int i = 0; i < rs.s.nchild && i < TRAJ_MAX_CHILD; ++i) row[rs.s.entry[i] >> 16] = (float)(rs.s.entry[i] & 0xFFFF) / den;
        }
        if (z) z[m] = rs.z;
        poss.push_back(rs.s.pos);
        ++m;
        ++tail;
    }
    p->replay_tail = tail;
    if (obs && m > 0) {
        int r = kb_encode_planes(reinterpret_cast<const kb_position*>(poss.data()), m, obs);
        if (r) return r;
    }
    *count = m;
    return KB_OK;
}

}  // extern "C"
