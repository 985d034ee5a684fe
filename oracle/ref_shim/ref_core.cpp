// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C shim over the UNMODIFIED reference (codeandkey/kami) Env + MCTS, compiled from the
// sources where they lie under /root/reference (see oracle/Makefile).  It is loaded with
// ctypes by tests/ and by bench.py's reference arm to (a) pin the oracle restatement
// (oracle/kami_oracle.c) and (b) generate golden vectors.  Nothing here is reference
// code: it only calls the reference's public API (kami/env.h, kami/mcts.h,
// kami/options.h).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <atomic>
#include <iostream>
#include <sstream>
#include <iomanip>
#include <random>
#include <stdexcept>
#include <cmath>

// The shim peeks at Env::game (zobrist key, clocks) for differential tests only.
#define private public
#include "kami/env.h"
#include "kami/mcts.h"
#undef private
#include "kami/options.h"

using namespace kami;

extern "C" {

// ---- options (kami/options.h:5-18) -------------------------------------------------
void ref_opt_set_int(const char* key, int v) { options::setInt(key, v); }
void ref_opt_set_float(const char* key, float v) { options::setFloat(key, v); }
void ref_opt_set_str(const char* key, const char* v) { options::setStr(key, v); }
void ref_srand(unsigned s) { srand(s); }
int ref_rand() { return rand(); }

// ---- Env (kami/env.h:41-485) -------------------------------------------------------
void* ref_env_new() { return new Env(); }
void ref_env_free(void* e) { delete (Env*)e; }
int ref_env_ply(void* e) { return ((Env*)e)->ply(); }
float ref_env_turn(void* e) { return ((Env*)e)->turn(); }
void ref_env_push(void* e, int action) { ((Env*)e)->push(action); }
void ref_env_pop(void* e) { ((Env*)e)->pop(); }
int ref_env_actions(void* e, int* out, int cap) {
    std::vector<int>& a = ((Env*)e)->actions();
    int n = (int)a.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = a[i];
    return n;
}
void ref_env_observe(void* e, float* dst) { ((Env*)e)->observe(dst); }
int ref_env_terminal(void* e, float* value) {
    float v = 0.0f;
    bool t = ((Env*)e)->terminal(&v);
    *value = v;
    return t ? 1 : 0;
}
int ref_env_terminal_str(void* e, float* value, char* out, int cap) {
    float v = 0.0f;
    std::string s;
    bool t = ((Env*)e)->terminal_str(&v, s);
    *value = v;
    if (out && cap > 0) {
        strncpy(out, s.c_str(), cap - 1);
        out[cap - 1] = 0;
    }
    return t ? 1 : 0;
}
int ref_env_decode(void* e, int action) { return ((Env*)e)->decode(action); }
int ref_env_encode(void* e, int move) { return ((Env*)e)->encode(move); }
float ref_env_bootstrap(void* e, float window) { return ((Env*)e)->bootstrap_value(window); }
void ref_env_fen(void* e, char* out, int cap) {
    std::string s = ((Env*)e)->print();
    strncpy(out, s.c_str(), cap - 1);
    out[cap - 1] = 0;
}

// ---- MCTS (kami/mcts.h:66-349) -----------------------------------------------------
void* ref_mcts_new() { return new MCTS(); }
void ref_mcts_free(void* t) { delete (MCTS*)t; }
int ref_mcts_n(void* t) { return ((MCTS*)t)->n(); }
int ref_mcts_select(void* t, float* obs) { return ((MCTS*)t)->select(obs) ? 1 : 0; }
void ref_mcts_expand(void* t, float* policy, float value, int disable_bootstrap) {
    ((MCTS*)t)->expand(policy, value, disable_bootstrap != 0);
}
int ref_mcts_pick(void* t, float alpha) { return ((MCTS*)t)->pick(alpha); }
int ref_mcts_push(void* t, int action) {
    try {
        ((MCTS*)t)->push(action);
    } catch (std::exception&) {
        return -1;
    }
    return 0;
}
void ref_mcts_reset(void* t) { ((MCTS*)t)->reset(); }
void ref_mcts_snapshot(void* t, float* pspace) { ((MCTS*)t)->snapshot(pspace); }
void* ref_mcts_env(void* t) { return &((MCTS*)t)->get_env(); }
// Root children in list order: action, n, w, p.
int ref_mcts_root_children(void* t, int* action, int* n, float* w, float* p, int cap) {
    Node* r = ((MCTS*)t)->root;
    int k = (int)r->children.size();
    for (int i = 0; i < k && i < cap; ++i) {
        action[i] = r->children[i]->action;
        n[i] = r->children[i]->n;
        w[i] = r->children[i]->w;
        p[i] = r->children[i]->p;
    }
    return k;
}
float ref_mcts_root_w(void* t) { return ((MCTS*)t)->root->w; }

// Total node count + a 64-bit digest over the whole tree (pre-order: action, n, bits(w),
// bits(p)) so two implementations can be compared on every node without shipping trees.
static void digest_node(Node* nd, uint64_t* h, long* cnt) {
    auto mix = [&](uint64_t v) {
        *h ^= v + 0x9E3779B97F4A7C15ULL + (*h << 6) + (*h >> 2);
    };
    uint32_t wb, pb;
    memcpy(&wb, &nd->w, 4);
    memcpy(&pb, &nd->p, 4);
    mix((uint64_t)(uint32_t)nd->action);
    mix((uint64_t)(uint32_t)nd->n);
    mix(wb);
    mix(pb);
    mix((uint64_t)nd->children.size());
    ++*cnt;
    for (Node* c : nd->children) digest_node(c, h, cnt);
}
uint64_t ref_mcts_digest(void* t, long* count) {
    uint64_t h = 0;
    long cnt = 0;
    digest_node(((MCTS*)t)->root, &h, &cnt);
    if (count) *count = cnt;
    return h;
}

uint64_t ref_env_key(void* e) { return ncPositionGetKey(&((Env*)e)->game); }
int ref_env_hmc(void* e) { return ncPositionHalfmoveClock(&((Env*)e)->game); }
int ref_env_check(void* e) { return ncPositionIsCheck(&((Env*)e)->game); }
int ref_env_eval(void* e) { return ncPositionEvaluate(&((Env*)e)->game); }
int ref_env_repcount(void* e) { return ncPositionRepCount(&((Env*)e)->game); }
int ref_env_castle(void* e) { Env* x = (Env*)e; return x->game.ply[x->game.nply - 1].castle_rights; }
int ref_env_ep(void* e) { Env* x = (Env*)e; return x->game.ply[x->game.nply - 1].en_passant; }
int ref_env_piece_at(void* e, int sq) { return ncBoardGetPiece(&((Env*)e)->game.board, sq); }

// ---- raw neocortex accessors used to pin zobrist keys / eval -------------------------
uint64_t ref_zobrist_piece(int sq, int p) { return NC_ZOBRIST_PIECE_KEYS[sq][p]; }
uint64_t ref_zobrist_castle(int r) { return NC_ZOBRIST_CASTLE_KEYS[r]; }
uint64_t ref_zobrist_ep(int f) { return NC_ZOBRIST_EP_KEYS[f]; }
uint64_t ref_zobrist_btm() { return NC_ZOBRIST_BTM_KEY; }

}  // extern "C"
