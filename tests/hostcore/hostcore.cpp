// TEST SCAFFOLDING ONLY: compiles the device position core (kami_b200/csrc/chess.cuh) for the
// host with g++ so the CPU-only container can differential-test it against the oracle before
// any GPU run.  The product library never contains or calls this build.
#include <cstring>
#include "../../kami_b200/csrc/chess.cuh"
#include "../../kami_b200/csrc/zobrist.h"

namespace kb { u64 h_zobrist[ZK_COUNT]; }
using namespace kb;

extern "C" {
void hc_init() { make_zobrist(h_zobrist, ZK_COUNT); }
void hc_start(void* pos80) { start_position(*(Pos*)pos80); }
int hc_legal_actions(const void* pos80, int* out) {
    u16 a[MAX_MOVES];
    int n = legal_actions_scalar(*(const Pos*)pos80, a);
    for (int i = 0; i < n; ++i) out[i] = a[i];
    return n;
}
// every pseudo-legal move of the position: move_is_legal must give make_move<false>'s verdict; returns the number of
// disagreements and the number of moves tested
int hc_legality_diff(const void* pos80, int* tested) {
    const Pos& p = *(const Pos*)pos80;
    MoveList l;
    gen_pseudo_legal(p, l);
    int bad = 0;
    for (int i = 0; i < l.n && i < MAX_MOVES; ++i) {
        Pos tmp;
        if (make_move<false>(p, l.mv[i], tmp) != move_is_legal(p, l.mv[i])) ++bad;
    }
    *tested = l.n < MAX_MOVES ? l.n : MAX_MOVES;
    return bad;
}
// every legal action of the position: descend_move must produce make_move<false, true> + set_full_key byte for byte
int hc_descend_diff(const void* pos80, int* tested) {
    const Pos& p = *(const Pos*)pos80;
    u16 acts[MAX_MOVES];
    const int n = legal_actions_scalar(p, acts);
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        Pos a, b;
        memset(&a, 0, sizeof(a));
        memset(&b, 0, sizeof(b));
        const u16 mv = decode_action(p, acts[i]);
        make_move<false, true>(p, mv, a);
        set_full_key(a);
        descend_move(p, mv, b);
        if (memcmp(&a, &b, sizeof(Pos)) != 0) ++bad;
    }
    *tested = n;
    return bad;
}
// decode_action_tab against decode_action on every action code, for the side to move of the given position
int hc_decode_tab_diff(const void* pos80) {
    const Pos& p = *(const Pos*)pos80;
    int bad = 0;
    for (int a = 0; a < 4672; ++a) bad += decode_action(p, a) != decode_action_tab(p, a);
    return bad;
}
int hc_push(void* pos80, int action) {
    Pos cur = *(Pos*)pos80, nx;
    bool legal = make_move<true>(cur, decode_action(cur, action), nx);
    *(Pos*)pos80 = nx;
    return legal ? 1 : 0;
}
void hc_planes(const void* pos80, float* out) { encode_planes_scalar(*(const Pos*)pos80, out); }
int hc_eval(const void* pos80) { return static_eval(*(const Pos*)pos80); }
float hc_bootstrap(const void* pos80, float w) { return bootstrap_value(*(const Pos*)pos80, w); }
int hc_encode(const void* pos80, int mv) { return encode_action(*(const Pos*)pos80, (u16)mv); }
int hc_decode(const void* pos80, int a) { return decode_action(*(const Pos*)pos80, a); }
int hc_terminal_pre(const void* pos80, const uint64_t* hist, int n) { return terminal_before_movegen(*(const Pos*)pos80, hist, n); }
}
