// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C shim over the UNMODIFIED reference's Env::pgn() (kami/env.h:432-474), which prints SAN through
// the vendored thc library (kami/chess/thc/thc.cpp, compiled from where it lies under
// /root/reference by oracle/Makefile).  Used by tests/golden/make_pgn_golden.py to generate the
// PGN fixtures.  Nothing here is reference code: it only calls the reference's public API.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "kami/env.h"

using namespace kami;

extern "C" {

// A seeded random legal game on the reference Env (its own LCG, so the fixture does not depend on rand()): with
// probability bias_pct % the first legal action (the best capture by the reference's ordering), else a uniform one.
// Returns the number of plies (game played to a terminal position) or -1 if it does not fit.
int ref_random_game(unsigned long long seed, int bias_pct, int* actions, int cap) {
    Env e;
    unsigned long long x = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    auto next = [&]() { x = x * 6364136223846793005ULL + 1442695040888963407ULL; return (unsigned)(x >> 33); };
    int n = 0;
    float v;
    while (!e.terminal(&v)) {
        if (n >= cap) return -1;
        std::vector<int>& acts = e.actions();
        int a = (int)(next() % 100) < bias_pct ? acts[0] : acts[next() % acts.size()];
        actions[n++] = a;
        e.push(a);
    }
    return n;
}
// terminal reason of the position after `n` actions: the text of Env::terminal_str (env.h:288-385)
int ref_game_reason(const int* actions, int n, char* out, int cap) {
    Env e;
    for (int i = 0; i < n; ++i) e.push(actions[i]);
    float v;
    std::string s;
    bool t = e.terminal_str(&v, s);
    strncpy(out, s.c_str(), cap - 1);
    out[cap - 1] = 0;
    return t ? 1 : 0;
}

// Replays `n` actions from the start position and writes Env::pgn().  Returns the text length, -1 when the final
// position is not terminal, -2 when the buffer is too small, -3 when the game has a Q3 event (below).  q3_ply receives the first ply (0-based) at which a pawn
// reached the last rank WITHOUT promoting (queen promotions are ray moves that never promote), or -1: from
// there on thc's own board (which does promote) no longer follows the reference's game.
int ref_game_pgn(const int* actions, int n, char* out, int cap, int* q3_ply) {
    Env e;
    *q3_ply = -1;
    for (int i = 0; i < n; ++i) {
        ncMove mv = e.decode(actions[i]);
        int src = ncMoveSrc(mv), dst = ncMoveDst(mv);
        if (*q3_ply < 0 && ncMovePtype(mv) >= 12 && (dst / 8 == 0 || dst / 8 == 7)) {
            std::string fen = e.print();  // piece on src from the FEN board field
            int r = 7, f = 0;
            char pc = 0;
            for (char c : fen) {
                if (c == ' ') break;
                if (c == '/') { --r; f = 0; continue; }
                if (c >= '1' && c <= '8') { f += c - '0'; continue; }
                if (r * 8 + f == src) pc = c;
                ++f;
            }
            if (pc == 'P' || pc == 'p') *q3_ply = i;
        }
        e.push(actions[i]);
    }
    float v;
    if (!e.terminal(&v)) return -1;
    // thc's board promoted where the reference's game did not: Move::TerseIn fails on later moves, leaves the thc move
    // uninitialised and PushMove(garbage) can crash -- the reference's output is undefined for such games
    if (*q3_ply >= 0) return -3;
    std::string s = e.pgn();
    if ((int)s.size() + 1 > cap) return -2;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
}
