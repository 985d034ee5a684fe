// kami::Env over the B200 C ABI -- same public surface as the reference's kami/env.h:41-485.
// The game lives on the device (kb_env: a stack of compact positions); every method below is a
// kernel call through include/kami_b200.h.  Host code here is bookkeeping and text formatting.
#pragma once
#include <atomic>
#include <cassert>
#include <cctype>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/kami_b200.h"

// ---- the few neocortex C helpers callers use directly (types.h:240-314) -----------------------
typedef int ncMove;
typedef int ncSquare;
typedef int ncPiece;
static inline ncSquare ncMoveSrc(ncMove mv) { return (mv >> 6) & 0x3f; }
static inline ncSquare ncMoveDst(ncMove mv) { return mv & 0x3f; }
static inline ncPiece ncMovePtype(ncMove mv) { return (mv >> 12) & 0xF; }
static inline void ncMoveUCI(ncMove mv, char* dst) {
    dst[0] = (char)('a' + ncMoveSrc(mv) % 8);
    dst[1] = (char)('1' + ncMoveSrc(mv) / 8);
    dst[2] = (char)('a' + ncMoveDst(mv) % 8);
    dst[3] = (char)('1' + ncMoveDst(mv) / 8);
    dst[4] = dst[5] = '\0';
    if (ncMovePtype(mv) < 6) dst[4] = "pnbrqk"[ncMovePtype(mv)];
}
static inline ncMove ncMoveFromUci(char* uci) {
    int sf = uci[0] - 'a', sr = uci[1] - '1', df = uci[2] - 'a', dr = uci[3] - '1', pt = 0xF;
    if (uci[4]) {
        const char* types = "pnbrqk";
        const char* f = strchr(types, tolower(uci[4]));
        if (!f) return -1;
        pt = (int)(f - types);
    }
    if (sf < 0 || sf >= 8 || df < 0 || df >= 8 || sr < 0 || sr >= 8 || dr < 0 || dr >= 8) return -1;
    return (sr * 8 + sf) << 6 | (dr * 8 + df) | (pt << 12);
}

namespace kami {
constexpr int NFEATURES = 8 + 6 + 4 + 12;
constexpr int PSIZE = 73 * 64;
constexpr int WIDTH = 8;
constexpr int HEIGHT = 8;
constexpr int OBSIZE = WIDTH * HEIGHT * NFEATURES;

inline void kb_check(int rc) {
    if (rc != KB_OK) throw std::runtime_error(std::string("kami_b200: ") + kb_last_error());
}

// The reference initialises its lookup tables in a static constructor that also draws its zobrist
// keys from rand() (env.h:25-39, zobrist.c:25-54): 6344 draws before main().  Callers that never
// srand (test/encoding.cpp) depend on the stream position, so the draws are consumed here too.
class NCInit {
   public:
    NCInit() {
        static std::atomic<bool> initialized{false};
        if (!initialized.exchange(true)) {
            for (int i = 0; i < 793 * 8; ++i) (void)rand();
            std::cout << "Initialized neocortex lookup tables" << std::endl;
        }
    }
};
static NCInit nc_initializer;

class Env {
   private:
    kb_env* h = nullptr;
    float curturn = 1.0f;
    int nply = 0;
    std::vector<ncMove> history;
    std::vector<int> cur_actions;
    bool actions_utd = false;

    void replay_from(const Env& o) {
        kb_check(kb_env_reset(h));
        curturn = 1.0f;
        history.clear();
        actions_utd = false;  // (the reference's copy takes o.actions_utd / o.cur_actions; recomputing is equivalent)
        cur_actions.clear();
        for (ncMove mv : o.history) {
            int a = 0;
            kb_check(kb_env_encode(h, mv, &a));
            push(a);
        }
    }

   public:
    Env() { kb_check(kb_env_create(&h)); }
    Env(const Env& o) {
        kb_check(kb_env_create(&h));
        replay_from(o);
    }
    Env& operator=(const Env& o) {
        if (this != &o) replay_from(o);
        return *this;
    }
    ~Env() { kb_env_destroy(h); }

    int ply() { return (int)history.size(); }

    int encode(ncMove move) {
        int a = 0;
        kb_check(kb_env_encode(h, move, &a));
        return a;
    }
    ncMove decode(int action) {
        int mv = 0;
        kb_check(kb_env_decode(h, action, &mv));
        return mv;
    }
    void observe(float* dst) { kb_check(kb_env_observe(h, dst)); }
    void push(int action) {
        ncMove mv = decode(action);
        kb_check(kb_env_push(h, action));
        history.push_back(mv);
        curturn = -curturn;
        actions_utd = false;
    }
    void pop() {
        kb_check(kb_env_pop(h));
        history.pop_back();
        curturn = -curturn;
        actions_utd = false;
    }
    std::string debug_action(int action) {
        char uci[6];
        ncMoveUCI(decode(action), uci);
        return uci;
    }
    bool terminal_str(float* value, std::string& out) {
        int t = 0, reason = 0;
        kb_check(kb_env_terminal(h, &t, value, &reason));
        kb_position p;
        kb_check(kb_env_position(h, &p));
        static const char* names[] = {"", "Draw by 50-move rule", "Draw by threefold repetition", "Draw by insufficient material"};
        if (reason >= 1 && reason <= 3) out = names[reason];
        else if (reason == 4) out = p.ctm == 0 ? "White is checkmated" : "Black is checkmated";
        else if (reason == 5) out = p.ctm == 0 ? "White is stalemated" : "Black is stalemated";
        return t != 0;
    }
    bool terminal(float* value) {
        std::string unused;
        return terminal_str(value, unused);
    }
    float turn() { return curturn; }
    std::vector<int>& actions() {
        if (!actions_utd) {
            int32_t buf[KB_MAX_ACTIONS];
            int n = 0;
            kb_check(kb_env_actions(h, buf, KB_MAX_ACTIONS, &n));
            cur_actions.assign(buf, buf + n);
            actions_utd = true;
        }
        return cur_actions;
    }
    // FEN of the current position (reference: ncPositionToFen via Env::print, env.h:425-430)
    std::string print() {
        kb_position p;
        kb_check(kb_env_position(h, &p));
        std::string fen;
        for (int r = 7; r >= 0; --r) {
            int empty = 0;
            for (int f = 0; f < 8; ++f) {
                uint64_t b = 1ULL << (r * 8 + f);
                int t = -1;
                for (int i = 0; i < 6; ++i)
                    if (p.pieces[i] & b) t = i;
                if (t < 0) {
                    ++empty;
                    continue;
                }
                if (empty) fen += (char)('0' + empty), empty = 0;
                fen += (p.white & b) ? "PNBRQK"[t] : "pnbrqk"[t];
            }
            if (empty) fen += (char)('0' + empty);
            if (r) fen += '/';
        }
        fen += p.ctm == 0 ? " w " : " b ";
        if (!p.castle) fen += '-';
        else {
            if (p.castle & 1) fen += 'K';
            if (p.castle & 2) fen += 'Q';
            if (p.castle & 4) fen += 'k';
            if (p.castle & 8) fen += 'q';
        }
        fen += ' ';
        if (p.ep != 0xFF) fen += (char)('a' + p.ep % 8), fen += (char)('1' + p.ep / 8);
        else fen += '-';
        fen += " " + std::to_string((int)p.hmc) + " " + std::to_string(1 + (int)history.size() / 2);
        return fen;
    }
    // Movetext of the finished game in SAN (reference: env.h:432-474, which prints through the vendored thc library's
    // Move::NaturalOut, thc.cpp:6974-7188).  Same conventions: "e4" / "exf6" (en passant too) / "=N" suffixes, "O-O",
    // piece moves disambiguated by file, then rank, then both, and only against LEGAL moves of the same piece letter to
    // the same square; "+" for check, "#" for mate.  The game is replayed on a scratch device env; the rules are this
    // library's (= neocortex's), so a queen "promotion" that the reference's own rules never carry out (the
    // pawn stays a pawn) is written as the plain pawn move it is -- thc would print "=Q" there and lose track of the game.
    std::string pgn() {
        float value;
        std::string tstr;
        if (!terminal_str(&value, tstr)) throw std::runtime_error("Game must be in terminal state to write PGN!");
        struct Scratch {
            kb_env* e = nullptr;
            Scratch() { kb_check(kb_env_create(&e)); }
            ~Scratch() { kb_env_destroy(e); }
        } sc;
        auto type_at = [](const kb_position& p, int sq) {
            for (int i = 0; i < 6; ++i)
                if (p.pieces[i] >> sq & 1) return i;
            return -1;
        };
        auto legal_moves = [&](std::vector<ncMove>& out) {
            int32_t buf[KB_MAX_ACTIONS];
            int n = 0;
            kb_check(kb_env_actions(sc.e, buf, KB_MAX_ACTIONS, &n));
            out.resize(n);
            for (int i = 0; i < n; ++i) kb_check(kb_env_decode(sc.e, buf[i], &out[i]));
        };
        std::string out;
        int mn = 1;
        kb_position pos;
        std::vector<ncMove> legal, next_legal;
        kb_check(kb_env_position(sc.e, &pos));
        legal_moves(legal);
        for (size_t i = 0; i < history.size(); ++i) {
            if (i % 2 == 0) out += (mn == 1 ? "" : " ") + std::to_string(mn) + ".";
            else ++mn;
            const ncMove mv = history[i];
            const int src = ncMoveSrc(mv), dst = ncMoveDst(mv), promo = ncMovePtype(mv);
            const int piece = type_at(pos, src);
            const bool capture = type_at(pos, dst) >= 0 || (piece == 0 && pos.ep != 0xFF && dst == pos.ep);
            const char sf = (char)('a' + src % 8), sr = (char)('1' + src / 8), df = (char)('a' + dst % 8), dr = (char)('1' + dst / 8);
            std::string san;
            if (piece == 0) {
                if (capture) san = std::string(1, sf) + "x";
                san += df;
                san += dr;
                if (promo >= 1 && promo <= 4) san += std::string("=") + "PNBRQ"[promo];
            } else if (piece == 5 && (dst % 8 - src % 8 == 2 || dst % 8 - src % 8 == -2)) {
                san = dst % 8 > src % 8 ? "O-O" : "O-O-O";
            } else {
                // thc tries "Nd2", "Nbd2", "N1d2", "Nb1d2" in turn and keeps the first that exactly one legal move prints as
                int same = 0, same_file = 0, same_rank = 0;
                for (ncMove o : legal)
                    if (ncMoveDst(o) == dst && type_at(pos, ncMoveSrc(o)) == piece) {
                        ++same;
                        same_file += ncMoveSrc(o) % 8 == src % 8;
                        same_rank += ncMoveSrc(o) / 8 == src / 8;
                    }
                san = std::string(1, "PNBRQK"[piece]);
                if (same != 1) {
                    if (same_file == 1) san += sf;
                    else if (same_rank == 1) san += sr;
                    else san += sf, san += sr;
                }
                if (capture) san += 'x';
                san += df;
                san += dr;
            }
            int a = 0;
            kb_check(kb_env_encode(sc.e, mv, &a));
            kb_check(kb_env_push(sc.e, a));
            kb_check(kb_env_position(sc.e, &pos));
            legal_moves(next_legal);
            if (pos.check) san += next_legal.empty() ? '#' : '+';
            legal.swap(next_legal);
            out += " " + san;
        }
        std::string result = value < 0 ? "0-1" : value > 0 ? "1-0" : "1/2-1/2";
        return out + " " + result + " {" + tstr + "}";
    }
    float bootstrap_value(float window) {
        float v = 0.0f;
        kb_check(kb_env_bootstrap(h, window, &v));
        return v;
    }
};
}  // namespace kami
