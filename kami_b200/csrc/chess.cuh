// Position core of the B200 hot path: compact bitboard position, exact attack sets
// (branch-free line arithmetic, no lookup tables), pseudo-legal generation in the
// reference's emission order, SEE ordering, legality, action codec, terminal rules and
// the tapered static eval.  Under nvcc every function is __device__ ONLY: libkami_b200
// contains no host copy of the rules, so there is nothing a CPU fallback could call.
// The same source also compiles with plain g++ (tests/hostcore) so the CPU-only dev
// container can differential-test the rules before a GPU run; that build is test
// scaffolding and is never linked into the product.
//
// Reference: codeandkey/kami kami/chess/neocortex/{position,board,attacks,types}.c,
// eval.h, kami/env.h.  Behavioural quirks reproduced on purpose are tagged (Qn) after
// SURVEY.md Appendix A.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KB_HD __device__ __forceinline__
#define KB_HDN __device__
#else
#define KB_HD inline
#define KB_HDN inline
#endif

namespace kb {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

enum { PAWN = 0, KNIGHT, BISHOP, ROOK, QUEEN, KING };
enum { WHITE = 0, BLACK = 1 };
enum { T_NONE = 0, T_FIFTY = 1, T_REPETITION = 2, T_MATERIAL = 3, T_CHECKMATE = 4, T_STALEMATE = 5 };

constexpr u64 FILE_A = 0x0101010101010101ULL;
constexpr u64 FILE_H = 0x8080808080808080ULL;
constexpr u64 RANK_1 = 0xFFULL;
constexpr u64 DIAG_MAIN = 0x8040201008040201ULL;  // a1-h8
constexpr u64 DIAG_ANTI = 0x0102040810204080ULL;  // h1-a8
constexpr int MAX_MOVES = 128;                    // position.h:19
constexpr int PSIZE = 4672;
constexpr int NFEATURES = 30;
constexpr u16 NO_PROMO = 0xF000;

// Same 80-byte layout as kb_position in include/kami_b200.h.
struct Pos {
    u64 pc[6];
    u64 white;
    u64 bkey;
    u64 key;
    u8 ctm, castle, ep, hmc;
    u16 ply;
    u8 check, pad;
};
static_assert(sizeof(Pos) == 80, "kb_position layout");

// Zobrist keys: 768 piece-square, 16 castle, 8 ep file, 1 black-to-move (zobrist.c:35-54).
constexpr int ZK_CASTLE = 768, ZK_EP = 784, ZK_BTM = 792, ZK_COUNT = 793;
#if defined(__CUDACC__)
// defined here: exactly one translation unit (tree.cu) includes this header under nvcc
// __constant__: the tree descent reads it with warp-uniform indices (a broadcast from the constant cache instead of a
// global load in the middle of every key update); 793 keys = 6.3 KB
static __constant__ u64 d_zobrist[ZK_COUNT];
#define KB_ZK(i) (kb::d_zobrist[(i)])
#else
extern u64 h_zobrist[ZK_COUNT];
#define KB_ZK(i) (kb::h_zobrist[(i)])
#endif

// ---- bit helpers -------------------------------------------------------------------------
KB_HD int lsb(u64 b) {
#if defined(__CUDACC__)
    return __ffsll((long long)b) - 1;
#else
    return __builtin_ctzll(b);
#endif
}
KB_HD int popc(u64 b) {
#if defined(__CUDACC__)
    return __popcll(b);
#else
    return __builtin_popcountll(b);
#endif
}
KB_HD u64 brev(u64 b) {
#if defined(__CUDACC__)
    return __brevll(b);
#else
    b = ((b >> 1) & 0x5555555555555555ULL) | ((b & 0x5555555555555555ULL) << 1);
    b = ((b >> 2) & 0x3333333333333333ULL) | ((b & 0x3333333333333333ULL) << 2);
    b = ((b >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((b & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(b);
#endif
}
KB_HD u64 bit(int s) { return 1ULL << s; }
KB_HD int pop_lsb(u64& b) {
    int s = lsb(b);
    b &= b - 1;
    return s;
}
KB_HD u64 shl(u64 b, int d) { return d > 0 ? b << d : b >> (-d); }

// ---- attack sets (attacks.c builds the same sets as tables) --------------------------------
KB_HD u64 line_attacks(u64 occ, u64 mask, u64 sqbit) {
    // o^(o-2r) on the masked line in both directions (bit-reversal for the negative ray)
    u64 o = occ & mask;
    u64 fwd = o - 2 * sqbit;
    u64 rev = brev(brev(o) - 2 * brev(sqbit));
    return (fwd ^ rev) & mask;
}
KB_HD u64 rank_mask(int s) { return RANK_1 << (s & 56); }
KB_HD u64 file_mask(int s) { return FILE_A << (s & 7); }
KB_HD u64 diag_mask(int s) {
    int d = (s & 7) - (s >> 3);
    return d >= 0 ? DIAG_MAIN >> (8 * d) : DIAG_MAIN << (-8 * d);
}
KB_HD u64 anti_mask(int s) {
    int d = (s & 7) + (s >> 3) - 7;
    return d >= 0 ? DIAG_ANTI << (8 * d) : DIAG_ANTI >> (-8 * d);
}
KB_HD u64 rook_attacks(int s, u64 occ) {
    u64 b = bit(s);
    return line_attacks(occ, rank_mask(s), b) | line_attacks(occ, file_mask(s), b);
}
KB_HD u64 bishop_attacks(int s, u64 occ) {
    u64 b = bit(s);
    return line_attacks(occ, diag_mask(s), b) | line_attacks(occ, anti_mask(s), b);
}
KB_HD u64 king_attacks(int s) {
    u64 b = bit(s);
    u64 a = ((b << 1) & ~FILE_A) | ((b >> 1) & ~FILE_H);
    u64 row = a | b;
    return a | (row << 8) | (row >> 8);
}
KB_HD u64 knight_attacks(int s) {
    u64 b = bit(s);
    u64 l1 = (b >> 1) & ~FILE_H, l2 = (b >> 2) & 0x3F3F3F3F3F3F3F3FULL;
    u64 r1 = (b << 1) & ~FILE_A, r2 = (b << 2) & 0xFCFCFCFCFCFCFCFCULL;
    u64 h1 = l1 | r1, h2 = l2 | r2;
    return (h1 << 16) | (h1 >> 16) | (h2 << 8) | (h2 >> 8);
}
// squares a pawn of colour `col` standing on s attacks (attacks.c:69-80; empty on the last rank)
KB_HD u64 pawn_attacks(int col, int s) {
    u64 b = bit(s);
    return col == WHITE ? (((b & ~FILE_A) << 7) | ((b & ~FILE_H) << 9)) : (((b & ~FILE_A) >> 9) | ((b & ~FILE_H) >> 7));
}
// squares strictly between two aligned squares, else 0 (types.c:10-45)
KB_HD u64 between(int a, int b) {
    if (a == b) return 0;
    int fa = a & 7, ra = a >> 3, fb = b & 7, rb = b >> 3;
    u64 m;
    if (ra == rb) m = rank_mask(a);
    else if (fa == fb) m = file_mask(a);
    else if (fa - ra == fb - rb) m = diag_mask(a);
    else if (fa + ra == fb + rb) m = anti_mask(a);
    else return 0;
    int lo = a < b ? a : b, hi = a < b ? b : a;
    return m & (bit(hi) - 1) & ~(2 * bit(lo) - 1);
}

KB_HD u64 occ_all(const Pos& p) { return p.pc[0] | p.pc[1] | p.pc[2] | p.pc[3] | p.pc[4] | p.pc[5]; }
KB_HD u64 occ_col(const Pos& p, int col) { return col == WHITE ? p.white : (occ_all(p) ^ p.white); }
// piece type on s or -1
KB_HD int type_at(const Pos& p, int s) {
    u64 b = bit(s);
    int t = -1;
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (p.pc[i] & b) t = i;
    return t;
}
KB_HD int color_at(const Pos& p, int s) { return (p.white >> s) & 1 ? WHITE : BLACK; }

// board.c:201-217: pieces of either colour attacking s
KB_HD u64 attackers_to(const Pos& p, int s, u64 occ) {
    u64 black = occ ^ p.white;
    return (pawn_attacks(WHITE, s) & p.pc[PAWN] & black) | (pawn_attacks(BLACK, s) & p.pc[PAWN] & p.white) |
           (knight_attacks(s) & p.pc[KNIGHT]) | (bishop_attacks(s, occ) & (p.pc[BISHOP] | p.pc[QUEEN])) |
           (rook_attacks(s, occ) & (p.pc[ROOK] | p.pc[QUEEN])) | (king_attacks(s) & p.pc[KING]);
}
// board.c:230-238
KB_HD bool any_attacked(const Pos& p, u64 mask, int by_col) {
    u64 occ = occ_all(p);
    u64 by = by_col == WHITE ? p.white : (occ ^ p.white);
    while (mask)
        if (attackers_to(p, pop_lsb(mask), occ) & by) return true;
    return false;
}

// ---- piece placement with incremental board key (board.c:75-140) ----------------------------
KB_HD void toggle(Pos& p, int type, int col, int s) {
    u64 b = bit(s);
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (i == type) p.pc[i] ^= b;
    if (col == WHITE) p.white ^= b;
    p.bkey ^= KB_ZK(s * 12 + type * 2 + col);
}

KB_HD void set_full_key(Pos& p) {  // position.c:301-311
    u64 k = p.bkey ^ KB_ZK(ZK_CASTLE + p.castle);
    if (p.ep != 0xFF) k ^= KB_ZK(ZK_EP + (p.ep & 7));  // (Q5) xor-ed after every double push
    if (p.ctm == BLACK) k ^= KB_ZK(ZK_BTM);
    p.key = k;
}

KB_HDN void start_position(Pos& p) {  // position.c:19-36
    for (int i = 0; i < 6; ++i) p.pc[i] = 0;
    p.white = 0;
    p.bkey = 0;
    const int back[8] = {ROOK, KNIGHT, BISHOP, QUEEN, KING, BISHOP, KNIGHT, ROOK};
    for (int f = 0; f < 8; ++f) {
        toggle(p, back[f], WHITE, f);
        toggle(p, PAWN, WHITE, 8 + f);
        toggle(p, PAWN, BLACK, 48 + f);
        toggle(p, back[f], BLACK, 56 + f);
    }
    p.ctm = WHITE;
    p.castle = 0xF;
    p.ep = 0xFF;
    p.hmc = 0;
    p.ply = 0;
    p.check = 0;
    p.pad = 0;
    p.key = p.bkey;  // position.c:30: the initial key is the bare board key
}

// position.c:167-321.  `next` = position after `mv`; returns whether the mover's king is
// safe.  WITH_CHECK also computes next.check (only needed for positions that are kept).
// KNOWN_LEGAL: the move comes from a legal-action list (tree descent), the king-safety test is skipped.
template <bool WITH_CHECK, bool KNOWN_LEGAL = false>
KB_HD bool make_move(const Pos& cur, u16 mv, Pos& next) {
    next = cur;
    const int src = (mv >> 6) & 63, dst = mv & 63, promo = mv >> 12;
    const int us = cur.ctm;
    const int stype = type_at(cur, src);
    const int dtype = type_at(cur, dst);
    int hmc = cur.hmc + 1;
    toggle(next, stype, us, src);
    if (stype == PAWN) {
        hmc = 0;
        if (dst == cur.ep) toggle(next, PAWN, us ^ 1, (src & 56) | (dst & 7));  // en passant
    }
    if (stype == KING) {
        int df = (dst & 7) - (src & 7);
        if (df > 1 || df < -1) {  // castling: move the rook
            int base = us == WHITE ? 0 : 56;
            bool ks = dst > src;
            toggle(next, ROOK, us, base + (ks ? 7 : 0));
            toggle(next, ROOK, us, base + (ks ? 5 : 3));
        }
        next.castle &= us == WHITE ? ~0x3 : ~0xC;
    }
    if (dtype >= 0) {
        toggle(next, dtype, us ^ 1, dst);
        hmc = 0;
    }
    // (Q3) promotion only when ptype < 12; queen promotions arrive as NO_PROMO ray moves
    toggle(next, promo < 12 ? promo : stype, us, dst);
    u64 sd = bit(src) | bit(dst);
    if (sd & (bit(4) | bit(7))) next.castle &= ~1;
    if (sd & (bit(4) | bit(0))) next.castle &= ~2;
    if (sd & (bit(60) | bit(63))) next.castle &= ~4;
    if (sd & (bit(60) | bit(56))) next.castle &= ~8;
    int dr = (dst >> 3) - (src >> 3);
    next.ep = (stype == PAWN && (dr > 1 || dr < -1)) ? (u8)(dst + (us == WHITE ? -8 : 8)) : (u8)0xFF;
    next.hmc = (u8)hmc;
    next.ctm = (u8)(us ^ 1);
    next.ply = (u16)(cur.ply + 1);
    u64 occ = occ_all(next);
    u64 own = us == WHITE ? next.white : (occ ^ next.white);
    if (!KNOWN_LEGAL && any_attacked(next, next.pc[KING] & own, us ^ 1)) return false;
    if (WITH_CHECK) {
        set_full_key(next);
        next.check = any_attacked(next, next.pc[KING] & (occ ^ own), us) ? 1 : 0;
    }
    return true;
}

// make_move<false, true>(cur, mv, next) + set_full_key(next) for the tree descent, written for LATENCY: every lane of
// the warp runs this same scalar chain once per level (2.0 k of the 3.7 k cycles a level of the descent took), so what
// counts is the length of the dependent chain, not the instruction count.  Source and destination are cleared from all
// six boards at once and the arriving piece is or-ed into its board (independent per board) instead of three or four
// toggles that each scan the types; the two type lookups are or-trees instead of six-deep select chains; the Zobrist
// terms are requested together.  En passant, castling and promotions -- rare -- take the general routine.  Same fields
// as make_move (tests/hostcore: differential test on every legal move of random games); `check` is left as in `cur`
// like make_move<false> leaves it.
KB_HD void descend_move(const Pos& cur, u16 mv, Pos& next) {
    const int src = (mv >> 6) & 63, dst = mv & 63, promo = mv >> 12;
    const u64 sb = bit(src), db = bit(dst);
    const int us = cur.ctm;
    int stype = 0, dtype = 0;
    bool cap = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        stype |= (cur.pc[i] & sb) ? i : 0;
        dtype |= (cur.pc[i] & db) ? i : 0;
        cap |= (cur.pc[i] & db) != 0;
    }
    const int df = (dst & 7) - (src & 7);
    const bool special = (stype == PAWN && dst == cur.ep) || (stype == KING && (df > 1 || df < -1)) || promo < 12;
    if (special) {
        make_move<false, true>(cur, mv, next);
        set_full_key(next);
        return;
    }
    const u64 clr = ~(sb | db);
#pragma unroll
    for (int i = 0; i < 6; ++i) next.pc[i] = (cur.pc[i] & clr) | (i == stype ? db : 0ULL);
    next.white = (cur.white & clr) | (us == WHITE ? db : 0ULL);
    u64 bk = cur.bkey ^ KB_ZK(src * 12 + stype * 2 + us) ^ KB_ZK(dst * 12 + stype * 2 + us);
    if (cap) bk ^= KB_ZK(dst * 12 + dtype * 2 + (us ^ 1));
    next.bkey = bk;
    const u64 sd = sb | db;
    int castle = cur.castle;
    if (stype == KING) castle &= us == WHITE ? ~0x3 : ~0xC;
    if (sd & (bit(4) | bit(7))) castle &= ~1;
    if (sd & (bit(4) | bit(0))) castle &= ~2;
    if (sd & (bit(60) | bit(63))) castle &= ~4;
    if (sd & (bit(60) | bit(56))) castle &= ~8;
    next.castle = (u8)castle;
    const int dr = (dst >> 3) - (src >> 3);
    next.ep = (stype == PAWN && (dr > 1 || dr < -1)) ? (u8)(dst + (us == WHITE ? -8 : 8)) : (u8)0xFF;
    next.hmc = (stype == PAWN || cap) ? (u8)0 : (u8)(cur.hmc + 1);
    next.ctm = (u8)(us ^ 1);
    next.ply = (u16)(cur.ply + 1);
    next.check = cur.check;
    next.pad = cur.pad;
    set_full_key(next);
}

// The return value of make_move<false>(p, mv, next) -- "the mover's king is safe afterwards" (position.c:316-319) --
// without building `next`: the king's attackers are looked up in p's own bitboards under the occupancy the move leaves
// behind (source emptied, destination filled, en-passant victim removed, castling rook relocated), with the captured
// piece masked out of the attacker set.  Which piece moves does not matter for that, only whether it is the king or
// a pawn taking en passant (two bit tests instead of the six-way scans and key updates of a real make-move), so all
// lanes of a warp run the same few instructions whatever their move is.
KB_HD bool move_is_legal(const Pos& p, u16 mv) {
    const int src = (mv >> 6) & 63, dst = mv & 63;
    const int us = p.ctm;
    const u64 sb = bit(src), db = bit(dst);
    const u64 occ = occ_all(p);
    const u64 own = us == WHITE ? p.white : (occ ^ p.white);
    u64 enemy = (occ ^ own) & ~db;       // a captured piece no longer attacks
    u64 occ2 = (occ & ~sb) | db;
    const bool is_king = (p.pc[KING] & sb) != 0;
    if ((p.pc[PAWN] & sb) && dst == p.ep) {  // en passant: the victim stands beside the source (p.ep == 0xFF matches no square)
        const u64 vb = bit((src & 56) | (dst & 7));
        occ2 &= ~vb;
        enemy &= ~vb;
    }
    int ks = lsb(p.pc[KING] & own);
    if (is_king) {
        const int df = (dst & 7) - (src & 7);
        if (df > 1 || df < -1) {  // castling: the rook jumps over the king
            const int base = us == WHITE ? 0 : 56;
            const bool kside = dst > src;
            occ2 = (occ2 & ~bit(base + (kside ? 7 : 0))) | bit(base + (kside ? 5 : 3));
        }
        ks = dst;
    }
    const u64 att = (pawn_attacks(us, ks) & p.pc[PAWN]) | (knight_attacks(ks) & p.pc[KNIGHT]) |
                    (bishop_attacks(ks, occ2) & (p.pc[BISHOP] | p.pc[QUEEN])) | (rook_attacks(ks, occ2) & (p.pc[ROOK] | p.pc[QUEEN])) |
                    (king_attacks(ks) & p.pc[KING]);
    return (att & enemy) == 0;
}

// position.c:316-319: is the side to move in check
KB_HD u8 in_check(const Pos& p) {
    const u64 occ = occ_all(p);
    const u64 own = p.ctm == WHITE ? p.white : (occ ^ p.white);
    return any_attacked(p, p.pc[KING] & own, p.ctm ^ 1) ? 1 : 0;
}

// ---- pseudo-legal generation, reference emission order (position.c:360-561, 563-740) -------
struct MoveList {
    u16 mv[MAX_MOVES];
    int n;
};
KB_HD void ml_add(MoveList& l, int s, int d, int promo) {
    if (l.n < MAX_MOVES) l.mv[l.n] = (u16)((s << 6) | d | (promo << 12));
    l.n++;
}
KB_HD void emit_promos(MoveList& l, u64 dsts, int dir) {
    while (dsts) {
        int d = pop_lsb(dsts);
        ml_add(l, d - dir, d, QUEEN);  // position.c:394-397 order Q N R B
        ml_add(l, d - dir, d, KNIGHT);
        ml_add(l, d - dir, d, ROOK);
        ml_add(l, d - dir, d, BISHOP);
    }
}
KB_HD void emit_shifted(MoveList& l, u64 dsts, int dir) {
    while (dsts) {
        int d = pop_lsb(dsts);
        ml_add(l, d - dir, d, 0xF);
    }
}
template <int TYPE>
KB_HD void emit_piece(MoveList& l, u64 srcs, u64 occ, u64 allowed) {
    while (srcs) {
        int s = pop_lsb(srcs);
        u64 a;
        if (TYPE == QUEEN) a = rook_attacks(s, occ) | bishop_attacks(s, occ);
        else if (TYPE == ROOK) a = rook_attacks(s, occ);
        else if (TYPE == BISHOP) a = bishop_attacks(s, occ);
        else if (TYPE == KNIGHT) a = knight_attacks(s);
        else a = king_attacks(s);
        a &= allowed;
        while (a) ml_add(l, s, pop_lsb(a), 0xF);
    }
}
KB_HDN void gen_pseudo_legal(const Pos& p, MoveList& l) {
    l.n = 0;
    const int us = p.ctm;
    const u64 occ = occ_all(p);
    const u64 own = us == WHITE ? p.white : (occ ^ p.white), opp = occ ^ own;
    const u64 epm = p.ep != 0xFF ? bit(p.ep) : 0;
    const u64 pawns = own & p.pc[PAWN];
    const u64 promo_rank = us == WHITE ? (RANK_1 << 48) : (RANK_1 << 8);
    const u64 start_rank = us == WHITE ? (RANK_1 << 8) : (RANK_1 << 48);
    const int adv = us == WHITE ? 8 : -8, left = us == WHITE ? 7 : -9, right = us == WHITE ? 9 : -7;
    const u64 pp = pawns & promo_rank, np = pawns & ~pp;
    if (!p.check) {
        emit_promos(l, shl(pp & ~FILE_A, left) & opp, left);
        emit_promos(l, shl(pp & ~FILE_H, right) & opp, right);
        emit_promos(l, shl(pp, adv) & ~occ, adv);
        emit_shifted(l, shl(np, adv) & ~occ, adv);
        emit_shifted(l, shl(np & ~FILE_A, left) & (opp | epm), left);
        emit_shifted(l, shl(np & ~FILE_H, right) & (opp | epm), right);
        emit_shifted(l, shl(shl(pawns & start_rank, adv) & ~occ, adv) & ~occ, 2 * adv);
        emit_piece<QUEEN>(l, own & p.pc[QUEEN], occ, ~own);
        emit_piece<ROOK>(l, own & p.pc[ROOK], occ, ~own);
        emit_piece<KNIGHT>(l, own & p.pc[KNIGHT], occ, ~own);
        emit_piece<BISHOP>(l, own & p.pc[BISHOP], occ, ~own);
        emit_piece<KING>(l, own & p.pc[KING], occ, ~own);
        // castling (position.c:523-556): right, empty squares, e/f/g (e/d/c) unattacked
        const u64 crank = us == WHITE ? RANK_1 : (RANK_1 << 56);
        const int ksrc = us == WHITE ? 4 : 60;
        if ((p.castle & (1 << (us * 2))) && !(occ & crank & 0x6060606060606060ULL) &&
            !any_attacked(p, crank & 0x7070707070707070ULL, us ^ 1))
            ml_add(l, ksrc, ksrc + 2, 0xF);
        if ((p.castle & (1 << (us * 2 + 1))) && !(occ & crank & 0x0E0E0E0E0E0E0E0EULL) &&
            !any_attacked(p, crank & 0x1C1C1C1C1C1C1C1CULL, us ^ 1))
            ml_add(l, ksrc, ksrc - 2, 0xF);
        return;
    }
    // evasions: king steps, then (single checker) captures of it / ep / interpositions
    const int ksq = lsb(own & p.pc[KING]);
    const u64 checkers = attackers_to(p, ksq, occ) & opp;
    emit_piece<KING>(l, own & p.pc[KING], occ, ~own);
    if (popc(checkers) > 1) return;
    const u64 block = between(ksq, lsb(checkers));
    emit_promos(l, shl(pp & ~FILE_A, left) & checkers, left);
    emit_promos(l, shl(pp & ~FILE_H, right) & checkers, right);
    emit_promos(l, shl(pp, adv) & ~occ & block, adv);
    emit_shifted(l, shl(np, adv) & ~occ & block, adv);
    emit_shifted(l, shl(np & ~FILE_A, left) & (checkers | epm), left);
    emit_shifted(l, shl(np & ~FILE_H, right) & (checkers | epm), right);
    emit_shifted(l, shl(shl(pawns & start_rank, adv) & ~occ, adv) & ~occ & block, 2 * adv);
    emit_piece<QUEEN>(l, own & p.pc[QUEEN], occ, block | checkers);
    emit_piece<ROOK>(l, own & p.pc[ROOK], occ, block | checkers);
    emit_piece<KNIGHT>(l, own & p.pc[KNIGHT], occ, block | checkers);
    emit_piece<BISHOP>(l, own & p.pc[BISHOP], occ, block | checkers);
}

// ---- SEE (position.c:960-1080), iterative ---------------------------------------------------
// (Q7) the reference indexes its per-PIECE material table by piece TYPE.
KB_HD int see_value(int type) {
    // P=100 N=-100 B=300 R=-300 Q=300 K=-300
    return type == PAWN ? 100 : ((type & 1) ? -(type == KNIGHT ? 100 : 300) : 300);
}
// Score of a capture (or en-passant) move: value(captured) - SEE(dst, opponent), where SEE is
// the forced exchange sequence with least-valuable-attacker order P,B,N,R,Q,K and no stand-pat.
KB_HDN int see_capture(const Pos& p, u16 mv) {
    const int src = (mv >> 6) & 63, dst = mv & 63;
    u64 pc[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) pc[i] = p.pc[i];
    u64 occ = occ_all(p);
    u64 white = p.white;
    const int us = p.ctm;
    int mover = type_at(p, src);
    int victim;
    const u64 sb = bit(src), db = bit(dst);
    if (mover == PAWN && p.ep != 0xFF && dst == p.ep) {
        int csq = dst + (us == WHITE ? -8 : 8);
        victim = PAWN;
        pc[PAWN] ^= bit(csq);
        occ ^= bit(csq);
        if (us == BLACK) white ^= bit(csq);
        occ |= db;
    } else {
        victim = type_at(p, dst);
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (i == victim) pc[i] ^= db;
        if (us == BLACK) white ^= db;  // the captured piece was white
    }
    // move the capturing piece src -> dst
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (i == mover) pc[i] ^= sb | db;
    occ ^= sb;
    if (us == WHITE) white ^= sb | db;
    int total = see_value(victim);
    int sign = -1;
    int on_sq = mover;   // type of the piece now standing on dst
    int col = us ^ 1;    // side to capture next
    for (;;) {
        u64 own = col == WHITE ? white : (occ ^ white);
        u64 ba = bishop_attacks(dst, occ), ra = rook_attacks(dst, occ);
        u64 a;
        int lt;
        if ((a = pawn_attacks(col ^ 1, dst) & pc[PAWN] & own)) lt = PAWN;
        else if ((a = ba & own & pc[BISHOP])) lt = BISHOP;
        else if ((a = knight_attacks(dst) & own & pc[KNIGHT])) lt = KNIGHT;
        else if ((a = ra & own & pc[ROOK])) lt = ROOK;
        else if ((a = (ba | ra) & own & pc[QUEEN])) lt = QUEEN;
        else if ((a = king_attacks(dst) & own & pc[KING])) lt = KING;
        else break;
        u64 fb = a & (0 - a);  // least significant attacker (ncBitboardPop)
        total += sign * see_value(on_sq);
        sign = -sign;
        // attacker leaves fb, replaces the piece on dst
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i == on_sq) pc[i] ^= db;
            if (i == lt) pc[i] ^= fb | db;
        }
        occ ^= fb;
        if (col == WHITE) white ^= fb | db;  // white piece arrives, black piece on dst leaves
        else white ^= db;                    // white piece on dst is captured
        on_sq = lt;
        col ^= 1;
    }
    return total;
}
// move score for ordering (position.c:927-935): SEE for captures and ep, 0 for the rest
KB_HD int order_score(const Pos& p, u16 mv) {
    const int src = (mv >> 6) & 63, dst = mv & 63;
    const u64 occ = occ_all(p);
    int s = 0;
    if (occ & bit(dst)) s += see_capture(p, mv);
    if (p.ep != 0xFF && dst == p.ep && (p.pc[PAWN] & bit(src))) s += see_capture(p, mv);
    return s;
}

// ---- action codec (env.h:60-200) -------------------------------------------------------------
KB_HD int encode_action(const Pos& p, u16 mv) {
    int src = (mv >> 6) & 63, dst = mv & 63, promo = mv >> 12;
    const int type = type_at(p, src);
    if (p.ctm == BLACK) {
        src = 63 - src;
        dst = 63 - dst;
    }
    const int df = (dst & 7) - (src & 7), dr = (dst >> 3) - (src >> 3);
    const int adf = df < 0 ? -df : df, adr = dr < 0 ? -dr : dr;
    if (type == KNIGHT) return 73 * src + 56 + (dr < 0 ? 4 : 0) + (df > 0 ? 2 : 0) + adr - 1;
    if (type == PAWN && promo >= KNIGHT && promo <= ROOK)  // under-promotions; (Q3) queen falls through
        return 73 * src + 64 + df + (promo == KNIGHT ? 1 : (promo == BISHOP ? 4 : 7));
    const int dist = (adf > adr ? adf : adr) - 1;
    int base;
    if (df == 0) base = dr > 0 ? 0 : 7;
    else if (dr == 0) base = df > 0 ? 14 : 21;
    else if (dr > 0) base = df > 0 ? 28 : 35;
    else base = df > 0 ? 42 : 49;
    return 73 * src + base + dist;
}
// decode_action by table (the tree descent's copy: one uniform constant load instead of two divisions and three select
// chains on the critical path of every level): entry t = move type 0..72, low byte = destination - source (white's view),
// high byte = promotion piece or 0xF.  tests/hostcore checks it against decode_action on all 4 672 actions, both sides.
#if defined(__CUDACC__)
#define KB_CTAB static __constant__ const
#else
#define KB_CTAB static const
#endif
#define KB_MT(d, p) (short)((((p) & 0xF) << 8) | ((d) & 0xFF))
KB_CTAB short kb_move_type[73] = {
    KB_MT(8, 15),   KB_MT(16, 15),  KB_MT(24, 15),  KB_MT(32, 15),  KB_MT(40, 15),  KB_MT(48, 15),  KB_MT(56, 15),   // N
    KB_MT(-8, 15),  KB_MT(-16, 15), KB_MT(-24, 15), KB_MT(-32, 15), KB_MT(-40, 15), KB_MT(-48, 15), KB_MT(-56, 15),  // S
    KB_MT(1, 15),   KB_MT(2, 15),   KB_MT(3, 15),   KB_MT(4, 15),   KB_MT(5, 15),   KB_MT(6, 15),   KB_MT(7, 15),    // E
    KB_MT(-1, 15),  KB_MT(-2, 15),  KB_MT(-3, 15),  KB_MT(-4, 15),  KB_MT(-5, 15),  KB_MT(-6, 15),  KB_MT(-7, 15),   // W
    KB_MT(9, 15),   KB_MT(18, 15),  KB_MT(27, 15),  KB_MT(36, 15),  KB_MT(45, 15),  KB_MT(54, 15),  KB_MT(63, 15),   // NE
    KB_MT(7, 15),   KB_MT(14, 15),  KB_MT(21, 15),  KB_MT(28, 15),  KB_MT(35, 15),  KB_MT(42, 15),  KB_MT(49, 15),   // NW
    KB_MT(-7, 15),  KB_MT(-14, 15), KB_MT(-21, 15), KB_MT(-28, 15), KB_MT(-35, 15), KB_MT(-42, 15), KB_MT(-49, 15),  // SE
    KB_MT(-9, 15),  KB_MT(-18, 15), KB_MT(-27, 15), KB_MT(-36, 15), KB_MT(-45, 15), KB_MT(-54, 15), KB_MT(-63, 15),  // SW
    KB_MT(6, 15),   KB_MT(15, 15),  KB_MT(10, 15),  KB_MT(17, 15),  KB_MT(-10, 15), KB_MT(-17, 15), KB_MT(-6, 15), KB_MT(-15, 15),  // knight
    KB_MT(7, KNIGHT), KB_MT(8, KNIGHT), KB_MT(9, KNIGHT), KB_MT(7, BISHOP), KB_MT(8, BISHOP), KB_MT(9, BISHOP),
    KB_MT(7, ROOK),   KB_MT(8, ROOK),   KB_MT(9, ROOK)};  // underpromotions
#undef KB_MT
KB_HD u16 decode_action_tab(const Pos& p, int action) {
    int src = action / 73;
    const int e = kb_move_type[action - src * 73];
    int dst = src + (int)(signed char)(e & 0xFF);
    if (p.ctm == BLACK) {
        src = 63 - src;
        dst = 63 - dst;
    }
    return (u16)(((src & 63) << 6) | (dst & 63) | ((e >> 8) << 12));
}
KB_HD u16 decode_action(const Pos& p, int action) {
    int src = action / 73, t = action % 73, dst, promo = 0xF;
    if (t < 56) {
        const int dir = t / 7, k = t % 7 + 1;
        // N S E W NE NW SE SW
        const int step = dir == 0 ? 8 : dir == 1 ? -8 : dir == 2 ? 1 : dir == 3 ? -1 : dir == 4 ? 9 : dir == 5 ? 7 : dir == 6 ? -7 : -9;
        dst = src + step * k;
    } else if (t < 64) {
        const int k = t - 56;
        // W-NW, N-NW, E-NE, N-NE, W-SW, S-SW, E-SE, S-SE
        const int step = k == 0 ? 6 : k == 1 ? 15 : k == 2 ? 10 : k == 3 ? 17 : k == 4 ? -10 : k == 5 ? -17 : k == 6 ? -6 : -15;
        dst = src + step;
    } else {
        const int k = t - 64;
        dst = src + 7 + k % 3;
        promo = k / 3 == 0 ? KNIGHT : (k / 3 == 1 ? BISHOP : ROOK);
    }
    if (p.ctm == BLACK) {
        src = 63 - src;
        dst = 63 - dst;
    }
    return (u16)(((src & 63) << 6) | (dst & 63) | (promo << 12));
}

// ---- legal action list, scalar form (env.h:398-423) -------------------------------------------
// The CUDA kernels run a warp-cooperative version of the same steps (tree.cuh); this one is the
// single-thread statement of it, also used by the CPU-side differential tests.
KB_HDN int legal_actions_scalar(const Pos& p, u16* actions, u16* moves_out = nullptr) {
    MoveList l;
    gen_pseudo_legal(p, l);
    int n = l.n < MAX_MOVES ? l.n : MAX_MOVES;
    int score[MAX_MOVES];
    for (int i = 0; i < n; ++i) score[i] = order_score(p, l.mv[i]);
    // stable descending insertion sort (position.c:937-957, Q14)
    for (int i = 1; i < n; ++i) {
        int j = i;
        while (j > 0 && score[j - 1] < score[j]) {
            int ts = score[j]; score[j] = score[j - 1]; score[j - 1] = ts;
            u16 tm = l.mv[j]; l.mv[j] = l.mv[j - 1]; l.mv[j - 1] = tm;
            --j;
        }
    }
    int k = 0;
    for (int i = 0; i < n; ++i)
        if (move_is_legal(p, l.mv[i])) {
            if (moves_out) moves_out[k] = l.mv[i];
            actions[k++] = (u16)encode_action(p, l.mv[i]);
        }
    return k;
}

// ---- terminal rules (env.h:288-385) -----------------------------------------------------------
// hist[0..nhist) = keys of the earlier positions of the game/path, oldest first.  Only the last
// hmc entries can match (Q5), so the scan is bounded.
KB_HD int repetition_count(const Pos& p, const u64* hist, int nhist) {
    int c = 0;
    int lim = p.hmc < nhist ? p.hmc : nhist;
    for (int i = 1; i <= lim; ++i) c += hist[nhist - i] == p.key;
    return c;
}
KB_HD bool insufficient_material(const Pos& p) {
    const u64 all = occ_all(p), k = p.pc[KING], n = p.pc[KNIGHT], b = p.pc[BISHOP];
    const bool even = popc(p.white) == popc(all ^ p.white);
    return k == all || (all == (k | b) && (popc(b) == 1 || (even && popc(b) == 2))) ||
           (all == (k | n) && (popc(n) == 1 || (even && popc(n) == 2)));
}
// Pre-movegen part: returns T_FIFTY / T_REPETITION / T_MATERIAL or T_NONE (Q4, Q5)
KB_HD int terminal_before_movegen(const Pos& p, const u64* hist, int nhist) {
    if (p.hmc >= 50) return T_FIFTY;
    if (repetition_count(p, hist, nhist) > 3) return T_REPETITION;
    if (insufficient_material(p)) return T_MATERIAL;
    return T_NONE;
}
// value of a position with no legal moves (absolute, White POV)
KB_HD float no_moves_value(const Pos& p, int* reason) {
    if (p.check) {
        *reason = T_CHECKMATE;
        return p.ctm == WHITE ? -1.0f : 1.0f;
    }
    *reason = T_STALEMATE;
    return 0.0f;
}

// ---- static eval (position.c:1082-1298, eval.h) -------------------------------------------------
KB_HD int guard_value(const Pos& p, int s, u64 occ) {  // board.c:219-228, eval.h:32-40
    u64 a = attackers_to(p, s, occ);
    const u64 w = p.white;
    int v = 0;
    v += 9 * (popc(a & p.pc[PAWN] & w) - popc(a & p.pc[PAWN] & ~w));
    v += 6 * (popc(a & p.pc[KNIGHT] & w) - popc(a & p.pc[KNIGHT] & ~w));
    v += 5 * (popc(a & p.pc[BISHOP] & w) - popc(a & p.pc[BISHOP] & ~w));
    v += 2 * (popc(a & p.pc[ROOK] & w) - popc(a & p.pc[ROOK] & ~w));
    v += 1 * (popc(a & p.pc[QUEEN] & w) - popc(a & p.pc[QUEEN] & ~w));
    v += 1 * (popc(a & p.pc[KING] & w) - popc(a & p.pc[KING] & ~w));
    return v;
}
KB_HD u64 north_fill(u64 b) {
    b |= b << 8;
    b |= b << 16;
    b |= b << 32;
    return b;
}
KB_HD u64 south_fill(u64 b) {
    b |= b >> 8;
    b |= b >> 16;
    b |= b >> 32;
    return b;
}
KB_HD u64 front_spans(u64 pawns, int col) { return col == WHITE ? north_fill(pawns << 8) : south_fill(pawns >> 8); }
KB_HD u64 attack_spans(u64 pawns, int col) {
    u64 f = front_spans(pawns, col);
    return ((f << 1) & ~FILE_A) | ((f >> 1) & ~FILE_H);
}
// board.c:281-294 walks the pawns from the lowest square up and tests each one only against the pawns that come LATER
// in that walk: a pawn counts as isolated unless some pawn on an adjacent file stands on a higher square -- a higher rank,
// or the same rank one file to the right.  Closed form of exactly that set (no per-pawn loop):
KB_HD u64 isolated_pawns(u64 pawns) {
    const u64 below = south_fill(pawns >> 8);  // squares strictly below some pawn of the same file
    const u64 has_higher_neighbour = ((below << 1) & ~FILE_A) | ((below >> 1) & ~FILE_H) | ((pawns >> 1) & ~FILE_H);
    return pawns & ~has_higher_neighbour;
}
KB_HD u64 backward_pawns(u64 own_pawns, u64 opp_pawns, int col) {  // board.c:296-309
    u64 stops = shl(own_pawns, col == WHITE ? 8 : -8);
    u64 oa = shl(opp_pawns & ~FILE_A, col == WHITE ? 7 : -9) | shl(opp_pawns & ~FILE_H, col == WHITE ? 9 : -7);
    stops &= ~attack_spans(own_pawns, col);
    stops &= oa;
    return shl(stops, col == WHITE ? -8 : 8);
}
// The guard terms of the eval (centre control, king zones: position.c:1110-1145) are 20 independent
// attackers_to() evaluations -- four fifths of the eval's work.  Term i in [0, 20): 0..3 the centre squares,
// 4..11 the i-th square of the white king's zone, 12..19 of the black king's zone (absent squares add 0).
constexpr int EVAL_GUARD_TERMS = 20;
KB_HD u64 nth_bit(u64 b, int n) {
    for (int i = 0; i < n; ++i) b &= b - 1;
    return b & (~b + 1);
}
KB_HD void eval_guard_term(const Pos& p, int i, int& mg, int& eg) {
    const u64 occ = occ_all(p), W = p.white;
    if (i < 4) {
        const int sq = i == 0 ? 27 : i == 1 ? 28 : i == 2 ? 35 : 36;
        const int v = guard_value(p, sq, occ);
        mg += v * 20;
        eg += v * 8;
        return;
    }
    const bool white_zone = i < 12;
    const u64 king = p.pc[KING] & (white_zone ? W : (occ ^ W));
    const u64 one = nth_bit(king_attacks(lsb(king)), white_zone ? i - 4 : i - 12);
    if (!one) return;
    int g = guard_value(p, lsb(one), occ);
    if (white_zone ? g > 0 : g < 0) g = 0;
    mg += g * 7;
    eg += g * 6;
}
KB_HDN int static_eval(const Pos& p, const int* guard_mg = nullptr, const int* guard_eg = nullptr) {
    const u64 occ = occ_all(p), W = p.white, B = occ ^ p.white;
    int mg = 0, eg = 0;
    {
        const int mat = 100 * (popc(p.pc[PAWN] & W) - popc(p.pc[PAWN] & B)) + 300 * (popc(p.pc[KNIGHT] & W) - popc(p.pc[KNIGHT] & B)) +
                        300 * (popc(p.pc[BISHOP] & W) - popc(p.pc[BISHOP] & B)) + 500 * (popc(p.pc[ROOK] & W) - popc(p.pc[ROOK] & B)) +
                        900 * (popc(p.pc[QUEEN] & W) - popc(p.pc[QUEEN] & B)) + 1200 * (popc(p.pc[KING] & W) - popc(p.pc[KING] & B));
        mg += mat;
        eg += mat;
    }
    const int wk = lsb(W & p.pc[KING]), bk = lsb(B & p.pc[KING]);
    const u64 wka = king_attacks(wk), bka = king_attacks(bk);
    if (guard_mg) {  // the caller (a warp) has already summed the guard terms below over its lanes
        mg += *guard_mg;
        eg += *guard_eg;
    } else {
        for (int i = 0; i < EVAL_GUARD_TERMS; ++i) eval_guard_term(p, i, mg, eg);
    }
    const u64 minors = p.pc[KNIGHT] | p.pc[BISHOP];
    const int dev = popc(minors & W & 0x000000FFFFFF0000ULL) + popc(minors & B & 0x0000FFFFFF000000ULL);
    const int edge = popc(p.pc[KNIGHT] & (FILE_A | FILE_H));  // both colours, DEVELOPMENT weights (:1158-1161)
    mg += (dev + edge) * 35;
    eg += (dev + edge) * 20;
    const u64 wp = p.pc[PAWN] & W, bp = p.pc[PAWN] & B;
    const u64 wpass = ~(front_spans(bp, BLACK) | attack_spans(bp, BLACK)) & wp;
    const u64 bpass = ~(front_spans(wp, WHITE) | attack_spans(wp, WHITE)) & bp;
    mg += (popc(wpass) + popc(bpass)) * 15;
    eg += (popc(wpass) + popc(bpass)) * 30;
    if ((wk >> 3) == 0) {
        int c = popc(wka & wp & (RANK_1 << 8));
        mg += 10 + c * 8;
        eg += -10 + c * 8;
    }
    if ((bk >> 3) == 7) {
        int c = popc(bka & bp & (RANK_1 << 8));  // RANK_2 for black too (:1195)
        mg -= 10 + c * 8;
        eg -= -10 + c * 8;
    }
    for (u64 z = wpass; z;) {
        int d = (pop_lsb(z) >> 3) - 1;
        mg += d * 15;
        eg += d * 15;  // MG weight in the endgame term (:1210)
    }
    for (u64 z = bpass; z;) {
        int d = 6 - (pop_lsb(z) >> 3);
        mg -= d * 15;
        eg -= d * 15;
    }
    {   // position.c:1221-1245, one pass over the files in the reference.  (a) Open files: for every file without pawns the
        // rooks and queens on the NEXT file mask are counted -- the mask is advanced before the test uses it (:1227-1232),
        // and advancing the h-file mask gives a2..a8.  All files at once: the pawnless files shifted left by one bit.
        const u64 with_pawns = north_fill(p.pc[PAWN]) | south_fill(p.pc[PAWN]);
        const u64 nxt = ~with_pawns << 1;
        const int r = popc(nxt & p.pc[ROOK] & W) - popc(nxt & p.pc[ROOK] & B);
        const int q = popc(nxt & p.pc[QUEEN] & W) - popc(nxt & p.pc[QUEEN] & B);
        mg += (r + q) * 5;
        eg += (r + q) * 5;
        // (b) Doubled pawns: sum over the 8 files of (nw_f - 1) * -10 - (nb_f - 1) * -10; the -1 terms cancel.
        const int dp = popc(bp) - popc(wp);
        mg += dp * 10;
        eg += dp * 20;
    }
    const int wc = popc((((wp & ~FILE_A) << 7) | ((wp & ~FILE_H) << 9)) & wp);
    const int bc = popc((((bp & ~FILE_A) << 7) | ((wp & ~FILE_H) << 9)) & wp);  // (:1259) mixes colours
    mg += (wc - bc) * 4;
    eg += (wc - bc) * 4;
    const int iso = popc(isolated_pawns(wp)) - popc(isolated_pawns(bp));
    mg += iso * -10;
    eg += iso * -10;
    const int bw = popc(backward_pawns(wp, bp, WHITE)) - popc(backward_pawns(bp, wp, BLACK));
    mg += bw * -10;
    eg += bw * -10;
    int phase = 24 - popc(p.pc[KNIGHT]) - popc(p.pc[BISHOP]) - 2 * popc(p.pc[ROOK]) - 4 * popc(p.pc[QUEEN]);
    phase = (phase * 256) / 24;
    return (mg * (256 - phase) + eg * phase) / 256;
}
// env.h:476-484
KB_HD float bootstrap_value(const Pos& p, float window) {
#if defined(__CUDACC__)
    float s = __fdiv_rn((float)static_eval(p), window);
#else
    float s = (float)static_eval(p) / window;
#endif
    s = s < 1.0f ? s : 1.0f;
    s = s > -1.0f ? s : -1.0f;
    return s;
}

// ---- plane encoder (env.h:202-262) ---------------------------------------------------------------
// header features 0..17 for every square; piece one-hots 18..29 at the POV-rotated square
KB_HD float header_feature(const Pos& p, int f) {
    if (f < 8) return (float)((p.ply >> f) & 1);   // (Q13) low 8 bits of the ply
    if (f < 14) return (float)((p.hmc >> (f - 8)) & 1);
    // (Q2) raw mask values our-K, our-Q, opp-K, opp-Q
    const int wm = f == 14 ? 1 : f == 15 ? 2 : f == 16 ? 4 : 8;
    const int m = p.ctm == WHITE ? wm : (wm < 4 ? wm << 2 : wm >> 2);
    return (float)(p.castle & m);
}
// feature value f (0..29) of POV square q
KB_HD float plane_value(const Pos& p, int q, int f) {
    if (f < 18) return header_feature(p, f);
    const int s = p.ctm == BLACK ? 63 - q : q;
    const int t = type_at(p, s);
    if (t < 0) return 0.0f;
    const int idx = 18 + (color_at(p, s) != p.ctm ? 6 : 0) + t;
    return idx == f ? 1.0f : 0.0f;
}
KB_HDN void encode_planes_scalar(const Pos& p, float* dst) {
    for (int q = 0; q < 64; ++q)
        for (int f = 0; f < NFEATURES; ++f) dst[q * NFEATURES + f] = plane_value(p, q, f);
}

}  // namespace kb
