/* TEST INFRASTRUCTURE ONLY -- see kami_oracle.h.  Plain-C restatement of kami's CPU hot
 * path; every function cites the reference file:line it follows.  Written independently
 * (copy-make positions, ray-walk sliders) so that it is a genuine second opinion. */
#include "kami_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define BB(s) (1ULL << (s))
#define FILE_A 0x0101010101010101ULL
#define FILE_H (FILE_A << 7)
#define RANK(n) (0xFFULL << (8 * ((n)-1)))

enum { PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING };

static int g_init = 0;
static uint64_t Z_PIECE[64][12], Z_CASTLE[16], Z_EP[8], Z_BTM;
static uint64_t KING_ATT[64], KNIGHT_ATT[64], PAWN_ATT[2][64];
static uint64_t FRONTSPAN[2][64], ATTACKSPAN[2][64];
static uint64_t BETWEEN[64][64], RAYS[64][8];

static inline int lsb(uint64_t b) { return __builtin_ctzll(b); }
static inline int popcnt(uint64_t b) { return __builtin_popcountll(b); }
static inline int pop(uint64_t* b) {
    int s = lsb(*b);
    *b &= *b - 1;
    return s;
}
static inline uint64_t shift(uint64_t b, int d) { return d > 0 ? b << d : b >> -d; }

/* ---- glibc rand() (TYPE_3 additive feedback, default seed 1).  zobrist.c:25-33 draws 8
 * values per key with `rand() & 0xFF`; NCInit runs it before main (env.h:25-39), so the
 * keys are the first 6344 outputs of the seed-1 stream. */
typedef struct {
    int32_t r[34];
    uint32_t ring[31];
    int idx;
} glibc_rng;
static void grng_seed(glibc_rng* g, uint32_t seed) {
    int32_t r[344 + 34];
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        int64_t v = (16807LL * r[i - 1]) % 2147483647LL;
        if (v < 0) v += 2147483647LL;
        r[i] = (int32_t)v;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    uint32_t* u = (uint32_t*)r;
    for (int i = 34; i < 344; i++) u[i] = u[i - 31] + u[i - 3];
    for (int i = 0; i < 31; i++) g->ring[i] = u[344 - 31 + i];
    g->idx = 0;
}
static uint32_t grng_next(glibc_rng* g) {
    /* o_k = o_{k-31} + o_{k-3}; ring holds the last 31 raw values, idx = oldest. */
    uint32_t v = g->ring[g->idx] + g->ring[(g->idx + 28) % 31];
    g->ring[g->idx] = v;
    g->idx = (g->idx + 1) % 31;
    return v >> 1;
}

/* ---- attack sets (mathematical facts; reference uses magic tables, attacks.c:25-185) */
static uint64_t walk(int sq, uint64_t occ, int df, int dr) {
    uint64_t out = 0;
    int f = sq % 8 + df, r = sq / 8 + dr;
    while (f >= 0 && f < 8 && r >= 0 && r < 8) {
        uint64_t m = BB(r * 8 + f);
        out |= m;
        if (occ & m) break;
        f += df;
        r += dr;
    }
    return out;
}
static uint64_t rook_att(int sq, uint64_t occ) {
    return walk(sq, occ, 1, 0) | walk(sq, occ, -1, 0) | walk(sq, occ, 0, 1) | walk(sq, occ, 0, -1);
}
static uint64_t bishop_att(int sq, uint64_t occ) {
    return walk(sq, occ, 1, 1) | walk(sq, occ, -1, 1) | walk(sq, occ, 1, -1) | walk(sq, occ, -1, -1);
}

void ok_init(void) {
    if (g_init) return;
    glibc_rng g;
    grng_seed(&g, 1);
    uint64_t* keys[4] = {&Z_PIECE[0][0], Z_CASTLE, Z_EP, &Z_BTM};
    int counts[4] = {64 * 12, 16, 8, 1};
    for (int t = 0; t < 4; t++)
        for (int i = 0; i < counts[t]; i++) {
            uint64_t k = 0;
            for (int b = 0; b < 8; b++) k |= (uint64_t)(grng_next(&g) & 0xFF) << (8 * b);
            keys[t][i] = k;
        }
    for (int sq = 0; sq < 64; sq++) {
        int f = sq % 8, r = sq / 8;
        static const int kd[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};
        static const int nd[8][2] = {{2, 1}, {2, -1}, {-2, 1}, {-2, -1}, {1, 2}, {-1, 2}, {1, -2}, {-1, -2}};
        for (int i = 0; i < 8; i++) {
            int ff = f + kd[i][0], rr = r + kd[i][1];
            if (ff >= 0 && ff < 8 && rr >= 0 && rr < 8) KING_ATT[sq] |= BB(rr * 8 + ff);
            ff = f + nd[i][0];
            rr = r + nd[i][1];
            if (ff >= 0 && ff < 8 && rr >= 0 && rr < 8) KNIGHT_ATT[sq] |= BB(rr * 8 + ff);
        }
        /* attacks.c:69-80: pawn attack sets, empty from the last rank */
        for (int c = 0; c < 2; c++) {
            int rr = r + (c == 0 ? 1 : -1);
            if (rr >= 0 && rr < 8) {
                if (f > 0) PAWN_ATT[c][sq] |= BB(rr * 8 + f - 1);
                if (f < 7) PAWN_ATT[c][sq] |= BB(rr * 8 + f + 1);
            }
        }
        /* attacks.c:42-67: front spans / attack spans */
        for (int rr = r + 1; rr < 8; rr++) {
            FRONTSPAN[0][sq] |= BB(rr * 8 + f);
            if (f > 0) ATTACKSPAN[0][sq] |= BB(rr * 8 + f - 1);
            if (f < 7) ATTACKSPAN[0][sq] |= BB(rr * 8 + f + 1);
        }
        for (int rr = r - 1; rr >= 0; rr--) {
            FRONTSPAN[1][sq] |= BB(rr * 8 + f);
            if (f > 0) ATTACKSPAN[1][sq] |= BB(rr * 8 + f - 1);
            if (f < 7) ATTACKSPAN[1][sq] |= BB(rr * 8 + f + 1);
        }
        /* types.c:47-85 ray order: N S E W NE NW SE SW */
        static const int rd[8][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};
        for (int d = 0; d < 8; d++) RAYS[sq][d] = walk(sq, 0, rd[d][0], rd[d][1]);
    }
    /* types.c:10-45: squares strictly between two aligned squares, else 0 */
    for (int a = 0; a < 64; a++)
        for (int d = 0; d < 8; d++) {
            uint64_t ray = RAYS[a][d];
            while (ray) {
                int b = pop(&ray);
                BETWEEN[a][b] = RAYS[a][d] & RAYS[b][d < 4 ? (d ^ 1) : 11 - d];
            }
        }
    g_init = 1;
}

uint64_t ok_zobrist_piece(int sq, int p) { return Z_PIECE[sq][p]; }
uint64_t ok_zobrist_castle(int r) { return Z_CASTLE[r]; }
uint64_t ok_zobrist_ep(int f) { return Z_EP[f]; }
uint64_t ok_zobrist_btm(void) { return Z_BTM; }

/* ---- board primitives (board.c:75-140) ---------------------------------------------- */
static void put(ok_pos* p, int sq, int pc) {
    p->sq[sq] = (int8_t)pc;
    p->pieces[pc >> 1] ^= BB(sq);
    p->colors[pc & 1] ^= BB(sq);
    p->board_key ^= Z_PIECE[sq][pc];
}
static int take(ok_pos* p, int sq) {
    int pc = p->sq[sq];
    p->sq[sq] = -1;
    p->pieces[pc >> 1] ^= BB(sq);
    p->colors[pc & 1] ^= BB(sq);
    p->board_key ^= Z_PIECE[sq][pc];
    return pc;
}
static inline uint64_t occ_of(const ok_pos* p) { return p->colors[0] | p->colors[1]; }

/* board.c:201-217: all pieces of either colour attacking sq */
static uint64_t attackers(const ok_pos* p, int sq) {
    uint64_t occ = occ_of(p);
    uint64_t wp = p->pieces[PAWN] & p->colors[0], bp = p->pieces[PAWN] & p->colors[1];
    return (PAWN_ATT[0][sq] & bp) | (PAWN_ATT[1][sq] & wp) | (KNIGHT_ATT[sq] & p->pieces[KNIGHT]) |
           (bishop_att(sq, occ) & (p->pieces[BISHOP] | p->pieces[QUEEN])) |
           (rook_att(sq, occ) & (p->pieces[ROOK] | p->pieces[QUEEN])) | (KING_ATT[sq] & p->pieces[KING]);
}
/* board.c:230-238 */
static int is_attacked(const ok_pos* p, uint64_t mask, int by_col) {
    while (mask)
        if (attackers(p, pop(&mask)) & p->colors[by_col]) return 1;
    return 0;
}

static void full_key(ok_pos* p) { /* position.c:301-311 */
    uint64_t k = p->board_key;
    if (p->ep >= 0) k ^= Z_EP[p->ep % 8];
    k ^= Z_CASTLE[p->castle];
    if (p->ctm == 1) k ^= Z_BTM;
    p->key = k;
}

static void start_position(ok_pos* p) { /* position.c:19-36, board.c:70-73 */
    memset(p, 0, sizeof(*p));
    memset(p->sq, -1, sizeof(p->sq));
    static const int back[8] = {ROOK, KNIGHT, BISHOP, QUEEN, KING, BISHOP, KNIGHT, ROOK};
    /* The reference parses the FEN rank 8 first, so placement order is a8..h8, ..., a1..h1;
     * order only matters for nothing observable (xor/sum are commutative). */
    for (int f = 0; f < 8; f++) {
        put(p, 56 + f, back[f] * 2 + 1);
        put(p, 48 + f, PAWN * 2 + 1);
        put(p, 8 + f, PAWN * 2 + 0);
        put(p, f, back[f] * 2 + 0);
    }
    p->ctm = 0;
    p->castle = 0xF;
    p->ep = -1;
    p->hmc = 0;
    p->check = 0;
    p->key = p->board_key; /* position.c:30: initial key is the bare board key */
}

/* position.c:167-321.  Returns legality; `next` is always fully formed like the reference's
 * pushed ply (the caller discards it when illegal). */
static int make_move(const ok_pos* cur, int move, ok_pos* next) {
    *next = *cur;
    ok_pos* p = next;
    int src = (move >> 6) & 63, dst = move & 63, promo = (move >> 12) & 15;
    int us = cur->ctm;
    p->hmc = cur->hmc + 1;
    p->ep = -1;
    int spc = take(p, src);
    int dpc = p->sq[dst];
    int stype = spc >> 1;
    if (stype == PAWN) p->hmc = 0;
    if (stype == PAWN && dst == cur->ep) { /* en passant: remove the pawn beside us */
        take(p, (src / 8) * 8 + dst % 8);
        p->hmc = 0;
    }
    if (stype == KING && abs(src % 8 - dst % 8) > 1) { /* castling: move the rook too */
        int rank = us == 0 ? 0 : 7, ks = dst > src;
        put(p, rank * 8 + (ks ? 5 : 3), take(p, rank * 8 + (ks ? 7 : 0)));
    }
    if (dpc >= 0) {
        take(p, dst);
        put(p, dst, spc);
        p->hmc = 0;
    } else
        put(p, dst, spc);
    if (promo < 12) { /* position.c:252: ncPieceValid(ptype), 0xF = none */
        take(p, dst);
        put(p, dst, promo * 2 + us);
    }
    if (stype == KING) p->castle &= ~(us == 0 ? 0x3 : 0xC);
    uint64_t sd = BB(src) | BB(dst);
    if (sd & (BB(4) | BB(7))) p->castle &= ~1;
    if (sd & (BB(4) | BB(0))) p->castle &= ~2;
    if (sd & (BB(60) | BB(63))) p->castle &= ~4;
    if (sd & (BB(60) | BB(56))) p->castle &= ~8;
    if (stype == PAWN && abs(dst / 8 - src / 8) > 1) p->ep = dst + (us == 0 ? -8 : 8);
    p->ctm = !us;
    full_key(p);
    p->check = -1;
    if (is_attacked(p, p->pieces[KING] & p->colors[us], p->ctm)) return 0;
    p->check = is_attacked(p, p->pieces[KING] & p->colors[p->ctm], us);
    return 1;
}

#define MV(s, d) (((s) << 6) | (d) | 0xF000)
#define MVP(s, d, t) (((s) << 6) | (d) | ((t) << 12))

static int emit_promos(int* out, int n, uint64_t dsts, int dir) {
    static const int order[4] = {QUEEN, KNIGHT, ROOK, BISHOP}; /* position.c:394-397 */
    while (dsts) {
        int d = pop(&dsts);
        for (int i = 0; i < 4; i++) out[n++] = MVP(d - dir, d, order[i]);
    }
    return n;
}
static int emit_shift(int* out, int n, uint64_t dsts, int dir) {
    while (dsts) {
        int d = pop(&dsts);
        out[n++] = MV(d - dir, d);
    }
    return n;
}
static int emit_piece(int* out, int n, const ok_pos* p, int type, uint64_t allowed) {
    uint64_t srcs = p->pieces[type] & p->colors[p->ctm], occ = occ_of(p);
    while (srcs) {
        int s = pop(&srcs);
        uint64_t a = 0;
        switch (type) {
            case QUEEN: a = rook_att(s, occ) | bishop_att(s, occ); break;
            case ROOK: a = rook_att(s, occ); break;
            case BISHOP: a = bishop_att(s, occ); break;
            case KNIGHT: a = KNIGHT_ATT[s]; break;
            case KING: a = KING_ATT[s]; break;
        }
        a &= allowed;
        while (a) out[n++] = MV(s, pop(&a));
    }
    return n;
}

/* position.c:360-561 (not in check) and 563-740 (evasions): pseudo-legal moves in the
 * reference's emission order. */
static int pseudo_legal(const ok_pos* p, int* out) {
    int n = 0, us = p->ctm;
    uint64_t own = p->colors[us], opp = p->colors[!us], occ = own | opp;
    uint64_t epm = p->ep >= 0 ? BB(p->ep) : 0;
    uint64_t pawns = own & p->pieces[PAWN];
    uint64_t promo_rank = us == 0 ? RANK(7) : RANK(2), start_rank = us == 0 ? RANK(2) : RANK(7);
    int adv = us == 0 ? 8 : -8, left = us == 0 ? 7 : -9, right = us == 0 ? 9 : -7;
    uint64_t pp = pawns & promo_rank, np = pawns & ~pp;
    if (!p->check) {
        n = emit_promos(out, n, shift(pp & ~FILE_A, left) & opp, left);
        n = emit_promos(out, n, shift(pp & ~FILE_H, right) & opp, right);
        n = emit_promos(out, n, shift(pp, adv) & ~occ, adv);
        n = emit_shift(out, n, shift(np, adv) & ~occ, adv);
        n = emit_shift(out, n, shift(np & ~FILE_A, left) & (opp | epm), left);
        n = emit_shift(out, n, shift(np & ~FILE_H, right) & (opp | epm), right);
        n = emit_shift(out, n, shift(shift(pawns & start_rank, adv) & ~occ, adv) & ~occ, 2 * adv);
        n = emit_piece(out, n, p, QUEEN, ~own);
        n = emit_piece(out, n, p, ROOK, ~own);
        n = emit_piece(out, n, p, KNIGHT, ~own);
        n = emit_piece(out, n, p, BISHOP, ~own);
        n = emit_piece(out, n, p, KING, ~own);
        /* position.c:523-556 castling: rights bit, empty squares, no attacked squares */
        uint64_t crank = us == 0 ? RANK(1) : RANK(8);
        int ksrc = us == 0 ? 4 : 60;
        if ((p->castle & (1 << (us * 2))) && !(occ & crank & (FILE_A << 5 | FILE_A << 6)) &&
            !is_attacked(p, crank & (FILE_A << 4 | FILE_A << 5 | FILE_A << 6), !us))
            out[n++] = MV(ksrc, ksrc + 2);
        if ((p->castle & (1 << (us * 2 + 1))) && !(occ & crank & (FILE_A << 1 | FILE_A << 2 | FILE_A << 3)) &&
            !is_attacked(p, crank & (FILE_A << 4 | FILE_A << 3 | FILE_A << 2), !us))
            out[n++] = MV(ksrc, ksrc - 2);
        return n;
    }
    /* evasions */
    int ksq = lsb(own & p->pieces[KING]);
    uint64_t checkers = attackers(p, ksq) & opp;
    n = emit_piece(out, n, p, KING, ~own);
    if (popcnt(checkers) > 1) return n;
    uint64_t block = BETWEEN[ksq][lsb(checkers)];
    n = emit_promos(out, n, shift(pp & ~FILE_A, left) & checkers, left);
    n = emit_promos(out, n, shift(pp & ~FILE_H, right) & checkers, right);
    n = emit_promos(out, n, shift(pp, adv) & ~occ & block, adv);
    n = emit_shift(out, n, shift(np, adv) & ~occ & block, adv);
    n = emit_shift(out, n, shift(np & ~FILE_A, left) & (checkers | epm), left);
    n = emit_shift(out, n, shift(np & ~FILE_H, right) & (checkers | epm), right);
    n = emit_shift(out, n, shift(shift(pawns & start_rank, adv) & ~occ, adv) & ~occ & block, 2 * adv);
    n = emit_piece(out, n, p, QUEEN, block | checkers);
    n = emit_piece(out, n, p, ROOK, block | checkers);
    n = emit_piece(out, n, p, KNIGHT, block | checkers);
    n = emit_piece(out, n, p, BISHOP, block | checkers);
    return n;
}

/* eval.h:12-20 table is per PIECE but SEE indexes it by piece TYPE (position.c:986,1065):
 * P=100 N=-100 B=300 R=-300 Q=300 K=-300. */
static const int SEE_VALUE[6] = {100, -100, 300, -300, 300, -300};

/* position.c:1000-1080: forced exchange on sq, least valuable attacker first in the order
 * pawn, bishop, knight, rook, queen, king; no stand-pat. */
static int see(ok_pos* p, int sq, int col) {
    uint64_t own = p->colors[col], occ = occ_of(p);
    uint64_t a;
    int from = -1;
    uint64_t ba = bishop_att(sq, occ), ra = rook_att(sq, occ);
    if ((a = PAWN_ATT[!col][sq] & p->pieces[PAWN] & own)) from = lsb(a);
    else if ((a = ba & own & p->pieces[BISHOP])) from = lsb(a);
    else if ((a = KNIGHT_ATT[sq] & own & p->pieces[KNIGHT])) from = lsb(a);
    else if ((a = ra & own & p->pieces[ROOK])) from = lsb(a);
    else if ((a = (ba | ra) & own & p->pieces[QUEEN])) from = lsb(a);
    else if ((a = KING_ATT[sq] & own & p->pieces[KING])) from = lsb(a);
    if (from < 0) return 0;
    int mover = take(p, from);
    int victim = take(p, sq);
    put(p, sq, mover);
    int ret = SEE_VALUE[victim >> 1] - see(p, sq, !col);
    take(p, sq);
    put(p, sq, victim);
    put(p, from, mover);
    return ret;
}
/* position.c:960-998 */
static int see_capture(const ok_pos* cur, int move) {
    ok_pos tmp = *cur;
    ok_pos* p = &tmp;
    int src = (move >> 6) & 63, dst = move & 63;
    int victim;
    if ((p->sq[src] >> 1) == PAWN && p->ep >= 0 && dst == p->ep) {
        victim = take(p, dst + (p->ctm == 0 ? -8 : 8));
        put(p, dst, take(p, src));
    } else {
        int mover = take(p, src);
        victim = take(p, dst);
        put(p, dst, mover);
    }
    return SEE_VALUE[victim >> 1] - see(p, dst, !p->ctm);
}
/* position.c:920-958: captures/ep scored by SEE, everything else 0; stable descending
 * insertion sort. */
static void order_moves(const ok_pos* p, int* moves, int n) {
    int score[OK_MAX_MOVES];
    for (int i = 0; i < n; i++) {
        int src = (moves[i] >> 6) & 63, dst = moves[i] & 63;
        score[i] = 0;
        if (p->sq[dst] >= 0) score[i] += see_capture(p, moves[i]);
        if ((p->sq[src] >> 1) == PAWN && dst == p->ep) score[i] += see_capture(p, moves[i]);
    }
    for (int i = 1; i < n; i++) {
        int j = i;
        while (j > 0 && score[j - 1] < score[j]) {
            int t = score[j]; score[j] = score[j - 1]; score[j - 1] = t;
            t = moves[j]; moves[j] = moves[j - 1]; moves[j - 1] = t;
            j--;
        }
    }
}

/* ---- Env ---------------------------------------------------------------------------- */
#define CUR(e) (&(e)->stack[(e)->n - 1])

ok_env* ok_env_new(void) {
    ok_init();
    ok_env* e = (ok_env*)malloc(sizeof(ok_env));
    ok_env_reset(e);
    return e;
}
void ok_env_free(ok_env* e) { free(e); }
void ok_env_reset(ok_env* e) {
    e->n = 1;
    start_position(&e->stack[0]);
    e->actions_valid = 0;
    e->n_actions = 0;
}
int ok_env_ply(const ok_env* e) { return e->n - 1; }
float ok_env_turn(const ok_env* e) { return (e->n - 1) % 2 == 0 ? 1.0f : -1.0f; } /* env.h:53,269 */

/* env.h:60-143 */
static int encode_move(const ok_pos* p, int move) {
    int src = (move >> 6) & 63, dst = move & 63, promo = (move >> 12) & 15;
    int type = p->sq[src] >> 1;
    if (p->ctm == 1) { src = 63 - src; dst = 63 - dst; }
    int sr = src / 8, dr = dst / 8, sf = src % 8, df = dst % 8;
    if (type == KNIGHT) {
        int ind = (sr > dr ? 4 : 0) + (sf < df ? 2 : 0) + abs(sr - dr) - 1;
        return 73 * src + 56 + ind;
    }
    if (type == PAWN && promo < 6 && promo != QUEEN) {
        if (promo == KNIGHT) return 73 * src + 64 + (df - sf) + 1;
        if (promo == BISHOP) return 73 * src + 64 + (df - sf) + 4;
        if (promo == ROOK) return 73 * src + 64 + (df - sf) + 7;
        /* promo == PAWN/KING fall through like the reference's switch */
    }
    int ind = popcnt(BETWEEN[src][dst]);
    for (int d = 0; d < 7; d++)
        if (RAYS[src][d] & BB(dst)) return 73 * src + 7 * d + ind;
    return 73 * src + 49 + ind;
}
/* env.h:145-200 */
static int decode_action(const ok_pos* p, int action) {
    int src = action / 73, t = action % 73, dst, promo = 0xF;
    static const int ray[8] = {8, -8, 1, -1, 9, 7, -7, -9};
    static const int kn[8] = {-1 + 7, 8 + 7, 1 + 9, 8 + 9, -1 - 9, -8 - 9, 1 - 7, -8 - 7};
    if (t < 56) dst = src + ray[t / 7] * (t % 7 + 1);
    else if (t < 64) dst = src + kn[t - 56];
    else {
        static const int pd[3] = {7, 8, 9};
        static const int pt[3] = {KNIGHT, BISHOP, ROOK};
        dst = src + pd[(t - 64) % 3];
        promo = pt[(t - 64) / 3];
    }
    if (p->ctm == 1) { src = 63 - src; dst = 63 - dst; }
    return (src << 6) | dst | (promo << 12);
}
int ok_env_encode(const ok_env* e, int move) { return encode_move(CUR(e), move); }
int ok_env_decode(const ok_env* e, int action) { return decode_action(CUR(e), action); }

/* env.h:202-262 */
void ok_env_observe(const ok_env* e, float* dst) {
    const ok_pos* p = CUR(e);
    float hdr[18];
    int ply = e->n - 1;
    for (int i = 0; i < 8; i++) hdr[i] = (float)((ply >> i) & 1);
    for (int i = 0; i < 6; i++) hdr[8 + i] = (float)((p->hmc >> i) & 1);
    int ok_ = p->ctm == 0 ? 1 : 4, oq = p->ctm == 0 ? 2 : 8, pk = p->ctm == 0 ? 4 : 1, pq = p->ctm == 0 ? 8 : 2;
    hdr[14] = (float)(p->castle & ok_); /* raw mask values, not 0/1 */
    hdr[15] = (float)(p->castle & oq);
    hdr[16] = (float)(p->castle & pk);
    hdr[17] = (float)(p->castle & pq);
    memset(dst, 0, sizeof(float) * OK_OBSIZE);
    for (int s = 0; s < 64; s++) memcpy(dst + s * OK_NFEATURES, hdr, sizeof(hdr));
    for (int s = 0; s < 64; s++) {
        int pc = p->sq[s];
        if (pc < 0) continue;
        int pov = p->ctm == 1 ? 63 - s : s;
        dst[pov * OK_NFEATURES + 18 + ((pc & 1) != p->ctm ? 6 : 0) + (pc >> 1)] = 1.0f;
    }
}

void ok_env_push(ok_env* e, int action) { /* env.h:264-271 */
    int mv = decode_action(CUR(e), action);
    make_move(CUR(e), mv, &e->stack[e->n]);
    e->n++;
    e->actions_valid = 0;
}
void ok_env_pop(ok_env* e) { /* env.h:273-279 */
    e->n--;
    e->actions_valid = 0;
}

static void refresh_actions(ok_env* e) { /* env.h:398-423 */
    if (e->actions_valid) return;
    int mv[OK_MAX_MOVES];
    const ok_pos* p = CUR(e);
    int n = pseudo_legal(p, mv);
    order_moves(p, mv, n);
    e->n_actions = 0;
    ok_pos tmp;
    for (int i = 0; i < n; i++)
        if (make_move(p, mv[i], &tmp)) e->actions[e->n_actions++] = encode_move(p, mv[i]);
    e->actions_valid = 1;
}
int ok_env_actions(ok_env* e, int* out, int cap) {
    refresh_actions(e);
    for (int i = 0; i < e->n_actions && i < cap; i++) out[i] = e->actions[i];
    return e->n_actions;
}
int ok_env_repcount(const ok_env* e) { /* position.c:1347-1357 */
    int c = 0;
    for (int i = e->n - 2; i >= 0; i--) c += e->stack[i].key == CUR(e)->key;
    return c;
}
/* env.h:288-385.  reason: 1 fifty, 2 repetition, 3 material, 4 checkmate, 5 stalemate */
int ok_env_terminal(ok_env* e, float* value, int* reason) {
    const ok_pos* p = CUR(e);
    int r = 0;
    *value = 0.0f;
    if (p->hmc >= 50) r = 1;
    else if (ok_env_repcount(e) > 3) r = 2;
    else {
        uint64_t k = p->pieces[KING], n = p->pieces[KNIGHT], b = p->pieces[BISHOP], all = occ_of(p);
        int even = popcnt(p->colors[0]) == popcnt(p->colors[1]);
        if (k == all || (all == (k | b) && (popcnt(b) == 1 || (even && popcnt(b) == 2))) ||
            (all == (k | n) && (popcnt(n) == 1 || (even && popcnt(n) == 2))))
            r = 3;
        else {
            refresh_actions(e);
            if (e->n_actions == 0) {
                if (p->check) {
                    r = 4;
                    *value = p->ctm == 0 ? -1.0f : 1.0f;
                } else
                    r = 5;
            }
        }
    }
    if (reason) *reason = r;
    return r != 0;
}

/* board.c:219-228 */
static int guard(const ok_pos* p, int sq) {
    static const int G[6] = {9, 6, 5, 2, 1, 1}; /* eval.h:32-40, sign by colour */
    int v = 0;
    uint64_t a = attackers(p, sq);
    while (a) {
        int pc = p->sq[pop(&a)];
        v += (pc & 1) ? -G[pc >> 1] : G[pc >> 1];
    }
    return v;
}
static uint64_t spans(const ok_pos* p, int col, int front, int att) { /* board.c:240-279 */
    uint64_t out = 0, pw = p->pieces[PAWN] & p->colors[col];
    while (pw) {
        int s = pop(&pw);
        if (front) out |= FRONTSPAN[col][s];
        if (att) out |= ATTACKSPAN[col][s];
    }
    return out;
}
static uint64_t isolated(const ok_pos* p, int col) { /* board.c:281-294: only LATER pawns are tested */
    uint64_t pw = p->pieces[PAWN] & p->colors[col], out = 0;
    while (pw) {
        int s = pop(&pw), f = s % 8;
        uint64_t nb = (f > 0 ? FILE_A << (f - 1) : 0) | (f < 7 ? FILE_A << (f + 1) : 0);
        if (!(nb & pw)) out |= BB(s);
    }
    return out;
}
static uint64_t backward(const ok_pos* p, int col) { /* board.c:296-309 */
    uint64_t pw = p->pieces[PAWN] & p->colors[col], op = p->pieces[PAWN] & p->colors[!col];
    uint64_t stops = shift(pw, col == 0 ? 8 : -8);
    uint64_t oa = shift(op & ~FILE_A, col == 0 ? 7 : -9) | shift(op & ~FILE_H, col == 0 ? 9 : -7);
    stops &= ~spans(p, col, 0, 1);
    stops &= oa;
    return shift(stops, col == 0 ? -8 : 8);
}
/* position.c:1082-1298 with eval.h weights; integer, reproduces the reference's slips
 * (black chain from white pawns :1259, edge knights on DEVELOPMENT weights :1158-1161,
 * both sides' development added :1147-1150, black shield on RANK_2 :1195, advanced-passer
 * MG weight in the endgame term :1210,1219). */
static int evaluate(const ok_pos* p) {
    int mg = 0, eg = 0;
    static const int MAT[6] = {100, 300, 300, 500, 900, 1200};
    for (int t = 0; t < 6; t++) {
        int d = MAT[t] * (popcnt(p->pieces[t] & p->colors[0]) - popcnt(p->pieces[t] & p->colors[1]));
        mg += d;
        eg += d;
    }
    int wk = lsb(p->colors[0] & p->pieces[KING]), bk = lsb(p->colors[1] & p->pieces[KING]);
    static const int C[4] = {27, 28, 35, 36};
    for (int i = 0; i < 4; i++) {
        int v = guard(p, C[i]);
        mg += v * 20;
        eg += v * 8;
    }
    uint64_t z = KING_ATT[wk];
    while (z) {
        int g = guard(p, pop(&z));
        if (g > 0) g = 0;
        mg += g * 7;
        eg += g * 6;
    }
    z = KING_ATT[bk];
    while (z) {
        int g = guard(p, pop(&z));
        if (g < 0) g = 0;
        mg += g * 7;
        eg += g * 6;
    }
    uint64_t minors = p->pieces[KNIGHT] | p->pieces[BISHOP];
    int wd = popcnt(minors & p->colors[0] & (RANK(3) | RANK(4) | RANK(5)));
    int bd = popcnt(minors & p->colors[1] & (RANK(4) | RANK(5) | RANK(6)));
    mg += (wd + bd) * 35;
    eg += (wd + bd) * 20;
    uint64_t edge = p->pieces[KNIGHT] & (FILE_A | FILE_H);
    int ne = popcnt(edge & p->colors[0]) + popcnt(edge & p->colors[1]);
    mg += ne * 35;
    eg += ne * 20;
    uint64_t wpass = ~spans(p, 1, 1, 1) & p->pieces[PAWN] & p->colors[0];
    uint64_t bpass = ~spans(p, 0, 1, 1) & p->pieces[PAWN] & p->colors[1];
    mg += (popcnt(wpass) + popcnt(bpass)) * 15;
    eg += (popcnt(wpass) + popcnt(bpass)) * 30;
    if (wk / 8 == 0) {
        int c = popcnt(KING_ATT[wk] & p->colors[0] & p->pieces[PAWN] & RANK(2));
        mg += 10 + c * 8;
        eg += -10 + c * 8;
    }
    if (bk / 8 == 7) {
        int c = popcnt(KING_ATT[bk] & p->colors[1] & p->pieces[PAWN] & RANK(2));
        mg -= 10 + c * 8;
        eg -= -10 + c * 8;
    }
    z = wpass;
    while (z) {
        int d = pop(&z) / 8 - 1;
        mg += d * 15;
        eg += d * 15;
    }
    z = bpass;
    while (z) {
        int d = 6 - pop(&z) / 8;
        mg -= d * 15;
        eg -= d * 15;
    }
    for (int f = 0; f < 8; f++) {
        uint64_t file = FILE_A << f, fp = p->pieces[PAWN] & file;
        /* position.c:1226-1232: the file mask is shifted BEFORE the open-file test uses it,
         * so rooks/queens are counted on file f+1 (for f == 7: a2..a8). */
        uint64_t nxt = file << 1;
        if (!fp) {
            int r = popcnt(nxt & p->pieces[ROOK] & p->colors[0]) - popcnt(nxt & p->pieces[ROOK] & p->colors[1]);
            int q = popcnt(nxt & p->pieces[QUEEN] & p->colors[0]) - popcnt(nxt & p->pieces[QUEEN] & p->colors[1]);
            mg += (r + q) * 5;
            eg += (r + q) * 5;
        }
        int nw = popcnt(fp & p->colors[0]), nb = popcnt(fp & p->colors[1]);
        mg += (nw - 1) * -10 - (nb - 1) * -10;
        eg += (nw - 1) * -20 - (nb - 1) * -20;
    }
    uint64_t wp = p->pieces[PAWN] & p->colors[0], bp = p->pieces[PAWN] & p->colors[1];
    int wc = popcnt((shift(wp & ~FILE_A, 7) | shift(wp & ~FILE_H, 9)) & wp);
    int bc = popcnt((shift(bp & ~FILE_A, 7) | shift(wp & ~FILE_H, 9)) & wp);
    mg += (wc - bc) * 4;
    eg += (wc - bc) * 4;
    int iso = popcnt(isolated(p, 0)) - popcnt(isolated(p, 1));
    mg += iso * -10;
    eg += iso * -10;
    int bw = popcnt(backward(p, 0)) - popcnt(backward(p, 1));
    mg += bw * -10;
    eg += bw * -10;
    static const int PH[5] = {0, 1, 1, 2, 4};
    int phase = 24;
    for (int t = 0; t < 5; t++) phase -= popcnt(p->pieces[t]) * PH[t];
    phase = (phase * 256) / 24;
    return (mg * (256 - phase) + eg * phase) / 256;
}
int ok_env_eval(const ok_env* e) { return evaluate(CUR(e)); }
float ok_env_bootstrap(const ok_env* e, float window) { /* env.h:476-484 */
    float s = (float)evaluate(CUR(e)) / window;
    if (s > 1.0f) s = 1.0f;
    if (s < -1.0f) s = -1.0f;
    return s;
}
uint64_t ok_env_key(const ok_env* e) { return CUR(e)->key; }
int ok_env_hmc(const ok_env* e) { return CUR(e)->hmc; }
int ok_env_check(const ok_env* e) { return CUR(e)->check; }
int ok_env_castle(const ok_env* e) { return CUR(e)->castle; }
int ok_env_ep(const ok_env* e) { return CUR(e)->ep; }
int ok_env_piece_at(const ok_env* e, int sq) { return CUR(e)->sq[sq]; }

void ok_env_export(const ok_env* e, void* out80) {
    const ok_pos* p = CUR(e);
    uint8_t* o = (uint8_t*)out80;
    memcpy(o, p->pieces, 48);
    memcpy(o + 48, &p->colors[0], 8);
    memcpy(o + 56, &p->board_key, 8);
    memcpy(o + 64, &p->key, 8);
    o[72] = (uint8_t)p->ctm;
    o[73] = (uint8_t)p->castle;
    o[74] = (uint8_t)(p->ep < 0 ? 0xFF : p->ep);
    o[75] = (uint8_t)p->hmc;
    uint16_t ply = (uint16_t)(e->n - 1);
    memcpy(o + 76, &ply, 2);
    o[78] = (uint8_t)(p->check ? 1 : 0);
    o[79] = 0;
}

/* ---- MCTS (kami/mcts.h) --------------------------------------------------------------- */
typedef struct ok_node {
    int n;
    float w, p;
    int action;
    float turn;
    struct ok_node* parent;
    struct ok_node** child;
    int nchild;
} ok_node;

struct ok_mcts {
    ok_env env;
    ok_node* root;
    ok_node* target;
    ok_mcts_cfg cfg;
    double cpuct;       /* mcts.h:70,89 */
    float fpu;          /* unvisited_node_value */
    float bw, bwindow, bamp;
    uint64_t rng;
};

static ok_node* node_new(void) {
    ok_node* n = (ok_node*)calloc(1, sizeof(ok_node));
    n->action = -1;
    return n;
}
static void node_free(ok_node* n) {
    for (int i = 0; i < n->nchild; i++) node_free(n->child[i]);
    free(n->child);
    free(n);
}
void ok_mcts_default_cfg(ok_mcts_cfg* c) {
    c->cpuct = 1.0f;
    c->force_expand_unvisited = 0;
    c->unvisited_node_value_pct = 100;
    c->bootstrap_weight = 0;
    c->bootstrap_window = 1600;
    c->bootstrap_amp_pct = 75;
    c->scale_cpuct_by_actions = 0;
    c->noise_weight = 0.05f;
    c->seed = 0;
}
ok_mcts* ok_mcts_new(const ok_mcts_cfg* c) { /* mcts.h:85-100 */
    ok_init();
    ok_mcts* t = (ok_mcts*)malloc(sizeof(ok_mcts));
    ok_env_reset(&t->env);
    t->cfg = *c;
    t->cpuct = (double)c->cpuct;
    t->fpu = (float)c->unvisited_node_value_pct / 100.0f;
    t->bw = (float)c->bootstrap_weight / 100.0f;
    t->bwindow = (float)c->bootstrap_window;
    t->bamp = (float)c->bootstrap_amp_pct / 100.0f;
    t->rng = c->seed * 0x9E3779B97F4A7C15ULL + 1;
    t->root = node_new();
    t->root->turn = -ok_env_turn(&t->env);
    t->target = NULL;
    return t;
}
void ok_mcts_free(ok_mcts* t) {
    node_free(t->root);
    free(t);
}
int ok_mcts_n(const ok_mcts* t) { return t->root->n; }
ok_env* ok_mcts_env(ok_mcts* t) { return &t->env; }

static void backprop(ok_node* nd, float value) { /* mcts.h:35-42 */
    for (; nd; nd = nd->parent) {
        nd->n += 1;
        volatile float half = (value * nd->turn) / 2.0f;
        volatile float inc = 0.5f + half;
        nd->w = nd->w + inc;
    }
}
static void unwind(ok_mcts* t) {
    while (t->target != t->root) {
        ok_env_pop(&t->env);
        t->target = t->target->parent;
    }
    t->target = NULL;
}
int ok_mcts_select(ok_mcts* t, float* obs) { /* mcts.h:186-255 (iterative) */
    if (!t->target) t->target = t->root;
    for (;;) {
        ok_node* tg = t->target;
        if (tg->nchild == 0) {
            float value;
            if (ok_env_terminal(&t->env, &value, NULL)) {
                backprop(tg, value);
                unwind(t);
                return 0;
            }
            ok_env_observe(&t->env, obs);
            return 1;
        }
        double best = -1000.0;
        ok_node* bc = NULL;
        float cpuct = (float)t->cpuct;
        if (t->cfg.scale_cpuct_by_actions) cpuct /= (float)tg->nchild;
        for (int i = 0; i < tg->nchild; i++) {
            ok_node* c = tg->child[i];
            if (t->cfg.force_expand_unvisited && !c->n) {
                bc = c;
                break;
            }
            volatile float fpu = t->fpu * c->turn;
            volatile float q = c->n > 0 ? c->w / (float)c->n : fpu;
            volatile float pc = c->p * cpuct;
            volatile double num = (double)pc * sqrt((double)tg->n);
            volatile double u = num / (double)(c->n + 1);
            double uct = (double)q + u;
            if (uct > best) {
                best = uct;
                bc = c;
            }
        }
        ok_env_push(&t->env, bc->action);
        t->target = bc;
    }
}
static double rng_u01(uint64_t* s) { /* splitmix64 */
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return ((double)(z >> 11) + 0.5) / 9007199254740992.0;
}
void ok_mcts_expand(ok_mcts* t, const float* policy, float value, int disable_bootstrap) { /* mcts.h:257-327 */
    ok_node* tg = t->target;
    int acts[OK_MAX_MOVES];
    int n = ok_env_actions(&t->env, acts, OK_MAX_MOVES);
    volatile float ptotal = 0.0f;
    for (int i = 0; i < n; i++) ptotal = ptotal + policy[acts[i]];
    float noise[OK_MAX_MOVES];
    volatile float tn = 0.0f;
    float nw = t->cfg.noise_weight;
    for (int i = 0; i < n; i++) {
        /* reference: gamma_distribution<double>(1,1) from a time-seeded mt19937, i.e. Exp(1);
         * not reproducible by design, so only nw == 0 is bit-comparable. */
        noise[i] = nw != 0.0f ? (float)(-log(rng_u01(&t->rng))) : 1.0f;
        tn = tn + noise[i];
    }
    tg->child = (ok_node**)malloc(sizeof(ok_node*) * (n > 0 ? n : 1));
    tg->nchild = n;
    for (int i = 0; i < n; i++) {
        ok_node* c = node_new();
        c->action = acts[i];
        c->parent = tg;
        c->turn = -tg->turn;
        volatile float a = (1 - nw) * policy[acts[i]];
        volatile float b = a / ptotal;
        volatile float d = noise[i] / tn;
        volatile float g = nw * d;
        c->p = b + g;
        tg->child[i] = c;
    }
    value *= tg->turn;
    if (!disable_bootstrap && t->bw > 0.0f) {
        volatile float a = (1 - t->bw) * value;
        volatile float b = t->bw * ok_env_bootstrap(&t->env, t->bwindow);
        volatile float c = b * t->bamp;
        value = a + c;
    }
    backprop(tg, value);
    unwind(t);
}
int ok_mcts_pick(ok_mcts* t, float alpha, double u01) { /* mcts.h:137-184 */
    ok_node* r = t->root;
    if (!r->nchild) return -2;
    if (alpha < 0.1f) {
        int bn = 0, ba = -1;
        for (int i = 0; i < r->nchild; i++)
            if (r->child[i]->n > bn) {
                bn = r->child[i]->n;
                ba = r->child[i]->action;
            }
        return ba;
    }
    double dist[OK_MAX_MOVES], len = 0.0;
    for (int i = 0; i < r->nchild; i++) {
        dist[i] = pow((double)r->child[i]->n, (double)(1.0f / alpha));
        len += dist[i];
    }
    double ind = u01;
    for (int i = 0; i < r->nchild; i++) {
        ind -= dist[i] / len;
        if (ind <= 0.0) return r->child[i]->action;
    }
    return r->child[r->nchild - 1]->action;
}
int ok_mcts_push(ok_mcts* t, int action) { /* mcts.h:113-135 */
    ok_node* r = t->root;
    ok_node* next = NULL;
    for (int i = 0; i < r->nchild; i++)
        if (r->child[i]->action == action) next = r->child[i];
    if (!next) return -1;
    for (int i = 0; i < r->nchild; i++)
        if (r->child[i] != next) node_free(r->child[i]);
    free(r->child);
    free(r);
    t->root = next;
    next->parent = NULL;
    ok_env_push(&t->env, action);
    return 0;
}
void ok_mcts_reset(ok_mcts* t) { /* mcts.h:331-339 */
    node_free(t->root);
    ok_env_reset(&t->env);
    t->target = NULL;
    t->root = node_new();
    t->root->turn = -ok_env_turn(&t->env);
}
void ok_mcts_snapshot(const ok_mcts* t, float* ps) { /* mcts.h:341-348 */
    for (int i = 0; i < OK_PSIZE; i++) ps[i] = 0.0f;
    for (int i = 0; i < t->root->nchild; i++)
        ps[t->root->child[i]->action] = (float)t->root->child[i]->n / (float)(t->root->n - 1);
}
int ok_mcts_root_children(const ok_mcts* t, int* action, int* n, float* w, float* p, int cap) {
    int k = t->root->nchild;
    for (int i = 0; i < k && i < cap; i++) {
        action[i] = t->root->child[i]->action;
        n[i] = t->root->child[i]->n;
        w[i] = t->root->child[i]->w;
        p[i] = t->root->child[i]->p;
    }
    return k;
}
float ok_mcts_root_w(const ok_mcts* t) { return t->root->w; }
static void digest(const ok_node* nd, uint64_t* h, long* cnt) {
    uint32_t wb, pb;
    memcpy(&wb, &nd->w, 4);
    memcpy(&pb, &nd->p, 4);
    uint64_t v[5] = {(uint32_t)nd->action, (uint32_t)nd->n, wb, pb, (uint64_t)nd->nchild};
    for (int i = 0; i < 5; i++) *h ^= v[i] + 0x9E3779B97F4A7C15ULL + (*h << 6) + (*h >> 2);
    ++*cnt;
    for (int i = 0; i < nd->nchild; i++) digest(nd->child[i], h, cnt);
}
uint64_t ok_mcts_digest(const ok_mcts* t, long* count) {
    uint64_t h = 0;
    long c = 0;
    digest(t->root, &h, &c);
    if (count) *count = c;
    return h;
}
