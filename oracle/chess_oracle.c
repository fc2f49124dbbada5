/*
 * chess_oracle.c -- TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
 *
 * A deliberately simple mailbox (8x8 array, ray loops, make-and-test legality) chess
 * rules engine that restates the subset of python-chess 1.10.0 semantics which the
 * reference hot path depends on.  python-chess (`chess==1.10.0`, reference
 * environment.yml:21) is an un-vendored third-party dependency that is not present in
 * /root/reference and cannot be installed here, so the published behaviour of that
 * library is restated from its documented algorithm; parity is anchored on
 *   - public perft tables (standard + Chess960), Scharnagl numbering anchors,
 *   - the reference's own call sites:
 *       chess_tensor.py:53   piece_at            chess_tensor.py:69   Board.from_chess960_pos
 *       chess_tensor.py:91   move in legal_moves  chess_tensor.py:95   push
 *       chess_tensor.py:101  is_repetition(2/3)   chess_tensor.py:113  move_stack
 *       chess_tensor.py:114  has_*_castling_rights chess_tensor.py:118 halfmove_clock
 *       chess_tensor.py:161  is_game_over/outcome  sim.py:46,86        is_game_over/result
 *       mcts.py:58           deepcopy(Board incl. move stack)
 * PARITY STATUS: "parity unpinned" by the reference (it ships no tests / golden vectors
 * for this boundary); pinned instead by the external known-answer tests in tests/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.
 *
 * Squares: 0 = a1 ... 7 = h1 ... 63 = h8 (python-chess numbering).
 * Pieces : 0 empty, 1..6 = white P N B R Q K, -1..-6 = black.
 * Moves  : from | to << 6 | promo << 12   (promo: 0 or 2..5 = N B R Q, python-chess piece types)
 *          Castling is written the way the board's mode writes it: vanilla boards use
 *          the two-square king move (e1g1), chess960 boards use king-takes-own-rook.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ctype.h>

#define WHITE 1
#define BLACK 0

typedef struct {
    int8_t  sq[64];
    uint8_t turn;        /* 1 = white to move */
    uint8_t rights[2];   /* [color] bitmask over files of castling rooks on that colour's back rank */
    int8_t  ep;          /* en-passant square set after ANY double pawn push, else -1 */
    uint16_t halfmove;
} OraPos;

typedef struct OraGame {
    int n;               /* number of moves played = len(move_stack) */
    int cap;
    int chess960;
    OraPos *stack;       /* stack[0] initial ... stack[n] current */
    uint16_t *moves;     /* moves[i]: stack[i] -> stack[i+1], in internal king-takes-rook form */
} OraGame;

static int color_of(int p) { return p > 0 ? WHITE : BLACK; }
static int backrank(int color) { return color == WHITE ? 0 : 7; }

/* ---------------------------------------------------------------- attacks */

static const int KN[8][2] = {{1,2},{2,1},{2,-1},{1,-2},{-1,-2},{-2,-1},{-2,1},{-1,2}};
static const int KG[8][2] = {{0,1},{1,1},{1,0},{1,-1},{0,-1},{-1,-1},{-1,0},{-1,1}};

/* is square s attacked by any piece of colour `by` on board b? */
static int attacked(const int8_t *b, int s, int by)
{
    int r = s >> 3, f = s & 7, i;
    int me = by == WHITE ? 1 : -1;
    /* pawns: a white pawn on (r-1, f+-1) attacks (r,f) */
    int pr = by == WHITE ? r - 1 : r + 1;
    if (pr >= 0 && pr < 8) {
        if (f > 0 && b[pr * 8 + f - 1] == me * 1) return 1;
        if (f < 7 && b[pr * 8 + f + 1] == me * 1) return 1;
    }
    for (i = 0; i < 8; i++) {
        int rr = r + KN[i][1], ff = f + KN[i][0];
        if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8 && b[rr * 8 + ff] == me * 2) return 1;
    }
    for (i = 0; i < 8; i++) {
        int rr = r + KG[i][1], ff = f + KG[i][0];
        if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8 && b[rr * 8 + ff] == me * 6) return 1;
    }
    for (i = 0; i < 8; i++) {
        int dr = KG[i][1], df = KG[i][0];
        int diag = dr != 0 && df != 0;
        int rr = r + dr, ff = f + df;
        while (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) {
            int p = b[rr * 8 + ff];
            if (p) {
                if (p == me * 5) return 1;
                if (diag && p == me * 3) return 1;
                if (!diag && p == me * 4) return 1;
                break;
            }
            rr += dr; ff += df;
        }
    }
    return 0;
}

static int king_square(const int8_t *b, int color)
{
    int k = color == WHITE ? 6 : -6, s;
    for (s = 0; s < 64; s++) if (b[s] == k) return s;
    return -1;
}

/* ---------------------------------------------------------------- make move */

#define MV(from, to, promo) ((uint16_t)((from) | ((to) << 6) | ((promo) << 12)))
#define MV_FROM(m) ((m) & 63)
#define MV_TO(m) (((m) >> 6) & 63)
#define MV_PROMO(m) (((m) >> 12) & 7)

static int is_castling_move(const OraPos *p, uint16_t m)
{
    int from = MV_FROM(m), to = MV_TO(m);
    int pc = p->sq[from];
    int me = p->turn == WHITE ? 1 : -1;
    return pc == me * 6 && p->sq[to] == me * 4;   /* king moves onto own rook */
}

/* python-chess Board.push semantics (castling given as king-takes-rook) */
static void make_move(const OraPos *p, uint16_t m, OraPos *o)
{
    int from = MV_FROM(m), to = MV_TO(m), promo = MV_PROMO(m);
    int us = p->turn, me = us == WHITE ? 1 : -1;
    int pc = p->sq[from], cap = p->sq[to];
    int old_ep = p->ep;
    *o = *p;
    o->ep = -1;
    o->halfmove = (uint16_t)(p->halfmove + 1);
    /* zeroing: pawn move or capture of an enemy piece */
    if (abs(pc) == 1 || (cap != 0 && color_of(cap) != us)) o->halfmove = 0;
    /* castling rights: any rook square touched as from/to loses its right */
    if ((from >> 3) == 0) o->rights[WHITE] &= (uint8_t)~(1u << (from & 7));
    if ((from >> 3) == 7) o->rights[BLACK] &= (uint8_t)~(1u << (from & 7));
    if ((to >> 3) == 0) o->rights[WHITE] &= (uint8_t)~(1u << (to & 7));
    if ((to >> 3) == 7) o->rights[BLACK] &= (uint8_t)~(1u << (to & 7));
    if (abs(pc) == 6) o->rights[us] = 0;

    if (pc == me * 6 && cap == me * 4) {            /* castling */
        int br = backrank(us) * 8;
        int a_side = (to & 7) < (from & 7);
        o->sq[from] = 0; o->sq[to] = 0;
        o->sq[br + (a_side ? 2 : 6)] = (int8_t)(me * 6);
        o->sq[br + (a_side ? 3 : 5)] = (int8_t)(me * 4);
    } else {
        o->sq[from] = 0;
        if (abs(pc) == 1) {
            int diff = to - from;
            if (diff == 16 && (from >> 3) == 1) o->ep = (int8_t)(from + 8);
            else if (diff == -16 && (from >> 3) == 6) o->ep = (int8_t)(from - 8);
            else if (to == old_ep && (abs(diff) == 7 || abs(diff) == 9) && cap == 0)
                o->sq[old_ep + (us == WHITE ? -8 : 8)] = 0;   /* en-passant capture */
        }
        o->sq[to] = (int8_t)(promo ? me * promo : pc);
    }
    o->turn = (uint8_t)!us;
}

/* ---------------------------------------------------------------- move generation */

static void add_move(uint16_t *out, int *n, int from, int to, int promo) { out[(*n)++] = MV(from, to, promo); }

static void add_pawn(uint16_t *out, int *n, int from, int to)
{
    int r = to >> 3;
    if (r == 0 || r == 7) {
        add_move(out, n, from, to, 5); add_move(out, n, from, to, 4);
        add_move(out, n, from, to, 3); add_move(out, n, from, to, 2);
    } else add_move(out, n, from, to, 0);
}

static int gen_pseudo(const OraPos *p, uint16_t *out)
{
    int n = 0, s, i;
    int us = p->turn, me = us == WHITE ? 1 : -1;
    for (s = 0; s < 64; s++) {
        int pc = p->sq[s], t, r = s >> 3, f = s & 7;
        if (pc == 0 || color_of(pc) != us) continue;
        t = abs(pc);
        if (t == 1) {
            int dir = us == WHITE ? 1 : -1, r1 = r + dir;
            if (r1 < 0 || r1 > 7) continue;
            if (p->sq[r1 * 8 + f] == 0) {
                add_pawn(out, &n, s, r1 * 8 + f);
                if (r == (us == WHITE ? 1 : 6) && p->sq[(r1 + dir) * 8 + f] == 0)
                    add_move(out, &n, s, (r1 + dir) * 8 + f, 0);
            }
            for (i = -1; i <= 1; i += 2) {
                int ff = f + i, to;
                if (ff < 0 || ff > 7) continue;
                to = r1 * 8 + ff;
                if (p->sq[to] != 0 && color_of(p->sq[to]) != us) add_pawn(out, &n, s, to);
                else if (p->sq[to] == 0 && to == p->ep && r == (us == WHITE ? 4 : 3))
                    add_move(out, &n, s, to, 0);
            }
        } else if (t == 2 || t == 6) {
            const int (*d)[2] = t == 2 ? KN : KG;
            for (i = 0; i < 8; i++) {
                int rr = r + d[i][1], ff = f + d[i][0], to;
                if (rr < 0 || rr > 7 || ff < 0 || ff > 7) continue;
                to = rr * 8 + ff;
                if (p->sq[to] == 0 || color_of(p->sq[to]) != us) add_move(out, &n, s, to, 0);
            }
        } else {
            for (i = 0; i < 8; i++) {
                int dr = KG[i][1], df = KG[i][0], diag = dr != 0 && df != 0, rr, ff;
                if (t == 3 && !diag) continue;
                if (t == 4 && diag) continue;
                rr = r + dr; ff = f + df;
                while (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) {
                    int to = rr * 8 + ff;
                    if (p->sq[to] == 0) add_move(out, &n, s, to, 0);
                    else { if (color_of(p->sq[to]) != us) add_move(out, &n, s, to, 0); break; }
                    rr += dr; ff += df;
                }
            }
        }
    }
    (void)me;
    return n;
}

/* python-chess generate_castling_moves: king-takes-rook form */
static int gen_castling(const OraPos *p, uint16_t *out)
{
    int n = 0, us = p->turn, me = us == WHITE ? 1 : -1, br = backrank(us) * 8;
    int k = king_square(p->sq, us), rf;
    if (k < 0 || (k >> 3) != backrank(us)) return 0;
    for (rf = 7; rf >= 0; rf--) {
        int rook, a_side, king_to, rook_to, lo, hi, s, ok = 1;
        int8_t b[64];
        if (!(p->rights[us] & (1u << rf))) continue;
        rook = br + rf;
        if (p->sq[rook] != me * 4) continue;
        a_side = rook < k;
        king_to = br + (a_side ? 2 : 6);
        rook_to = br + (a_side ? 3 : 5);
        /* squares the king/rook pass over, plus both destinations, must be empty once king and rook are lifted */
        memcpy(b, p->sq, 64);
        b[k] = 0; b[rook] = 0;
        lo = k < king_to ? k : king_to; hi = k < king_to ? king_to : k;
        for (s = lo; s <= hi; s++) if (b[s]) ok = 0;
        lo = rook < rook_to ? rook : rook_to; hi = rook < rook_to ? rook_to : rook;
        for (s = lo; s <= hi; s++) if (b[s]) ok = 0;
        if (!ok) continue;
        /* king start square and every square strictly between start and destination: unattacked with the king lifted */
        memcpy(b, p->sq, 64);
        b[k] = 0;
        lo = k < king_to ? k : king_to; hi = k < king_to ? king_to : k;
        for (s = lo; s <= hi; s++) {
            if (s == king_to && s != k) continue;       /* destination handled below */
            if (attacked(b, s, !us)) ok = 0;
        }
        if (!ok) continue;
        /* destination: unattacked with king and rook lifted and the rook standing on rook_to */
        b[rook] = 0;
        b[rook_to] = (int8_t)(me * 4);
        if (attacked(b, king_to, !us)) continue;
        add_move(out, &n, k, rook, 0);
    }
    return n;
}

static int gen_legal_internal(const OraPos *p, uint16_t *out)
{
    uint16_t tmp[512];
    int n = gen_pseudo(p, tmp), i, m = 0;
    for (i = 0; i < n; i++) {
        OraPos q;
        make_move(p, tmp[i], &q);
        if (!attacked(q.sq, king_square(q.sq, p->turn), !p->turn)) out[m++] = tmp[i];
    }
    m += gen_castling(p, out + m);
    return m;
}

static int has_legal_ep(const OraPos *p)
{
    uint16_t mv[512];
    int n, i;
    if (p->ep < 0) return 0;
    n = gen_legal_internal(p, mv);
    for (i = 0; i < n; i++)
        if (MV_TO(mv[i]) == p->ep && abs(p->sq[MV_FROM(mv[i])]) == 1 && (MV_FROM(mv[i]) & 7) != (p->ep & 7))
            return 1;
    return 0;
}

/* ---------------------------------------------------------------- board-mode move notation */

/* internal (king-takes-rook) -> the notation the board's mode uses */
static uint16_t to_external(const OraGame *g, const OraPos *p, uint16_t m)
{
    if (!g->chess960 && is_castling_move(p, m)) {
        int from = MV_FROM(m), to = MV_TO(m);
        if (from == 4 && to == 7) return MV(4, 6, 0);
        if (from == 4 && to == 0) return MV(4, 2, 0);
        if (from == 60 && to == 63) return MV(60, 62, 0);
        if (from == 60 && to == 56) return MV(60, 58, 0);
    }
    return m;
}

/* python-chess _to_chess960: e1g1 -> e1h1 when a king stands on e1 and no rook on g1 */
static uint16_t to_internal(const OraPos *p, uint16_t m)
{
    int from = MV_FROM(m), to = MV_TO(m);
    if (MV_PROMO(m)) return m;
    if (from == 4 && abs(p->sq[4]) == 6) {
        if (to == 6 && abs(p->sq[6]) != 4) return MV(4, 7, 0);
        if (to == 2 && abs(p->sq[2]) != 4) return MV(4, 0, 0);
    } else if (from == 60 && abs(p->sq[60]) == 6) {
        if (to == 62 && abs(p->sq[62]) != 4) return MV(60, 63, 0);
        if (to == 58 && abs(p->sq[58]) != 4) return MV(60, 56, 0);
    }
    return m;
}

/* ---------------------------------------------------------------- castling-right cleaning */

static void clean_rights(OraPos *p, int chess960)
{
    int c;
    for (c = 0; c < 2; c++) {
        int me = c == WHITE ? 1 : -1, br = backrank(c) * 8, f, k = -1;
        uint8_t r = 0;
        for (f = 0; f < 8; f++) if ((p->rights[c] & (1u << f)) && p->sq[br + f] == me * 4) r |= (uint8_t)(1u << f);
        for (f = 0; f < 8; f++) if (p->sq[br + f] == me * 6) k = f;
        if (!chess960) {
            r &= 0x81;
            if (k != 4) r = 0;
        } else {
            if (k < 0) r = 0;
            else {
                int a = -1, h = -1;
                for (f = 0; f < 8; f++) if (r & (1u << f)) { if (a < 0) a = f; h = f; }
                r = 0;
                if (a >= 0 && a < k) r |= (uint8_t)(1u << a);
                if (h >= 0 && h > k) r |= (uint8_t)(1u << h);
            }
        }
        p->rights[c] = r;
    }
}

/* ---------------------------------------------------------------- game object */

static OraGame *game_alloc(int chess960)
{
    OraGame *g = (OraGame *)calloc(1, sizeof(OraGame));
    g->cap = 64;
    g->chess960 = chess960;
    g->stack = (OraPos *)calloc((size_t)g->cap + 1, sizeof(OraPos));
    g->moves = (uint16_t *)calloc((size_t)g->cap + 1, sizeof(uint16_t));
    return g;
}

void ora_game_free(OraGame *g) { if (g) { free(g->stack); free(g->moves); free(g); } }

OraGame *ora_game_copy(const OraGame *s)
{
    OraGame *g = (OraGame *)calloc(1, sizeof(OraGame));
    *g = *s;
    g->stack = (OraPos *)calloc((size_t)g->cap + 1, sizeof(OraPos));
    g->moves = (uint16_t *)calloc((size_t)g->cap + 1, sizeof(uint16_t));
    memcpy(g->stack, s->stack, sizeof(OraPos) * (size_t)(s->n + 1));
    memcpy(g->moves, s->moves, sizeof(uint16_t) * (size_t)(s->n + 1));
    return g;
}

/* Scharnagl numbering of Chess960 start positions -> piece letters on files a..h */
void ora_chess960_backrank(int id, char *out8)
{
    static const int KNT[10][2] = {{0,1},{0,2},{0,3},{0,4},{1,2},{1,3},{1,4},{2,3},{2,4},{3,4}};
    char row[8];
    int n = id, i, k, free_idx;
    memset(row, 0, 8);
    row[(n % 4) * 2 + 1] = 'B'; n /= 4;       /* light-squared bishop: b d f h */
    row[(n % 4) * 2] = 'B'; n /= 4;           /* dark-squared bishop:  a c e g */
    k = n % 6; n /= 6;                        /* queen on the k-th free file */
    for (i = 0, free_idx = 0; i < 8; i++) if (!row[i]) { if (free_idx == k) { row[i] = 'Q'; break; } free_idx++; }
    for (i = 0, free_idx = 0; i < 8; i++) if (!row[i]) {
        if (free_idx == KNT[n][0] || free_idx == KNT[n][1]) row[i] = 'N';
        free_idx++;
    }
    for (i = 0, free_idx = 0; i < 8; i++) if (!row[i]) { row[i] = free_idx == 1 ? 'K' : 'R'; free_idx++; }
    memcpy(out8, row, 8);
}

static int piece_from_char(char c)
{
    const char *t = "PNBRQK";
    const char *q = strchr(t, toupper((unsigned char)c));
    int v;
    if (!q || !c) return 0;
    v = (int)(q - t) + 1;
    return isupper((unsigned char)c) ? v : -v;
}

OraGame *ora_game_new_startpos(int chess960_id)
{
    OraGame *g = game_alloc(chess960_id >= 0);
    OraPos *p = &g->stack[0];
    char row[8];
    int f;
    ora_chess960_backrank(chess960_id >= 0 ? chess960_id : 518, row);
    for (f = 0; f < 8; f++) {
        p->sq[f] = (int8_t)piece_from_char(row[f]);
        p->sq[56 + f] = (int8_t)-piece_from_char(row[f]);
        p->sq[8 + f] = 1; p->sq[48 + f] = -1;
        if (row[f] == 'R') { p->rights[WHITE] |= (uint8_t)(1u << f); p->rights[BLACK] |= (uint8_t)(1u << f); }
    }
    p->turn = WHITE; p->ep = -1; p->halfmove = 0;
    clean_rights(p, g->chess960);
    return g;
}

/* FEN / X-FEN / Shredder-FEN castling field */
OraGame *ora_game_new_fen(const char *fen, int chess960)
{
    OraGame *g = game_alloc(chess960);
    OraPos *p = &g->stack[0];
    int r = 7, f = 0;
    const char *c = fen;
    for (; *c && *c != ' '; c++) {
        if (*c == '/') { r--; f = 0; }
        else if (isdigit((unsigned char)*c)) f += *c - '0';
        else { if (r >= 0 && f < 8) p->sq[r * 8 + f] = (int8_t)piece_from_char(*c); f++; }
    }
    while (*c == ' ') c++;
    p->turn = (*c == 'w') ? WHITE : BLACK;
    while (*c && *c != ' ') c++;
    while (*c == ' ') c++;
    for (; *c && *c != ' '; c++) {
        int color, br, me, k, ff;
        if (*c == '-') continue;
        color = isupper((unsigned char)*c) ? WHITE : BLACK;
        br = backrank(color) * 8; me = color == WHITE ? 1 : -1;
        k = -1;
        for (ff = 0; ff < 8; ff++) if (p->sq[br + ff] == me * 6) k = ff;
        switch (tolower((unsigned char)*c)) {
        case 'k':   /* outermost rook on the h-side of the king */
            for (ff = 7; ff > k && k >= 0; ff--) if (p->sq[br + ff] == me * 4) { p->rights[color] |= (uint8_t)(1u << ff); break; }
            break;
        case 'q':
            for (ff = 0; ff < k; ff++) if (p->sq[br + ff] == me * 4) { p->rights[color] |= (uint8_t)(1u << ff); break; }
            break;
        default:
            if (tolower((unsigned char)*c) >= 'a' && tolower((unsigned char)*c) <= 'h')
                p->rights[color] |= (uint8_t)(1u << (tolower((unsigned char)*c) - 'a'));
        }
    }
    while (*c == ' ') c++;
    p->ep = -1;
    if (*c && *c != '-' && c[1]) p->ep = (int8_t)((c[1] - '1') * 8 + (c[0] - 'a'));
    while (*c && *c != ' ') c++;
    while (*c == ' ') c++;
    p->halfmove = (uint16_t)atoi(c);
    clean_rights(p, chess960);
    return g;
}

static const OraPos *cur(const OraGame *g) { return &g->stack[g->n]; }

int ora_turn(const OraGame *g) { return cur(g)->turn; }
int ora_ply(const OraGame *g) { return g->n; }
int ora_halfmove(const OraGame *g) { return cur(g)->halfmove; }
int ora_ep_square(const OraGame *g) { return cur(g)->ep; }
int ora_is_chess960(const OraGame *g) { return g->chess960; }
int ora_piece_at(const OraGame *g, int s) { return cur(g)->sq[s]; }
int ora_rights(const OraGame *g, int color) { return cur(g)->rights[color]; }
void ora_board(const OraGame *g, int8_t *out64) { memcpy(out64, cur(g)->sq, 64); }

int ora_legal_moves(const OraGame *g, uint16_t *out)
{
    const OraPos *p = cur(g);
    int n = gen_legal_internal(p, out), i;
    for (i = 0; i < n; i++) out[i] = to_external(g, p, out[i]);
    return n;
}

int ora_is_legal(const OraGame *g, uint16_t m)
{
    uint16_t mv[512];
    const OraPos *p = cur(g);
    uint16_t mi = to_internal(p, m);
    int n = gen_legal_internal(p, mv), i;
    for (i = 0; i < n; i++) if (mv[i] == mi) return 1;
    return 0;
}

int ora_move_at(const OraGame *g, int i)
{
    if (i < 0) i += g->n;
    if (i < 0 || i >= g->n) return -1;
    return to_external(g, &g->stack[i], g->moves[i]);
}

void ora_push(OraGame *g, uint16_t m)
{
    OraPos nxt;
    if (g->n + 2 > g->cap) {
        g->cap *= 2;
        g->stack = (OraPos *)realloc(g->stack, sizeof(OraPos) * ((size_t)g->cap + 1));
        g->moves = (uint16_t *)realloc(g->moves, sizeof(uint16_t) * ((size_t)g->cap + 1));
    }
    if (g->n == 0) clean_rights(&g->stack[0], g->chess960);
    m = to_internal(cur(g), m);
    make_move(cur(g), m, &nxt);
    g->moves[g->n] = m;
    g->stack[++g->n] = nxt;
}

int ora_pop(OraGame *g) { if (g->n == 0) return -1; g->n--; return 0; }

int ora_is_check(const OraGame *g)
{
    const OraPos *p = cur(g);
    int k = king_square(p->sq, p->turn);
    return k >= 0 && attacked(p->sq, k, !p->turn);
}

/* Board.has_{king,queen}side_castling_rights */
int ora_has_castling(const OraGame *g, int color, int kingside)
{
    const OraPos *p = cur(g);
    int me = color == WHITE ? 1 : -1, br = backrank(color) * 8, k = -1, f;
    for (f = 0; f < 8; f++) if (p->sq[br + f] == me * 6) k = f;
    if (k < 0) return 0;
    for (f = 0; f < 8; f++)
        if ((p->rights[color] & (1u << f)) && (kingside ? f > k : f < k)) return 1;
    return 0;
}

/* move i (stack[i] -> stack[i+1]) judged on the pre-move board: Board.is_irreversible */
static int irreversible(const OraGame *g, int i)
{
    const OraPos *p = &g->stack[i];
    uint16_t m = g->moves[i];
    int from = MV_FROM(m), to = MV_TO(m), us = p->turn;
    OraPos q;
    if (abs(p->sq[from]) == 1 || abs(p->sq[to]) == 1) return 1;              /* touched & pawns */
    if (p->sq[to] != 0 && color_of(p->sq[to]) != us) return 1;               /* capture */
    make_move(p, m, &q);
    if (q.rights[0] != p->rights[0] || q.rights[1] != p->rights[1]) return 1; /* reduces castling rights */
    if (has_legal_ep(p)) return 1;
    return 0;
}

static int same_position(const OraPos *a, const OraPos *b)
{
    int ea, eb;
    if (memcmp(a->sq, b->sq, 64) || a->turn != b->turn) return 0;
    if (a->rights[0] != b->rights[0] || a->rights[1] != b->rights[1]) return 0;
    ea = has_legal_ep(a) ? a->ep : -1;
    eb = has_legal_ep(b) ? b->ep : -1;
    return ea == eb;
}

/* Board.is_repetition(count) */
int ora_is_repetition(const OraGame *g, int count)
{
    const OraPos *now = cur(g);
    int i = g->n;          /* index of the position currently "on the board" while popping */
    for (;;) {
        if (count <= 1) return 1;
        if (i < count - 1) break;          /* len(move_stack) < count - 1 */
        i--;                                /* pop */
        if (irreversible(g, i)) break;
        if (same_position(&g->stack[i], now)) count--;
    }
    return 0;
}

static int insufficient_side(const OraPos *p, int color)
{
    int s, own = 0, pawns_rq = 0, knights = 0, bishops = 0;
    int opp_other = 0;       /* opponent pieces that are not king / queen */
    int all_b_light = 0, all_b_dark = 0, any_pawn = 0, any_knight = 0;
    for (s = 0; s < 64; s++) {
        int pc = p->sq[s], t = abs(pc);
        if (!pc) continue;
        if (t == 1) any_pawn = 1;
        if (t == 2) any_knight = 1;
        if (t == 3) { if (((s >> 3) + (s & 7)) & 1) all_b_light = 1; else all_b_dark = 1; }
        if (color_of(pc) == color) {
            own++;
            if (t == 1 || t == 4 || t == 5) pawns_rq = 1;
            if (t == 2) knights = 1;
            if (t == 3) bishops = 1;
        } else if (t != 6 && t != 5) opp_other = 1;
    }
    if (pawns_rq) return 0;
    if (knights) return own <= 2 && !opp_other;
    if (bishops) return !(all_b_light && all_b_dark) && !any_pawn && !any_knight;
    return 1;
}

/* Board.outcome(claim_draw=False): 0 none, 1 checkmate, 2 insufficient material, 3 stalemate,
 * 4 seventy-five moves, 5 fivefold repetition */
int ora_outcome(const OraGame *g)
{
    uint16_t mv[512];
    const OraPos *p = cur(g);
    int n = gen_legal_internal(p, mv);
    if (n == 0 && ora_is_check(g)) return 1;
    if (insufficient_side(p, WHITE) && insufficient_side(p, BLACK)) return 2;
    if (n == 0) return 3;
    if (p->halfmove >= 150) return 4;
    if (ora_is_repetition(g, 5)) return 5;
    return 0;
}

static uint64_t perft_rec(const OraPos *p, int depth)
{
    uint16_t mv[512];
    int n = gen_legal_internal(p, mv), i;
    uint64_t t = 0;
    if (depth <= 1) return (uint64_t)n;
    for (i = 0; i < n; i++) {
        OraPos q;
        make_move(p, mv[i], &q);
        t += perft_rec(&q, depth - 1);
    }
    return t;
}

uint64_t ora_perft(const OraGame *g, int depth)
{
    if (depth <= 0) return 1;
    return perft_rec(cur(g), depth);
}

/* per-root-move breakdown for debugging the CUDA engine: returns count, fills moves (external) and nodes */
int ora_divide(const OraGame *g, int depth, uint16_t *moves, uint64_t *nodes)
{
    uint16_t mv[512];
    const OraPos *p = cur(g);
    int n = gen_legal_internal(p, mv), i;
    for (i = 0; i < n; i++) {
        OraPos q;
        make_move(p, mv[i], &q);
        moves[i] = to_external(g, p, mv[i]);
        nodes[i] = depth <= 1 ? 1 : perft_rec(&q, depth - 1);
    }
    return n;
}

/* Export current position in the product's szb_pos wire layout (see include/szb200.h):
 * 12 piece bitboards (white P N B R Q K, black P N B R Q K), then scalar fields. */
void ora_export(const OraGame *g, uint64_t *bb12, int *turn, int *rights_w, int *rights_b,
                int *ep, int *halfmove, int *ply)
{
    const OraPos *p = cur(g);
    int s;
    memset(bb12, 0, 12 * sizeof(uint64_t));
    for (s = 0; s < 64; s++) {
        int pc = p->sq[s];
        if (pc > 0) bb12[pc - 1] |= 1ull << s;
        else if (pc < 0) bb12[6 + (-pc) - 1] |= 1ull << s;
    }
    *turn = p->turn; *rights_w = p->rights[WHITE]; *rights_b = p->rights[BLACK];
    *ep = p->ep; *halfmove = p->halfmove; *ply = g->n;
}
