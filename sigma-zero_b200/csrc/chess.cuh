// chess.cuh -- bitboard chess rules for the GPU self-play engine (one thread per position).
//
// Replaces the python-chess calls the reference makes on its hot path:
//   chess_tensor.py:91   `move in board.legal_moves`      -> gen_legal / index_to_move
//   chess_tensor.py:95   `board.push(move)`               -> make_move
//   chess_tensor.py:101  `board.is_repetition(2|3)`       -> count_repetitions (ancestor walk)
//   chess_tensor.py:114  `has_*_castling_rights`          -> castling_flags
//   chess_tensor.py:161  `board.is_game_over()/outcome()` -> outcome_of
//   chess_tensor.py:221  actionToTensor                   -> move_to_index
//   chess_tensor.py:309  tensorToAction                   -> index_to_move
//   chess_tensor.py:123  plane stacking + :131 flips      -> pack_planes
//
// Everything here is `__host__ __device__` so that the very same code can be compiled with g++ and
// checked against the oracle on a machine without a GPU (tests/host_harness).  The product only ever
// runs it on the device.
//
// Conventions: square 0 = a1 .. 63 = h8; bit s of a bitboard = square s.
// Sliding attacks use reversed-subtraction on line masks (o - s) ^ rev(rev(o) - rev(s)); the line masks
// and the knight/king step tables live in a 2 KB `Tables` block that kernels stage in shared memory.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SZB_HD __host__ __device__ __forceinline__
#define SZB_HDN inline __host__ __device__ __noinline__
#else
#define SZB_HD inline
#define SZB_HDN inline
#endif

namespace szb {

// ----------------------------------------------------------------------------------------------
// bit helpers
// ----------------------------------------------------------------------------------------------
SZB_HD int lsb(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
SZB_HD int popc(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
SZB_HD uint64_t bswap64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
#else
    return __builtin_bswap64(x);
#endif
}
SZB_HD uint64_t brev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    return __builtin_bswap64(x);
#endif
}
SZB_HD uint64_t bit(int s) { return 1ull << s; }

constexpr uint64_t FILE_A = 0x0101010101010101ull;
constexpr uint64_t FILE_H = 0x8080808080808080ull;
constexpr uint64_t RANK_1 = 0x00000000000000FFull;
constexpr uint64_t RANK_8 = 0xFF00000000000000ull;
constexpr uint64_t DARK_SQUARES = 0xAA55AA55AA55AA55ull;

// piece-type bitboard slots inside Pos::bb
enum { BB_WHITE = 0, BB_BLACK = 1, BB_P = 2, BB_N = 3, BB_B = 4, BB_R = 5, BB_Q = 6, BB_K = 7 };
// python-chess piece types (promotion codes)
enum { PT_PAWN = 1, PT_KNIGHT = 2, PT_BISHOP = 3, PT_ROOK = 4, PT_QUEEN = 5, PT_KING = 6 };

enum : uint8_t {
    F_WHITE = 1,      // white to move
    F_REP2 = 2,       // is_repetition(2) held when this position was reached
    F_REP3 = 4,       // is_repetition(3)
    F_IRREV = 8,      // the move that led here was irreversible as judged on the previous board
    F_EPLEGAL = 16,   // a legal en-passant capture exists here (part of the transposition key)
    F_960 = 32,       // game object was created in chess960 mode (castling written king-takes-rook)
    F_REP5 = 64,      // is_repetition(5)
};

enum : uint8_t { OUT_NONE = 0, OUT_CHECKMATE = 1, OUT_INSUFFICIENT = 2, OUT_STALEMATE = 3, OUT_75MOVES = 4, OUT_FIVEFOLD = 5 };

constexpr uint32_t NO_PREV = 0xFFFFFFFFu;
constexpr int MAX_MOVES = 224;          // 218 is the known maximum of legal moves in a chess position
constexpr int N_PLANES = 119;
constexpr int N_ACTIONS = 4672;
constexpr int MASK_WORDS = 73;          // 4672 / 64

// One position as the engine stores it in HBM: 96 bytes, six 16-byte vectors.
struct alignas(16) Pos {
    uint64_t bb[8];      // white, black, P, N, B, R, Q, K
    uint64_t key;        // hash of python-chess's transposition key (pieces, turn, clean rights, legal-ep square)
    uint32_t prev;       // slot of the previous position inside this game's state pool, NO_PREV at game start
    uint16_t ply;        // len(board.move_stack)
    uint8_t rights_w;    // files of castling rooks on rank 1
    uint8_t rights_b;    // ... rank 8
    int8_t ep;           // square behind the last double pawn push, -1 otherwise
    uint8_t halfmove;    // halfmove clock (game ends at 150)
    uint8_t flags;       // F_*
    uint8_t outcome;     // OUT_* (filled by analyse)
    uint8_t n_legal;     // number of legal moves (filled by analyse)
    uint8_t pad[3];
};
static_assert(sizeof(Pos) == 96, "Pos must stay 96 bytes");

struct Tables {
    uint64_t knight[64];
    uint64_t king[64];
    uint64_t diag[64];   // a1-h8 direction, square itself excluded
    uint64_t anti[64];   // h1-a8 direction, square itself excluded
};

inline void build_tables(Tables& t) {
    for (int s = 0; s < 64; s++) {
        int r = s >> 3, f = s & 7;
        uint64_t kn = 0, kg = 0, dg = 0, an = 0;
        const int kd[8][2] = {{1, 2}, {2, 1}, {2, -1}, {1, -2}, {-1, -2}, {-2, -1}, {-2, 1}, {-1, 2}};
        for (int i = 0; i < 8; i++) {
            int rr = r + kd[i][1], ff = f + kd[i][0];
            if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) kn |= 1ull << (rr * 8 + ff);
        }
        for (int dr = -1; dr <= 1; dr++)
            for (int df = -1; df <= 1; df++) {
                int rr = r + dr, ff = f + df;
                if ((dr || df) && rr >= 0 && rr < 8 && ff >= 0 && ff < 8) kg |= 1ull << (rr * 8 + ff);
            }
        for (int q = 0; q < 64; q++) {
            if (q == s) continue;
            int qr = q >> 3, qf = q & 7;
            if (qr - qf == r - f) dg |= 1ull << q;
            if (qr + qf == r + f) an |= 1ull << q;
        }
        t.knight[s] = kn; t.king[s] = kg; t.diag[s] = dg; t.anti[s] = an;
    }
}

// ----------------------------------------------------------------------------------------------
// attacks
// ----------------------------------------------------------------------------------------------
SZB_HD uint64_t rank_mask(int s) { return (0xFFull << (s & 56)) ^ bit(s); }
SZB_HD uint64_t file_mask(int s) { return (FILE_A << (s & 7)) ^ bit(s); }

SZB_HD uint64_t line_att(uint64_t occ, int s, uint64_t mask) {
    uint64_t o = occ & mask;
    uint64_t f = o - bit(s);
    uint64_t r = brev64(brev64(o) - bit(63 - s));
    return (f ^ r) & mask;
}
SZB_HD uint64_t rook_att(int s, uint64_t occ) { return line_att(occ, s, rank_mask(s)) | line_att(occ, s, file_mask(s)); }
SZB_HD uint64_t bishop_att(const Tables& T, int s, uint64_t occ) {
    return line_att(occ, s, T.diag[s]) | line_att(occ, s, T.anti[s]);
}

// squares strictly between a and b when they share a line, else 0
SZB_HD uint64_t between(const Tables& T, int a, int b) {
    if (a == b) return 0;
    uint64_t ba = bit(a), bbt = bit(b), ma;
    if ((a >> 3) == (b >> 3)) ma = rank_mask(a);
    else if ((a & 7) == (b & 7)) ma = file_mask(a);
    else if (T.diag[a] & bbt) ma = T.diag[a];
    else if (T.anti[a] & bbt) ma = T.anti[a];
    else return 0;
    uint64_t mb = (ma | ba) & ~bbt;
    return line_att(bbt, a, ma) & line_att(ba, b, mb);
}
// the whole line through a and b (both included) when aligned, else 0
SZB_HD uint64_t line_through(const Tables& T, int a, int b) {
    uint64_t bbt = bit(b);
    if ((a >> 3) == (b >> 3)) return rank_mask(a) | bit(a);
    if ((a & 7) == (b & 7)) return file_mask(a) | bit(a);
    if (T.diag[a] & bbt) return T.diag[a] | bit(a);
    if (T.anti[a] & bbt) return T.anti[a] | bit(a);
    return 0;
}

// squares from which a white / black pawn attacks the squares in b
SZB_HD uint64_t wpawn_sources(uint64_t b) { return ((b >> 7) & ~FILE_A) | ((b >> 9) & ~FILE_H); }
SZB_HD uint64_t bpawn_sources(uint64_t b) { return ((b << 7) & ~FILE_H) | ((b << 9) & ~FILE_A); }

// all pieces (either colour) attacking square s under occupancy occ
SZB_HD uint64_t attackers_to(const Tables& T, const uint64_t* bb, int s, uint64_t occ) {
    uint64_t b = bit(s);
    return (wpawn_sources(b) & bb[BB_P] & bb[BB_WHITE]) | (bpawn_sources(b) & bb[BB_P] & bb[BB_BLACK]) |
           (T.knight[s] & bb[BB_N]) | (T.king[s] & bb[BB_K]) |
           (rook_att(s, occ) & (bb[BB_R] | bb[BB_Q])) | (bishop_att(T, s, occ) & (bb[BB_B] | bb[BB_Q]));
}

SZB_HD int piece_type_at(const uint64_t* bb, int s) {
    uint64_t b = bit(s);
    for (int t = 0; t < 6; t++)
        if (bb[BB_P + t] & b) return t + 1;
    return 0;
}

// ----------------------------------------------------------------------------------------------
// moves: from | to << 6 | promo << 12.  Castling is always king-takes-own-rook internally.
// ----------------------------------------------------------------------------------------------
SZB_HD uint16_t mk_move(int from, int to, int promo) { return (uint16_t)(from | (to << 6) | (promo << 12)); }
SZB_HD int mv_from(uint16_t m) { return m & 63; }
SZB_HD int mv_to(uint16_t m) { return (m >> 6) & 63; }
SZB_HD int mv_promo(uint16_t m) { return (m >> 12) & 7; }
constexpr uint16_t MOVE_NONE = 0xFFFF;

SZB_HD void emit_targets(uint16_t* out, int& n, int from, uint64_t targets) {
    while (targets) {
        int t = lsb(targets);
        targets &= targets - 1;
        out[n++] = mk_move(from, t, 0);
    }
}

// Legal moves (python-chess Board.legal_moves as a set).  Returns the count.
SZB_HDN int gen_legal(const Tables& T, const Pos& p, uint16_t* out) {
    const bool white = p.flags & F_WHITE;
    const uint64_t own = p.bb[white ? BB_WHITE : BB_BLACK], opp = p.bb[white ? BB_BLACK : BB_WHITE];
    const uint64_t occ = own | opp;
    const uint64_t P = p.bb[BB_P], N = p.bb[BB_N], B = p.bb[BB_B], R = p.bb[BB_R], Q = p.bb[BB_Q], K = p.bb[BB_K];
    int n = 0;
    const uint64_t kbb = K & own;
    if (!kbb) return 0;
    const int ksq = lsb(kbb);
    const uint64_t checkers = attackers_to(T, p.bb, ksq, occ) & opp;

    // king steps: target must be unattacked once the king has left its square
    {
        uint64_t kt = T.king[ksq] & ~own;
        const uint64_t occ_nk = occ ^ kbb;
        while (kt) {
            int t = lsb(kt);
            kt &= kt - 1;
            if (!(attackers_to(T, p.bb, t, occ_nk) & opp)) out[n++] = mk_move(ksq, t, 0);
        }
    }
    if (checkers & (checkers - 1)) return n;                  // double check: king moves only

    uint64_t target_mask = ~own;
    if (checkers) target_mask &= between(T, ksq, lsb(checkers)) | checkers;

    // absolutely pinned own pieces
    uint64_t pinned = 0;
    {
        uint64_t snipers = (((rank_mask(ksq) | file_mask(ksq)) & (R | Q)) | ((T.diag[ksq] | T.anti[ksq]) & (B | Q))) & opp;
        while (snipers) {
            int s = lsb(snipers);
            snipers &= snipers - 1;
            uint64_t b = between(T, ksq, s) & occ;
            if (b && !(b & (b - 1)) && (b & own)) pinned |= b;
        }
    }

    // knights (a pinned knight has no moves)
    {
        uint64_t x = N & own & ~pinned;
        while (x) {
            int s = lsb(x);
            x &= x - 1;
            emit_targets(out, n, s, T.knight[s] & target_mask);
        }
    }
    // sliders
    {
        uint64_t x = (B | R | Q) & own;
        while (x) {
            int s = lsb(x);
            x &= x - 1;
            uint64_t a = 0, sb = bit(s);
            if ((B | Q) & sb) a |= bishop_att(T, s, occ);
            if ((R | Q) & sb) a |= rook_att(s, occ);
            a &= target_mask;
            if (pinned & sb) a &= line_through(T, ksq, s);
            emit_targets(out, n, s, a);
        }
    }
    // pawns
    {
        uint64_t x = P & own;
        const uint64_t last = white ? RANK_8 : RANK_1;
        while (x) {
            int s = lsb(x);
            x &= x - 1;
            uint64_t sb = bit(s), a = 0;
            if (white) {
                uint64_t one = (sb << 8) & ~occ;
                a = one | (((one << 8) & ~occ) & 0x00000000FF000000ull);
                a |= (((sb << 7) & ~FILE_H) | ((sb << 9) & ~FILE_A)) & opp;
            } else {
                uint64_t one = (sb >> 8) & ~occ;
                a = one | (((one >> 8) & ~occ) & 0x000000FF00000000ull);
                a |= (((sb >> 7) & ~FILE_A) | ((sb >> 9) & ~FILE_H)) & opp;
            }
            a &= target_mask;
            if (pinned & sb) a &= line_through(T, ksq, s);
            while (a) {
                int t = lsb(a);
                a &= a - 1;
                if (bit(t) & last) {
                    out[n++] = mk_move(s, t, PT_QUEEN);
                    out[n++] = mk_move(s, t, PT_ROOK);
                    out[n++] = mk_move(s, t, PT_BISHOP);
                    out[n++] = mk_move(s, t, PT_KNIGHT);
                } else out[n++] = mk_move(s, t, 0);
            }
        }
    }
    // en passant: make the capture on the occupancy and look at the king
    if (p.ep >= 0 && !(occ & bit(p.ep))) {
        const uint64_t eb = bit(p.ep);
        uint64_t caps = (white ? wpawn_sources(eb) : bpawn_sources(eb)) & P & own & (white ? 0x000000FF00000000ull : 0x00000000FF000000ull);
        const uint64_t victim = white ? (eb >> 8) : (eb << 8);
        while (caps) {
            int s = lsb(caps);
            caps &= caps - 1;
            uint64_t occ2 = (occ ^ bit(s) ^ victim) | eb;
            if (!(attackers_to(T, p.bb, ksq, occ2) & opp & ~victim)) out[n++] = mk_move(s, p.ep, 0);
        }
    }
    // castling (python-chess generate_castling_moves)
    {
        const uint8_t rights = white ? p.rights_w : p.rights_b;
        const int br = white ? 0 : 56;
        if (rights && (kbb & (0xFFull << br))) {
            for (int rf = 0; rf < 8; rf++) {
                if (!(rights & (1u << rf))) continue;
                const int rook = br + rf;
                const uint64_t rb = bit(rook);
                if (!(R & own & rb)) continue;
                const bool a_side = rook < ksq;
                const int kto = br + (a_side ? 2 : 6), rto = br + (a_side ? 3 : 5);
                const uint64_t kpath = between(T, ksq, kto);
                const uint64_t must_empty = kpath | between(T, rook, rto) | bit(kto) | bit(rto);
                if ((occ ^ kbb ^ rb) & must_empty) continue;
                uint64_t chk = kpath | kbb;
                bool bad = false;
                while (chk && !bad) {
                    int s = lsb(chk);
                    chk &= chk - 1;
                    bad = (attackers_to(T, p.bb, s, occ ^ kbb) & opp) != 0;
                }
                if (bad) continue;
                if (attackers_to(T, p.bb, kto, occ ^ kbb ^ rb ^ bit(rto)) & opp) continue;
                out[n++] = mk_move(ksq, rook, 0);
            }
        }
    }
    return n;
}

SZB_HD bool in_check(const Tables& T, const Pos& p) {
    const bool white = p.flags & F_WHITE;
    const uint64_t own = p.bb[white ? BB_WHITE : BB_BLACK], opp = p.bb[white ? BB_BLACK : BB_WHITE];
    const uint64_t kbb = p.bb[BB_K] & own;
    if (!kbb) return false;
    return (attackers_to(T, p.bb, lsb(kbb), own | opp) & opp) != 0;
}

// Board.has_legal_en_passant
SZB_HD bool has_legal_ep(const Tables& T, const Pos& p) {
    if (p.ep < 0) return false;
    const bool white = p.flags & F_WHITE;
    const uint64_t own = p.bb[white ? BB_WHITE : BB_BLACK], opp = p.bb[white ? BB_BLACK : BB_WHITE];
    const uint64_t occ = own | opp, eb = bit(p.ep);
    if (occ & eb) return false;
    const uint64_t kbb = p.bb[BB_K] & own;
    if (!kbb) return false;
    const int ksq = lsb(kbb);
    uint64_t caps = (white ? wpawn_sources(eb) : bpawn_sources(eb)) & p.bb[BB_P] & own & (white ? 0x000000FF00000000ull : 0x00000000FF000000ull);
    const uint64_t victim = white ? (eb >> 8) : (eb << 8);
    while (caps) {
        int s = lsb(caps);
        caps &= caps - 1;
        uint64_t occ2 = (occ ^ bit(s) ^ victim) | eb;
        if (!(attackers_to(T, p.bb, ksq, occ2) & opp & ~victim)) return true;
    }
    return false;
}

// ----------------------------------------------------------------------------------------------
// transposition key + equality (python-chess Board._transposition_key)
// ----------------------------------------------------------------------------------------------
SZB_HD uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
SZB_HD int key_ep(const Pos& p) { return (p.flags & F_EPLEGAL) ? p.ep : -1; }
SZB_HD uint64_t position_key(const Pos& p) {
    uint64_t h = 0x243F6A8885A308D3ull;
#pragma unroll
    for (int i = 0; i < 8; i++) h = mix64((h ^ p.bb[i]) + 0x9E3779B97F4A7C15ull);
    uint64_t tail = (uint64_t)(p.flags & F_WHITE) | ((uint64_t)p.rights_w << 8) | ((uint64_t)p.rights_b << 16) |
                    ((uint64_t)(uint8_t)key_ep(p) << 24);
    return mix64(h ^ tail);
}
SZB_HD bool same_position(const Pos& a, const Pos& b) {
    if (a.key != b.key) return false;
    bool eq = true;
#pragma unroll
    for (int i = 0; i < 8; i++) eq &= a.bb[i] == b.bb[i];
    return eq && ((a.flags ^ b.flags) & F_WHITE) == 0 && a.rights_w == b.rights_w && a.rights_b == b.rights_b &&
           key_ep(a) == key_ep(b);
}

// ----------------------------------------------------------------------------------------------
// make move (python-chess Board.push).  Fills everything except prev / repetition flags / outcome.
// ----------------------------------------------------------------------------------------------
SZB_HDN void make_move(const Tables& T, const Pos& p, uint16_t m, Pos& q) {
    const int from = mv_from(m), to = mv_to(m), promo = mv_promo(m);
    const bool white = p.flags & F_WHITE;
    const int us = white ? BB_WHITE : BB_BLACK, them = white ? BB_BLACK : BB_WHITE;
    const uint64_t fb = bit(from), tb = bit(to);
    q = p;
    const int pt = piece_type_at(p.bb, from);
    const bool capture = (p.bb[them] & tb) != 0;
    const bool zeroing = pt == PT_PAWN || capture;
    q.ep = -1;
    q.halfmove = zeroing ? 0 : (uint8_t)(p.halfmove + 1);
    // castling rights
    uint8_t rw = p.rights_w, rb = p.rights_b;
    if (from < 8) rw &= (uint8_t)~(1u << from);
    if (to < 8) rw &= (uint8_t)~(1u << to);
    if (from >= 56) rb &= (uint8_t)~(1u << (from - 56));
    if (to >= 56) rb &= (uint8_t)~(1u << (to - 56));
    if (pt == PT_KING) { if (white) rw = 0; else rb = 0; }
    q.rights_w = rw; q.rights_b = rb;

    if (pt == PT_KING && (p.bb[us] & tb)) {                 // castling: king onto own rook
        const int br = white ? 0 : 56;
        const bool a_side = to < from;
        const uint64_t kto = bit(br + (a_side ? 2 : 6)), rto = bit(br + (a_side ? 3 : 5));
        q.bb[BB_K] = (q.bb[BB_K] & ~fb) | kto;
        q.bb[BB_R] = (q.bb[BB_R] & ~tb) | rto;
        q.bb[us] = (q.bb[us] & ~(fb | tb)) | kto | rto;
    } else {
        q.bb[BB_P + pt - 1] &= ~fb;
        q.bb[us] &= ~fb;
        if (capture) {
#pragma unroll
            for (int t = 0; t < 6; t++) q.bb[BB_P + t] &= ~tb;
            q.bb[them] &= ~tb;
        }
        if (pt == PT_PAWN) {
            const int diff = to - from;
            if (diff == 16 && (from >> 3) == 1) q.ep = (int8_t)(from + 8);
            else if (diff == -16 && (from >> 3) == 6) q.ep = (int8_t)(from - 8);
            else if (to == p.ep && (diff == 7 || diff == 9 || diff == -7 || diff == -9) && !capture) {
                const uint64_t victim = white ? (tb >> 8) : (tb << 8);
                q.bb[BB_P] &= ~victim;
                q.bb[them] &= ~victim;
            }
        }
        q.bb[BB_P + (promo ? promo : pt) - 1] |= tb;
        q.bb[us] |= tb;
    }
    uint8_t fl = (uint8_t)((p.flags & F_960) | (white ? 0 : F_WHITE));
    if (zeroing || rw != p.rights_w || rb != p.rights_b || (p.flags & F_EPLEGAL)) fl |= F_IRREV;
    q.flags = fl;
    if (has_legal_ep(T, q)) q.flags |= F_EPLEGAL;
    q.ply = (uint16_t)(p.ply + 1);
    q.key = position_key(q);
    q.outcome = OUT_NONE;
    q.n_legal = 0;
}

// Number of earlier positions equal to pool[idx] reachable by walking back over reversible moves, capped
// at 4 (is_repetition(k) == result >= k-1).  The walk follows Pos::prev through this game's pool.
SZB_HD int count_repetitions(const Pos* pool, const Pos& now) {
    int matches = 0;
    const Pos* cur = &now;
    for (int guard = 0; guard < 512; guard++) {
        if (cur->flags & F_IRREV) break;
        if (cur->prev == NO_PREV) break;
        const Pos* pr = &pool[cur->prev];
        if ((uint16_t)(pr->ply + 1) != cur->ply) break;      // stale ring slot
        if (same_position(*pr, now) && ++matches >= 4) break;
        cur = pr;
    }
    return matches;
}

SZB_HD void set_repetition_flags(const Pos* pool, Pos& q) {
    int r = count_repetitions(pool, q);
    if (r >= 1) q.flags |= F_REP2;
    if (r >= 2) q.flags |= F_REP3;
    if (r >= 4) q.flags |= F_REP5;
}

// Board.has_insufficient_material(color)
SZB_HD bool insufficient_material(const Pos& p, bool white) {
    const uint64_t own = p.bb[white ? BB_WHITE : BB_BLACK], opp = p.bb[white ? BB_BLACK : BB_WHITE];
    if (own & (p.bb[BB_P] | p.bb[BB_R] | p.bb[BB_Q])) return false;
    if (own & p.bb[BB_N]) return popc(own) <= 2 && !(opp & ~p.bb[BB_K] & ~p.bb[BB_Q]);
    if (own & p.bb[BB_B]) {
        bool same = !(p.bb[BB_B] & DARK_SQUARES) || !(p.bb[BB_B] & ~DARK_SQUARES);
        return same && !p.bb[BB_P] && !p.bb[BB_N];
    }
    return true;
}

// Board.outcome(claim_draw=False); repetition flags must already be set.  `n_legal` from gen_legal.
SZB_HD uint8_t outcome_of(const Tables& T, const Pos& p, int n_legal) {
    if (n_legal == 0 && in_check(T, p)) return OUT_CHECKMATE;
    if (insufficient_material(p, true) && insufficient_material(p, false)) return OUT_INSUFFICIENT;
    if (n_legal == 0) return OUT_STALEMATE;
    if (p.halfmove >= 150) return OUT_75MOVES;
    if (p.flags & F_REP5) return OUT_FIVEFOLD;
    return OUT_NONE;
}

// has_kingside / has_queenside castling rights -> bits: 1 WK, 2 WQ, 4 BK, 8 BQ
SZB_HD int castling_flags(const Pos& p) {
    int out = 0;
    for (int c = 0; c < 2; c++) {
        const bool white = c == 0;
        const uint64_t k = p.bb[BB_K] & p.bb[white ? BB_WHITE : BB_BLACK] & (white ? RANK_1 : RANK_8);
        const uint8_t r = white ? p.rights_w : p.rights_b;
        if (!k || !r) continue;
        const int kf = lsb(k) & 7;
        if (r >> (kf + 1)) out |= white ? 1 : 4;
        if (r & ((1u << kf) - 1)) out |= white ? 2 : 8;
    }
    return out;
}

// ----------------------------------------------------------------------------------------------
// move <-> policy index (chess_tensor.py:221-410).  Index = plane*64 + row*8 + col in the mover's view.
// ----------------------------------------------------------------------------------------------
SZB_HD int view_square(int s, bool white) { return white ? (s ^ 56) : (s ^ 7); }   // row*8+col <-> square (involution)

SZB_HD int move_to_index(const Pos& p, uint16_t m) {
    const bool white = p.flags & F_WHITE;
    int from = mv_from(m), to = mv_to(m);
    const int promo = mv_promo(m);
    const uint64_t own = p.bb[white ? BB_WHITE : BB_BLACK];
    if (!(p.flags & F_960) && (p.bb[BB_K] & bit(from)) && (own & bit(to))) to = to < from ? from - 2 : from + 2;
    const int a = view_square(from, white), b = view_square(to, white);
    const int row = a >> 3, col = a & 7, dx = (b & 7) - col, dy = (b >> 3) - row;
    const int adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
    int plane;
    if (dx == 0 || dy == 0 || adx == ady) {
        if (promo >= PT_KNIGHT && promo <= PT_ROOK) plane = 64 + 3 * (promo - PT_KNIGHT) + (dx == 0 ? 0 : (dx > 0 ? 1 : 2));
        else {
            const int sx = (dx > 0) - (dx < 0), sy = (dy > 0) - (dy < 0);
            // (0,-1) N, (1,-1) NE, (1,0) E, (1,1) SE, (0,1) S, (-1,1) SW, (-1,0) W, (-1,-1) NW
            const int dir = sx == 0 ? (sy < 0 ? 0 : 4) : (sx > 0 ? 2 + sy : 6 - sy);
            plane = dir * 7 + (adx > ady ? adx : ady) - 1;
        }
    } else {
        // (1,-2) (2,-1) (2,1) (1,2) (-1,2) (-2,1) (-2,-1) (-1,-2)
        int k;
        if (dx > 0) k = dx == 1 ? (dy < 0 ? 0 : 3) : (dy < 0 ? 1 : 2);
        else k = dx == -1 ? (dy > 0 ? 4 : 7) : (dy > 0 ? 5 : 6);
        plane = 56 + k;
    }
    return plane * 64 + a;
}

// Inverse for the position the move is played from; returns MOVE_NONE when the index points off the board
// or at a square without an own piece.  Queen promotion is implied by a pawn reaching the last rank on a
// queen-move plane; castling is recognised per game mode and returned as king-takes-rook.
SZB_HD uint16_t index_to_move(const Pos& p, int index) {
    const bool white = p.flags & F_WHITE;
    const int plane = index >> 6, a = index & 63, row = a >> 3, col = a & 7;
    int trow, tcol, promo = 0;
    if (plane < 56) {
        const int dir = plane / 7, dist = plane % 7 + 1;
        const int sx = (dir >= 1 && dir <= 3) ? 1 : ((dir >= 5) ? -1 : 0);
        const int sy = (dir == 0 || dir == 1 || dir == 7) ? -1 : ((dir >= 3 && dir <= 5) ? 1 : 0);
        trow = row + sy * dist; tcol = col + sx * dist;
    } else if (plane < 64) {
        const int k = plane - 56;
        const int kx[8] = {1, 2, 2, 1, -1, -2, -2, -1}, ky[8] = {-2, -1, 1, 2, 2, 1, -1, -2};
        trow = row + ky[k]; tcol = col + kx[k];
    } else {
        const int k = plane - 64;
        trow = row - 1;
        tcol = col + (k % 3 == 1 ? 1 : (k % 3 == 2 ? -1 : 0));
        promo = PT_KNIGHT + k / 3;
    }
    if (trow < 0 || trow > 7 || tcol < 0 || tcol > 7) return MOVE_NONE;
    const int from = view_square(a, white);
    int to = view_square(trow * 8 + tcol, white);
    const uint64_t own = p.bb[white ? BB_WHITE : BB_BLACK];
    if (!(own & bit(from))) return MOVE_NONE;
    if ((p.bb[BB_P] & bit(from)) && plane < 56 && trow == 0) promo = PT_QUEEN;
    if (!(p.flags & F_960) && (p.bb[BB_K] & bit(from)) && (from & 7) == 4 && (from >> 3) == (to >> 3) &&
        (to - from == 2 || to - from == -2))
        to = (from & 56) + (to > from ? 7 : 0);
    return mk_move(from, to, promo);
}

// ----------------------------------------------------------------------------------------------
// input planes, bit-packed: out[k] holds plane k with bit (row*8+col) in the mover's view
// (chess_tensor.py:38-63,123-142; layout: SURVEY.md Appendix B).
// ----------------------------------------------------------------------------------------------
SZB_HD uint64_t view_bb(uint64_t x, bool white) { return white ? bswap64(x) : bswap64(brev64(x)); }

SZB_HD void pack_planes(const Pos* pool, const Pos& now, uint64_t* out) {
    const bool white = now.flags & F_WHITE;
    const int own = white ? BB_WHITE : BB_BLACK, opp = white ? BB_BLACK : BB_WHITE;
    const Pos* cur = &now;
    for (int t = 0; t < 8; t++) {
        uint64_t* o = out + 14 * t;
        if (cur) {
#pragma unroll
            for (int k = 0; k < 6; k++) {
                o[k] = view_bb(cur->bb[BB_P + k] & cur->bb[own], white);
                o[6 + k] = view_bb(cur->bb[BB_P + k] & cur->bb[opp], white);
            }
            o[12] = (cur->flags & F_REP2) ? ~0ull : 0ull;
            o[13] = (cur->flags & F_REP3) ? ~0ull : 0ull;
            const Pos* pr = (cur->prev == NO_PREV) ? nullptr : &pool[cur->prev];
            if (pr && (uint16_t)(pr->ply + 1) != cur->ply) pr = nullptr;
            cur = pr;
        } else {
#pragma unroll
            for (int k = 0; k < 14; k++) o[k] = 0;
        }
    }
    const int cf = now.ply ? castling_flags(now) : 15;       // hard-coded ones before the first move (chess_tensor.py:82)
    const int ownc = white ? (cf & 3) : (cf >> 2), oppc = white ? (cf >> 2) : (cf & 3);
    out[112] = white ? ~0ull : 0ull;
    out[113] = now.ply ? ~0ull : 0ull;
    out[114] = (ownc & 1) ? ~0ull : 0ull;
    out[115] = (ownc & 2) ? ~0ull : 0ull;
    out[116] = (oppc & 1) ? ~0ull : 0ull;
    out[117] = (oppc & 2) ? ~0ull : 0ull;
    out[118] = (now.ply && now.halfmove) ? ~0ull : 0ull;
}

// ----------------------------------------------------------------------------------------------
// position setup (host side helpers, also used by the host harness)
// ----------------------------------------------------------------------------------------------
// python-chess clean_castling_rights for the first position of a game
SZB_HD void clean_rights(Pos& p) {
    for (int c = 0; c < 2; c++) {
        const bool white = c == 0;
        const int br = white ? 0 : 56;
        const uint64_t own = p.bb[white ? BB_WHITE : BB_BLACK];
        uint8_t r = white ? p.rights_w : p.rights_b;
        r &= (uint8_t)((p.bb[BB_R] & own) >> br);
        const uint64_t k = p.bb[BB_K] & own & (0xFFull << br);
        if (!k) r = 0;
        else {
            const int kf = lsb(k) & 7;
            if (!(p.flags & F_960)) {
                r &= 0x81;
                if (kf != 4) r = 0;
            } else {
                uint8_t a = r ? (uint8_t)(r & (~r + 1)) : 0;          // lowest file
                uint8_t h = 0;
                for (int f = 7; f >= 0 && !h; f--) if (r & (1u << f)) h = (uint8_t)(1u << f);
                r = 0;
                if (a && a < (1u << kf)) r |= a;
                if (h && h > (1u << kf)) r |= h;
            }
        }
        if (white) p.rights_w = r; else p.rights_b = r;
    }
}

SZB_HD void finish_setup(const Tables& T, Pos& p) {
    p.prev = NO_PREV;
    p.flags &= (F_WHITE | F_960);
    clean_rights(p);
    if (has_legal_ep(T, p)) p.flags |= F_EPLEGAL;
    p.key = position_key(p);
    p.outcome = OUT_NONE;
    p.n_legal = 0;
    p.pad[0] = p.pad[1] = p.pad[2] = 0;
}

// Scharnagl numbering -> back rank piece types on files a..h
inline void chess960_backrank(int id, int* row8) {
    static const int KNT[10][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {1, 2}, {1, 3}, {1, 4}, {2, 3}, {2, 4}, {3, 4}};
    for (int i = 0; i < 8; i++) row8[i] = 0;
    int n = id;
    row8[(n % 4) * 2 + 1] = PT_BISHOP; n /= 4;
    row8[(n % 4) * 2] = PT_BISHOP; n /= 4;
    int q = n % 6; n /= 6;
    for (int i = 0, k = 0; i < 8; i++) if (!row8[i]) { if (k == q) { row8[i] = PT_QUEEN; break; } k++; }
    for (int i = 0, k = 0; i < 8; i++) if (!row8[i]) { if (k == KNT[n][0] || k == KNT[n][1]) row8[i] = PT_KNIGHT; k++; }
    for (int i = 0, k = 0; i < 8; i++) if (!row8[i]) { row8[i] = k == 1 ? PT_KING : PT_ROOK; k++; }
}

// id < 0: the vanilla start position (chess960 off); else Board.from_chess960_pos(id)
inline void start_position(const Tables& T, int id, Pos& p) {
    int row[8];
    chess960_backrank(id < 0 ? 518 : id, row);
    for (int i = 0; i < 8; i++) p.bb[i] = 0;
    p.rights_w = p.rights_b = 0;
    for (int f = 0; f < 8; f++) {
        p.bb[BB_P + row[f] - 1] |= bit(f) | bit(56 + f);
        if (row[f] == PT_ROOK) { p.rights_w |= (uint8_t)(1u << f); p.rights_b |= (uint8_t)(1u << f); }
    }
    p.bb[BB_P] = 0x00FF00000000FF00ull;
    p.bb[BB_WHITE] = 0x000000000000FFFFull;
    p.bb[BB_BLACK] = 0xFFFF000000000000ull;
    p.flags = (uint8_t)(F_WHITE | (id < 0 ? 0 : F_960));
    p.ep = -1; p.halfmove = 0; p.ply = 0;
    finish_setup(T, p);
}

}  // namespace szb
