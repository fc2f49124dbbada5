// Host build of the DEVICE chess code (sigma-zero_b200/csrc/chess.cuh) -- test infrastructure only.
// Lets `-m "not gpu"` tests run the exact functions the CUDA kernels call against the oracle on a
// machine without a GPU.  Never loaded by the product.
#include <vector>
#include <algorithm>
#include <cstring>
#include "../../sigma-zero_b200/csrc/chess.cuh"

using namespace szb;

struct Game {
    std::vector<Pos> pool;
};
static Tables g_tables;
static bool g_init = false;
static const Tables& tables() {
    if (!g_init) { build_tables(g_tables); g_init = true; }
    return g_tables;
}

static void analyse(Game* g) {
    Pos& p = g->pool.back();
    uint16_t mv[MAX_MOVES];
    int n = gen_legal(tables(), p, mv);
    p.n_legal = (uint8_t)n;
    p.outcome = outcome_of(tables(), p, n);
}

extern "C" {

void* hh_new(int start_id) {
    Game* g = new Game();
    Pos p;
    std::memset(&p, 0, sizeof(p));
    start_position(tables(), start_id, p);
    g->pool.push_back(p);
    analyse(g);
    return g;
}

void* hh_set(const uint64_t* bb12, int turn, int rights_w, int rights_b, int ep, int halfmove, int ply, int chess960) {
    Game* g = new Game();
    Pos p;
    std::memset(&p, 0, sizeof(p));
    for (int t = 0; t < 6; t++) {
        p.bb[BB_P + t] = bb12[t] | bb12[6 + t];
        p.bb[BB_WHITE] |= bb12[t];
        p.bb[BB_BLACK] |= bb12[6 + t];
    }
    p.flags = (uint8_t)((turn ? F_WHITE : 0) | (chess960 ? F_960 : 0));
    p.rights_w = (uint8_t)rights_w; p.rights_b = (uint8_t)rights_b;
    p.ep = (int8_t)ep; p.halfmove = (uint8_t)std::min(halfmove, 250); p.ply = (uint16_t)ply;
    finish_setup(tables(), p);
    g->pool.push_back(p);
    analyse(g);
    return g;
}

void hh_free(void* h) { delete (Game*)h; }

int hh_legal(void* h, uint16_t* idx_out) {
    Game* g = (Game*)h;
    const Pos& p = g->pool.back();
    uint16_t mv[MAX_MOVES];
    int n = gen_legal(tables(), p, mv);
    for (int i = 0; i < n; i++) idx_out[i] = (uint16_t)move_to_index(p, mv[i]);
    std::sort(idx_out, idx_out + n);
    return n;
}

// round trip check: every legal move must survive move -> index -> move
int hh_codec_roundtrip(void* h) {
    Game* g = (Game*)h;
    const Pos& p = g->pool.back();
    uint16_t mv[MAX_MOVES];
    int n = gen_legal(tables(), p, mv), bad = 0;
    for (int i = 0; i < n; i++) bad += index_to_move(p, move_to_index(p, mv[i])) != mv[i];
    return bad;
}

int hh_push_index(void* h, int idx) {
    Game* g = (Game*)h;
    const Pos p = g->pool.back();
    uint16_t m = index_to_move(p, idx);
    if (m == MOVE_NONE) return -1;
    uint16_t mv[MAX_MOVES];
    int n = gen_legal(tables(), p, mv);
    if (std::find(mv, mv + n, m) == mv + n) return -2;
    Pos q;
    make_move(tables(), p, m, q);
    q.prev = (uint32_t)(g->pool.size() - 1);
    set_repetition_flags(g->pool.data(), q);
    g->pool.push_back(q);
    analyse(g);
    return 0;
}

void hh_planes(void* h, uint64_t* out119) {
    Game* g = (Game*)h;
    pack_planes(g->pool.data(), g->pool.back(), out119);
}

int hh_outcome(void* h) { return ((Game*)h)->pool.back().outcome; }
int hh_flags(void* h) { return ((Game*)h)->pool.back().flags; }
int hh_turn(void* h) { return (((Game*)h)->pool.back().flags & F_WHITE) ? 1 : 0; }

static uint64_t perft_rec(const Pos& p, int depth) {
    uint16_t mv[MAX_MOVES];
    int n = gen_legal(tables(), p, mv);
    if (depth <= 1) return (uint64_t)n;
    uint64_t t = 0;
    for (int i = 0; i < n; i++) {
        Pos q;
        make_move(tables(), p, mv[i], q);
        t += perft_rec(q, depth - 1);
    }
    return t;
}

uint64_t hh_perft(void* h, int depth) {
    if (depth <= 0) return 1;
    return perft_rec(((Game*)h)->pool.back(), depth);
}

int hh_sizeof_pos() { return (int)sizeof(Pos); }
}

// ---- tree.cuh arithmetic (host build of the device functions) -----------------------------------
#include "../../sigma-zero_b200/csrc/tree.cuh"
extern "C" {
float hh_puct(int n, double w, float prior, int n_parent, float c) { return puct_score(n, w, prior, sqrt_parent(n_parent), c); }
float hh_cascade_sum(const float* x) {
    float part[32];
    for (int t = 0; t < 32; t++) part[t] = cascade_lane([&](int e) { return x[e]; }, t);
    return cascade_combine([&](int t) { return part[t]; });
}
// masked sum through the sparse walk (x must already be +0 outside the mask for the dense reference to agree)
float hh_cascade_sum_sparse(const float* x, const uint64_t* mask73) {
    float part[32];
    for (int t = 0; t < 32; t++) part[t] = cascade_lane_sparse(mask73, [&](int e) { return x[e]; }, t);
    return cascade_combine([&](int t) { return part[t]; });
}
float hh_noisy_prior(float p) { return noisy_prior(p); }
void hh_hash_eval(const uint64_t* words119, float* policy4672, float* value) {
    uint64_t h = he_fold(words119);
    for (int i = 0; i < N_ACTIONS; i++) policy4672[i] = he_policy(h, i);
    *value = he_value(h);
}
}
