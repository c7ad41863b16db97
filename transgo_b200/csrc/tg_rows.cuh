// tg_rows.cuh -- row-per-lane Go rules: one board ROW per lane, several games per warp.
//
// Second formulation of the rules engine (GoEnv/cpp_src/board.cc), used by the playout kernel.  tg_board.cuh keeps a
// whole board in every lane (3 words at 9x9, 7 at 19x19), so each whole-board operation is issued once per lane and the
// T lanes of a tile repeat it; that is what bounds k_playout<N,32> (ALU pipe 80 % busy, ncu profiles/r01b).  Here lane l
// of a warp holds ONE row of one game, games are laid side by side in the warp with an always-zero separator lane
// between them (9x9: lanes 1-9, 11-19, 21-29 = three games; 19x19: lanes 1-19 = one game), and
//     E/W neighbours  = a shift inside the lane,
//     N/S neighbours  = one SHFL from lane-1 / lane+1 (the separators supply the off-board zeros),
// so a whole-board operation on three games costs ONE warp instruction.  Per-group liberty counts (what Block::liberties
// holds, board.h:22) are never formed; the two places the rules need them are restated as whole-board floods:
//   * captures (board.cc:606-630): after the stone is placed, enemy stones NOT connected to a liberty are dead;
//   * suicide / capture test of a point with no empty neighbour (board.cc:90-158): a neighbour group matters only
//     through "has a liberty other than this point", i.e. connected to another liberty.
// Everything is a function of the position only, so results are bit-identical to the bitboard engine and to go_env.so
// (tests/test_env_gpu.py replays every ply on the oracle).
#pragma once
#include "tg_board.cuh"

namespace tg {

constexpr unsigned kFull = 0xffffffffu;

#ifdef TG_ROWS_STATS
__device__ unsigned long long g_rows_stats[8];
#define TG_STAT(i, n) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_rows_stats[i], (unsigned long long)(n)); } while (0)
#else
#define TG_STAT(i, n) do { } while (0)
#endif

template <int N>
struct RowGeo {
    static constexpr int LPG = N + 1;                 // lanes per game: N rows + one separator
    static constexpr int GPW = 31 / LPG;              // games per warp (lane 0 is the first separator)
    static constexpr uint32_t ROW = (1u << N) - 1u;
    static constexpr uint32_t EDGE = (1u << (N - 1)) | 1u;
};

// value of the lane above / below (row y-1 / y+1); separators and lanes 0 / 31 hold zeros
__device__ __forceinline__ uint32_t row_up(uint32_t x) { return __shfl_up_sync(kFull, x, 1); }
__device__ __forceinline__ uint32_t row_dn(uint32_t x) { return __shfl_down_sync(kFull, x, 1); }

// 4-neighbourhood of a set; m = this lane's row mask (ROW on row lanes of live games, 0 elsewhere)
__device__ __forceinline__ uint32_t rnbrs(uint32_t x, uint32_t m) {
    return ((x << 1) | (x >> 1) | row_up(x) | row_dn(x)) & m;
}

// One game per (warp, slot).  Row bits live in b / w; the scalar fields are replicated in the game's lanes.
template <int N>
struct RowPos {
    uint32_t b, w;
    int last1, last2, ko_loc, ko_age, step, player, ko_color, terminated;
};

template <int N>
struct RowLane {
    using R = RowGeo<N>;
    int lane, slot, r, lead;      // lead = lane holding row 0 of this game
    bool rowlane;                 // holds a row of an existing game
    uint32_t gm;                  // ballot mask of this game's row lanes (0 on non-row lanes)
    long long g;                  // game index
    __device__ __forceinline__ void init(long long warp, long long G) {
        lane = (int)(threadIdx.x & 31);
        const int l1 = lane - 1;
        slot = l1 < 0 ? R::GPW : l1 / R::LPG;
        r = l1 < 0 ? N : l1 - slot * R::LPG;
        g = warp * R::GPW + slot;
        rowlane = slot < R::GPW && r < N && g < G;
        lead = 1 + (slot < R::GPW ? slot : 0) * R::LPG;
        gm = rowlane ? (R::ROW << lead) : 0u;
    }
};

template <int N>
__device__ __forceinline__ uint32_t row_of(const BB<N>& a, int r) {
    using G = Geo<N>;
    typename G::word_t v = 0;
    const int wi = r / G::RPW;
    TG_FOR_W v = (i == wi) ? a.w[i] : v;
    return (uint32_t)(v >> ((r - wi * G::RPW) * G::STRIDE)) & RowGeo<N>::ROW;
}

template <int N>
__device__ __forceinline__ void rows_load(RowPos<N>& p, const RowLane<N>& L, const uint4* __restrict__ states, long long G) {
    p.b = p.w = 0;
    p.last1 = p.last2 = p.ko_loc = kInvalid; p.ko_age = 0; p.step = 1; p.player = kBlack; p.ko_color = 0; p.terminated = 1;
    if (L.rowlane) {
        Pos<N> s; pos_load<N>(s, states, L.g, G, 1);
        p.b = row_of<N>(s.stones[0], L.r); p.w = row_of<N>(s.stones[1], L.r);
        p.last1 = s.last1; p.last2 = s.last2; p.ko_loc = s.ko_loc; p.ko_age = s.ko_age; p.step = s.step;
        p.player = s.player; p.ko_color = s.ko_color; p.terminated = s.terminated;
    }
}

// gather the rows of each game into its lead lane and store the 3-word / 7-word state (once per kernel)
template <int N>
__device__ __forceinline__ void rows_store(const RowPos<N>& p, const RowLane<N>& L, uint4* __restrict__ states, long long G) {
    using GG = Geo<N>; using word_t = typename GG::word_t;
    Pos<N> s;
    s.stones[0] = bb_zero<N>(); s.stones[1] = bb_zero<N>();
#pragma unroll
    for (int y = 0; y < N; ++y) {
        const uint32_t rb = __shfl_sync(kFull, p.b, (L.lead + y) & 31), rw = __shfl_sync(kFull, p.w, (L.lead + y) & 31);
        const int wi = y / GG::RPW, sh = (y % GG::RPW) * GG::STRIDE;
        s.stones[0].w[wi] |= word_t(rb) << sh; s.stones[1].w[wi] |= word_t(rw) << sh;
    }
    if (L.rowlane && L.r == 0) {
        s.last1 = p.last1; s.last2 = p.last2; s.ko_loc = p.ko_loc; s.ko_age = p.ko_age; s.step = p.step;
        s.player = p.player; s.ko_color = p.ko_color; s.terminated = p.terminated;
        pos_store<N>(s, states, L.g, G, 1);
    }
}

// true eyes of `own` against `opp` (board.cc:665-714), row form.  m = row mask of this lane.
template <int N>
__device__ __forceinline__ uint32_t rtrue_eyes(uint32_t own, uint32_t opp, uint32_t m, int r) {
    using R = RowGeo<N>;
    const uint32_t empty = ~(own | opp) & m;
    const uint32_t eye = empty & ~rnbrs(~own & m, m);               // every on-board neighbour is own
    const uint32_t on = row_up(opp), os = row_dn(opp);              // opponent stones one row up / down
    const uint32_t d1 = (on << 1) & m, d2 = on >> 1, d3 = (os << 1) & m, d4 = os >> 1;
    const uint32_t anyd = d1 | d2 | d3 | d4;
    const uint32_t twod = (d1 & d2) | ((d1 | d2) & d3) | ((d1 | d2 | d3) & d4);
    const uint32_t edge = (r == 0 || r == N - 1) ? R::ROW : R::EDGE;
    return eye & ~((edge & anyd) | (~edge & twod));
}

// complete E/W fill of the seeds g (a subset of mk) along the runs of mk: the carry of mk + g ripples up each run, the
// same on the bit-reversed word ripples down (rmk = __brev(mk)).
__device__ __forceinline__ uint32_t row_hfill(uint32_t g, uint32_t mk, uint32_t rmk) {
    const uint32_t t = mk + g;
    const uint32_t t2 = rmk + __brev(g);
    return (((t ^ mk) & mk) | g) | __brev((t2 ^ rmk) & rmk);
}

// Same-colour connectivity of a position, prepared once per ply: flood steps move one row up and down and then fill
// each row completely, black groups through black and white through white.
struct RowLinks {
    uint32_t a, b, ra, rb, su, sd;       // colours, their bit reversals, "the stone above / below has my colour"
    __device__ __forceinline__ void init(uint32_t black, uint32_t white) {
        a = black; b = white; ra = __brev(a); rb = __brev(b);
        su = (a & row_up(a)) | (b & row_up(b));
        sd = (a & row_dn(a)) | (b & row_dn(b));
    }
    __device__ __forceinline__ uint32_t step_all(uint32_t g) const {
        const uint32_t n0 = g | (row_up(g) & su) | (row_dn(g) & sd);
        return row_hfill(n0 & a, a, ra) | row_hfill(n0 & b, b, rb);
    }
};

// Legal board points (row bits) of the player to move: board.cc:437-489 for every point at once.
// Warp-collective: every lane of the warp must call it.  m = ROW on row lanes of games that are still running, else 0.
template <int N>
__device__ __forceinline__ uint32_t rlegal(const RowPos<N>& p, const RowLane<N>& L, uint32_t m) {
    const uint32_t b = p.b & m, w = p.w & m;
    const uint32_t st = b | w;
    const uint32_t e = ~st & m;
    const uint32_t ne = rnbrs(e, m);
    uint32_t lg = e & ne;                                             // an empty neighbour: never suicide
    const uint32_t C = e & ~ne;                                       // enclosed points
    if (__any_sync(kFull, C != 0)) {
        TG_STAT(1, 1);
        const bool me_black = p.player == kBlack;
        const uint32_t own = me_black ? b : w, opp = me_black ? w : b;
        RowLinks lk; lk.init(b, w);
        // T: stones of groups that own an OPEN liberty (an empty point with an empty neighbour).  Such a group has a
        // liberty other than any enclosed point: own -> the point is safe to play, enemy -> it cannot be captured there.
        uint32_t T = rnbrs(lg, m) & st;
        while (true) {
            TG_STAT(2, 1);
            const uint32_t n = lk.step_all(T);
            const bool ch = n != T;
            T = n;
            if (!__any_sync(kFull, ch)) break;
        }
        const uint32_t K = st & ~T;                                   // closed groups: every liberty is an enclosed point
        uint32_t okC = C & rnbrs(own & T, m);
        uint32_t U = C & ~okC & rnbrs(K, m);                          // undecided: next to a closed group
        if (__any_sync(kFull, U != 0)) {
            RowLinks lkK; lkK.init(b & K, w & K);                      // closed groups are whole groups: links inside K
            do {
                // every game takes its lowest undecided point u and asks which closed groups keep a liberty other than u
                const uint32_t ub = __ballot_sync(kFull, U != 0) & L.gm;
                const bool mine = ub != 0 && L.lane == (__ffs((int)ub) - 1);
                const uint32_t u = mine ? (U & (0u - U)) : 0u;
                TG_STAT(3, 1);
                const uint32_t nuK = rnbrs(u, m) & K;                  // the closed neighbour stones of u
                uint32_t S = rnbrs(C & ~u, m) & K;
                while (true) {
                    if (!__any_sync(kFull, (nuK & ~S) != 0)) break;    // every one of them reached: all have another liberty
                    TG_STAT(4, 1);
                    const uint32_t n = lkK.step_all(S);
                    const bool ch = n != S;
                    S = n;
                    if (!__any_sync(kFull, ch)) break;
                }
                // legal: an own neighbour group with another liberty, or an enemy neighbour group with none (it is captured).
                // (after an early exit every game's nuK lies in S; after a fixed point S is exact)
                const uint32_t okb = __ballot_sync(kFull, ((nuK & own & S) | (nuK & opp & ~S)) != 0) & L.gm;
                okC |= okb ? u : 0u;
                U &= ~u;
            } while (__any_sync(kFull, U != 0));
        }
        lg |= okC;
    }
    if (p.ko_age == 0 && p.ko_color == p.player && p.ko_loc >= 0 && p.ko_loc < N * N) {       // board.cc:198-200
        const int ky = p.ko_loc / N;
        if (ky == L.r) lg &= ~(1u << (p.ko_loc - ky * N));
    }
    return lg;
}

// Place the stone `pb` (one bit in one lane of the game, zero elsewhere / for games that do not place a stone) for the
// player to move: captures + simple-ko bookkeeping (board.cc:546-653, 163-194).  Warp-collective.  `placing` says
// whether this lane's game places a stone; action = its point index.
template <int N>
__device__ __forceinline__ void rplace(RowPos<N>& p, const RowLane<N>& L, uint32_t m, uint32_t pb, bool placing, int action) {
    const bool me_black = p.player == kBlack;
    const uint32_t b = p.b & m, w = p.w & m;
    const uint32_t own = me_black ? b : w, opp = me_black ? w : b;
    const uint32_t e = ~(b | w) & m;
    const uint32_t around = rnbrs(pb, m);
    const uint32_t touch = around & opp;
    uint32_t dead = 0;
    if (__any_sync(kFull, touch != 0)) {
        TG_STAT(5, 1);
        // enemy stones connected to a liberty of the position with the new stone on it; stop as soon as every enemy
        // neighbour of every new stone is known to be alive (the common case), else run to the fixed point
        const uint32_t ropp = __brev(opp);
        uint32_t gl = rnbrs(e & ~pb, m) & opp;
        bool settled = false;
        while (true) {
            if (!__any_sync(kFull, (touch & ~gl) != 0)) { settled = true; break; }
            TG_STAT(6, 1);
            const uint32_t n1 = row_hfill((gl | row_up(gl) | row_dn(gl)) & opp, opp, ropp);
            const bool ch = n1 != gl;
            gl = n1;
            if (!__any_sync(kFull, ch)) break;
        }
        if (!settled) {
            const uint32_t hb = __ballot_sync(kFull, (touch & ~gl) != 0) & L.gm;
            dead = hb ? (opp & ~gl) : 0u;                              // the floods are complete: exact
        }
    }
    bool ko = false;
    if (__any_sync(kFull, dead != 0)) {
        TG_STAT(7, 1);
        // simple ko (board.cc:163-194, 561-570): the new stone has no empty and no own neighbour, and exactly one of the
        // captured neighbours is a single stone
        const uint32_t hit = around & dead;
        const uint32_t blocked = __ballot_sync(kFull, (around & (e | own)) != 0) & L.gm;
        const uint32_t cand = hit & ~rnbrs(opp, m);
        const int pc = __popc(cand);
        const uint32_t b1 = __ballot_sync(kFull, pc > 0) & L.gm, b2 = __ballot_sync(kFull, pc > 1) & L.gm;
        const int src = b1 ? (__ffs((int)b1) - 1) : L.lane;
        const int kloc = __shfl_sync(kFull, L.r * N + (__ffs((int)cand) - 1), src);
        ko = placing && blocked == 0 && b2 == 0 && __popc(b1) == 1;
        if (ko) { p.ko_loc = kloc; p.ko_color = 3 - p.player; p.ko_age = 0; }
    }
    if (placing) {
        if (!ko) p.ko_age += 1;
        const uint32_t own2 = own | pb, opp2 = opp & ~dead;
        p.b = me_black ? own2 : opp2; p.w = me_black ? opp2 : own2;
        p.player = 3 - p.player; p.last2 = p.last1; p.last1 = action; p.step += 1;
    }
}

}  // namespace tg
