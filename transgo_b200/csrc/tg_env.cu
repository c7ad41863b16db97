// tg_env.cu -- batched board kernels + the C ABI of include/transgo_b200.h (environment half).
// Replaces GoEnv/cpp_src/go_env.cc (reference C ABI) and the loops of board.cc / board_feature.cc behind it.
#include <cstdio>
#include <cstring>
#include <mutex>
#include "../../include/transgo_b200.h"
#include "tg_board.cuh"
#include "tg_common.cuh"
#include "tg_rows.cuh"
#include "tg_inc.cuh"

namespace tg {

// ---------------------------------------------------------------------------------------------- kernels
template <int N>
__global__ void k_reset(uint4* __restrict__ states, long long G) {
    long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= G) return;
    Pos<N> s; pos_reset(s);
    pos_store<N>(s, states, g, G, 1);
}

template <int N, int T>
__global__ void __launch_bounds__(kThreads)
k_step(uint4* __restrict__ states, long long G, const int16_t* __restrict__ actions, uint32_t* __restrict__ legal_out,
       uint8_t* __restrict__ done_out, uint8_t* __restrict__ acc_out, int max_step) {
    using GG = Geo<N>;
    long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / T;
    if (g >= G) return;                       // whole tiles leave together
    Pos<N> s; pos_load<N>(s, states, g, G, 1);
    // go_env.cc:51-80 with the lazy group analysis of the playout kernel (every branch is tile-uniform)
    int action = (int)actions[g];
    if (action == GG::P) action = kPass;                                  // go_env.cc:56-57
    bool acc = false, done = false;
    BB<N> l1b, l1w, known;
    if (s.terminated) done = true;                                        // go_env.cc:52-55
    else if (action == kPass || action == kResign) { done = apply_pass<N>(s, action, max_step); acc = true; }
    else if (action >= 0 && action < GG::P) {
        const BB<N> pb = bb_point<N>(action);
        const BB<N> lg0 = legal_points_lazy<N, T>(s, l1b, l1w, known);
        if (any(lg0 & pb)) { done = apply_stone_lazy<N, T>(s, pb, action, l1b, l1w, known, max_step); acc = true; }   // else board.cc:448-462
    }
    BB<N> lg = bb_zero<N>();
    if (legal_out) lg = legal_points_lazy<N, T>(s, l1b, l1w, known);     // uniform
    if (Tile<T>::lane() == 0) {
        if (acc) pos_store<N>(s, states, g, G, 1);
        if (done_out) done_out[g] = done ? 1 : 0;
        if (acc_out) acc_out[g] = acc ? 1 : 0;
        if (legal_out) {
            uint32_t d[GG::DQ * 4];
#pragma unroll
            for (int i = 0; i < GG::DQ * 4; ++i) d[i] = 0;
            bb_to_dense<N>(lg, d);
            if (!any(lg)) d[GG::P >> 5] |= 1u << (GG::P & 31);
            uint4* o = reinterpret_cast<uint4*>(legal_out) + g * GG::DQ;
#pragma unroll
            for (int k = 0; k < GG::DQ; ++k) o[k] = make_uint4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
        }
    }
}

template <int N, int T>
__global__ void __launch_bounds__(kThreads)
k_legal(const uint4* __restrict__ states, long long G, uint32_t* __restrict__ legal_out, int no_eye) {
    using GG = Geo<N>;
    long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / T;
    if (g >= G) return;
    Pos<N> s; pos_load<N>(s, states, g, G, 1);
    BB<N> l1b, l1w, known;
    BB<N> lg = legal_points_lazy<N, T>(s, l1b, l1w, known);
    if (no_eye) lg = andn(lg, true_eyes(stones_of(s, s.player), stones_of(s, 3 - s.player)));    // board.cc:510-512
    if (Tile<T>::lane() == 0) {
        uint32_t d[GG::DQ * 4];
#pragma unroll
        for (int i = 0; i < GG::DQ * 4; ++i) d[i] = 0;
        bb_to_dense<N>(lg, d);
        if (!any(lg)) d[GG::P >> 5] |= 1u << (GG::P & 31);
        uint4* o = reinterpret_cast<uint4*>(legal_out) + g * GG::DQ;
#pragma unroll
        for (int k = 0; k < GG::DQ; ++k) o[k] = make_uint4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
    }
}

template <int N, int T>
__global__ void __launch_bounds__(kThreads)
k_encode(const uint4* __restrict__ states, long long G, float* __restrict__ planes) {
    constexpr int GPC = kThreads / T;          // games per CTA (even -> the CTA's float range is 16 B aligned)
    extern __shared__ uint32_t sm_bits[];
    const long long g0 = blockIdx.x * (long long)GPC;
    const int local = threadIdx.x / T;
    const long long g = g0 + local;
    planes_smem_clear<N>(sm_bits, GPC);
    __syncthreads();
    if (g < G) {
        Pos<N> s; pos_load<N>(s, states, g, G, 1);
        BB<N> pl[10];
        feature_planes<N, T>(s, pl);
        if constexpr (T >= 8) planes_to_smem_staged<N, T>(pl, sm_bits, local, sm_bits + (planes_smem_bytes<N>(GPC) >> 2) + local * planes_stage_words<N>());
        else planes_to_smem<N, T>(pl, sm_bits, local);
    }
    __syncthreads();
    const int ng = (int)((G - g0) < GPC ? (G - g0) : GPC);
    planes_smem_to_global<N>(sm_bits, planes + g0 * (long long)Geo<N>::PLANE_BITS, ng);
}

template <int N, int T>
__global__ void __launch_bounds__(kThreads)
k_score(const uint4* __restrict__ states, long long G, float komi, float* __restrict__ score, float* __restrict__ territory) {
    using GG = Geo<N>;
    long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / T;
    if (g >= G) return;
    Pos<N> s; pos_load<N>(s, states, g, G, 1);
    BB<N> bt, wt;
    int raw = area_score<N, T>(s.stones[0], s.stones[1], bt, wt);
    if (Tile<T>::lane() == 0 && score) score[g] = (float)raw - komi;              // go_env.cc:129
    if (territory) {                                                             // go_env.cc:136-149
        float* t = territory + g * (long long)GG::P;
        for (int p = Tile<T>::lane(); p < GG::P; p += T)
            t[p] = bb_test(bt, p) ? 1.0f : (bb_test(wt, p) ? -1.0f : 0.0f);
    }
}

template <int N, int T>
__global__ void __launch_bounds__(kThreads)
k_playout(uint4* __restrict__ states, long long G, int plies, int max_step, unsigned long long seed, long long game_id0,
          int no_eye, int16_t* __restrict__ trace, unsigned long long* __restrict__ plies_done) {
    using GG = Geo<N>;
    long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / T;
    if (g >= G) return;
    Pos<N> s; pos_load<N>(s, states, g, G, 1);
    int played = 0;
    for (int t = 0; t < plies; ++t) {
        int action = kInvalid;
        if (!s.terminated) {                  // tile-uniform
            BB<N> l1b, l1w, known;
            BB<N> lg = legal_points_lazy<N, T>(s, l1b, l1w, known);
            if (no_eye) lg = andn(lg, true_eyes(stones_of(s, s.player), stones_of(s, 3 - s.player)));
            const int nl = count(lg);
            if (nl == 0) {
                action = GG::P;
                apply_pass<N>(s, kPass, max_step);
            } else {
                const uint32_t r = mix32(seed, (unsigned long long)(game_id0 + g), (unsigned long long)s.step);
                int wi, bit;
                bb_locate<N>(lg, (int)(((unsigned long long)r * (unsigned)nl) >> 32), wi, bit);
                BB<N> pb; TG_FOR_W pb.w[i] = (i == wi) ? (typename GG::word_t(1) << bit) : typename GG::word_t(0);
                action = wb_to_pt<N>(wi, bit);
                apply_stone_lazy<N, T>(s, pb, action, l1b, l1w, known, max_step);
            }
            played += 1;
        }
        if (trace && Tile<T>::lane() == 0) trace[(long long)t * G + g] = (int16_t)action;
    }
    if (Tile<T>::lane() == 0) {
        pos_store<N>(s, states, g, G, 1);
        if (plies_done && played) atomicAdd(plies_done, (unsigned long long)played);
    }
}

// Warp-uniform variant of k_playout for T < 32: the tiles that share a warp vote together (vote_any<T, true>), so the warp
// never executes two games' control paths one after the other.
template <int N, int T>
__global__ void __launch_bounds__(kThreads)
k_playout_wu(uint4* __restrict__ states, long long G, int plies, int max_step, unsigned long long seed, long long game_id0,
             int no_eye, int16_t* __restrict__ trace, unsigned long long* __restrict__ plies_done) {
    using GG = Geo<N>;
    long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / T;
    const bool exists = g < G;
    if (!__any_sync(0xffffffffu, exists)) return;
    Pos<N> s; pos_reset(s); s.terminated = 1;
    if (exists) pos_load<N>(s, states, g, G, 1);
    int played = 0;
    for (int t = 0; t < plies; ++t) {
        int action = kInvalid;
        const bool active = !s.terminated;
        if (__any_sync(0xffffffffu, active)) {
            BB<N> l1b, l1w, known;
            BB<N> lg = legal_points_lazy<N, T, true>(s, l1b, l1w, known);
            if (no_eye) lg = andn(lg, true_eyes(stones_of(s, s.player), stones_of(s, 3 - s.player)));
            const int nl = count(lg);
            const bool placing = active && nl > 0;
            const uint32_t r = mix32(seed, (unsigned long long)(game_id0 + g), (unsigned long long)s.step);
            int wi, bit;
            bb_locate<N>(lg, placing ? (int)(((unsigned long long)r * (unsigned)nl) >> 32) : -1, wi, bit);
            BB<N> pb; TG_FOR_W pb.w[i] = (placing && i == wi) ? (typename GG::word_t(1) << bit) : typename GG::word_t(0);
            const int a = wb_to_pt<N>(wi < 0 ? 0 : wi, bit);
            apply_stone_lazy<N, T, true>(s, pb, a, l1b, l1w, known, max_step, placing);
            if (active) {
                action = placing ? a : GG::P;
                if (!placing) apply_pass<N>(s, kPass, max_step);
                played += 1;
            }
        }
        if (trace && exists && Tile<T>::lane() == 0) trace[(long long)t * G + g] = (int16_t)action;
    }
    if (exists && Tile<T>::lane() == 0) {
        pos_store<N>(s, states, g, G, 1);
        if (plies_done && played) atomicAdd(plies_done, (unsigned long long)played);
    }
}

// Incremental engine (tg_inc.cuh): dense bit-string boards, the atari set kept in registers across plies.  T lanes per game
// (T >= 8: 32/T games per warp in lock step, lanes = the <= 5 groups around the new stone; T == 1: one thread per game).
template <int N, int T, bool NOEYE>
__global__ void __launch_bounds__(kThreads)
k_playout_inc(uint4* __restrict__ states, long long G, int plies, int max_step, unsigned long long seed, long long game_id0,
              int16_t* __restrict__ trace, unsigned long long* __restrict__ plies_done, int from_reset, float* __restrict__ score_out,
              float komi, int gw, const uint4* __restrict__ states_in) {
    using namespace inc;
    // gw = games per warp (T > 1): 32 / T fills the warp; fewer leave the upper tiles empty (they idle through the lock-step votes
    // without prolonging them), which trades lanes for a shorter slowest-of-the-warp flood and more warps per scheduler
    const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int slot = (int)(threadIdx.x & 31) / T;
    const long long g = T == 1 ? gt : (gt >> 5) * gw + slot;
    const bool exists = g < G && (T == 1 || slot < gw);
    if (T > 1 ? !__any_sync(0xffffffffu, exists) : !exists) return;
    const int j = (int)(threadIdx.x & (T - 1));
    const bool lead = j == 0;
    // lane roles: j = 0..3 -> the E / W / S / N neighbour of the new stone (an enemy group), j >= 4 -> the stone itself
    const int sl = j == 0 ? 1 : (j == 2 ? DGeo<N>::STR : 0), sr = j == 1 ? 1 : (j == 3 ? DGeo<N>::STR : 0);
    // from_reset: the games start from the empty board (Reset, go_env.cc:34-41) without reading the states first
    Pos<N> p0; pos_reset(p0); p0.terminated = exists && from_reset ? 0 : 1;
    if (exists && !from_reset) pos_load<N>(p0, states_in ? states_in : states, g, G, 1);   // states_in: read there, write `states`
    Game<N> s;
    const bool first_black = p0.player == kBlack;
    s.a = to_dense<N>(bb_select(first_black, p0.stones[0], p0.stones[1]));
    s.b = to_dense<N>(bb_select(first_black, p0.stones[1], p0.stones[0]));
    s.last1 = p0.last1; s.last2 = p0.last2; s.ko_loc = p0.ko_loc; s.ko_age = p0.ko_age; s.step = p0.step; s.player = p0.player;
    s.ko_color = p0.ko_color; s.terminated = p0.terminated;
    s.l1 = from_reset ? dzero<N>() : atari_from_scratch<N, T>(s.a, s.b);
    const unsigned long long hash_head = mix32_head(seed, (unsigned long long)(game_id0 + g));
    s.draw = mix32_tail(hash_head, (unsigned long long)s.step);
    int played = 0;
    for (int t = 0; t < plies; t += 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (t + h < plies) {                                   // kernel-uniform
                int action = kInvalid;
                if (vany<T>(!s.terminated)) {
                    action = h == 0 ? ply<N, T, NOEYE>(s, s.a, s.b, max_step, hash_head, sl, sr)
                                    : ply<N, T, NOEYE>(s, s.b, s.a, max_step, hash_head, sl, sr);
                    played += action != kInvalid;
                }
                if (trace && exists && lead) trace[(long long)(t + h) * G + g] = (int16_t)action;
            }
        }
    }
    if (score_out) {                                               // kernel-uniform: getScore of the final position (go_env.cc:126-130)
        const int raw = area_score_dense<N, T>(dsel<N>(first_black, s.a, s.b), dsel<N>(first_black, s.b, s.a));
        if (exists && lead) score_out[g] = (float)raw - komi;
    }
    if (exists && lead) {
        p0.stones[0] = from_dense<N>(dsel<N>(first_black, s.a, s.b));
        p0.stones[1] = from_dense<N>(dsel<N>(first_black, s.b, s.a));
        p0.last1 = s.last1; p0.last2 = s.last2; p0.ko_loc = s.ko_loc; p0.ko_age = s.ko_age; p0.step = s.step; p0.player = s.player;
        p0.ko_color = s.ko_color; p0.terminated = s.terminated;
        pos_store<N>(p0, states, g, G, 1);
        if (plies_done && played) atomicAdd(plies_done, (unsigned long long)played);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Stepwise API on the incremental engine.  `tg_env_step` re-derives the atari classes of the groups it touches on every call; here
// the atari set travels with the position as a SIDE RECORD (one record per game: the set as a dense bit string + a 32-bit digest of
// the stones it belongs to), so a step is the legal-mask formula, at most five early-exit floods around the new stone and the
// legal mask of the next position.  A record whose digest does not match the stones (the caller changed the position behind our
// back, or never initialised it) is rebuilt from scratch inside the same launch -- the record is a cache, never a source of truth.
template <int N> struct AtariRec {
    static constexpr int NW = inc::DGeo<N>::NW;
    static constexpr int Q = (NW + 1 + 3) / 4;                        // uint4 per game: NW words of the set + 1 digest word
};
template <int N> __device__ __forceinline__ uint32_t stones_digest(const inc::DS<N>& a, const inc::DS<N>& b) {
    unsigned long long z = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int i = 0; i < inc::DGeo<N>::NW; ++i) {
        z = (z ^ a.w[i]) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 29) ^ ((unsigned long long)b.w[i] << 32)) * 0x94D049BB133111EBull;
    }
    return (uint32_t)(z >> 32) | 1u;                                  // never 0: an all-zero record of a non-empty board is stale
}

template <int N, int T>
__global__ void __launch_bounds__(kThreads)
k_step_inc(uint4* __restrict__ states, uint32_t* __restrict__ atari, long long G, const int16_t* __restrict__ actions,
           uint32_t* __restrict__ legal_out, uint8_t* __restrict__ done_out, uint8_t* __restrict__ acc_out, int max_step) {
    using namespace inc;
    using GG = Geo<N>; using D = DGeo<N>;
    constexpr int RW = AtariRec<N>::Q * 4;                            // u32 words per record
    const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / T;
    const bool exists = g < G;
    if (!__any_sync(0xffffffffu, exists)) return;
    const int j = (int)(threadIdx.x & (T - 1));
    const bool lead = j == 0;
    const int sl = j == 0 ? 1 : (j == 2 ? D::STR : 0), sr = j == 1 ? 1 : (j == 3 ? D::STR : 0);
    Pos<N> p0; pos_reset(p0); p0.terminated = 1;
    if (exists) pos_load<N>(p0, states, g, G, 1);
    Game<N> s;
    const bool first_black = p0.player == kBlack;
    s.a = to_dense<N>(bb_select(first_black, p0.stones[0], p0.stones[1]));          // the side to move
    s.b = to_dense<N>(bb_select(first_black, p0.stones[1], p0.stones[0]));
    s.last1 = p0.last1; s.last2 = p0.last2; s.ko_loc = p0.ko_loc; s.ko_age = p0.ko_age; s.step = p0.step; s.player = p0.player;
    s.ko_color = p0.ko_color; s.terminated = p0.terminated; s.draw = 0;
    // the side record: valid iff its digest matches these stones (an empty board with an all-zero record matches by convention)
    const uint32_t dig = stones_digest<N>(first_black ? s.a : s.b, first_black ? s.b : s.a);
    const bool board_empty = !dany<N>(s.a | s.b);
    uint32_t* rec = atari + (exists ? g : 0) * RW;
    TG_DW s.l1.w[i] = exists ? rec[i] : 0u;
    const uint32_t have = exists ? rec[D::NW] : 0u;
    const bool stale = exists && !(have == dig || (board_empty && have == 0u && !dany<N>(s.l1)));
    if (__any_sync(0xffffffffu, stale)) {                             // rare: rebuild in lock step, keep the valid ones
        const DS<N> fresh = atari_from_scratch<N, T>(s.a, s.b);
        s.l1 = dsel<N>(stale, fresh, s.l1);
    }
    int action = exists ? (int)actions[g] : (int)kInvalid;
    if (action == GG::P) action = kPass;                              // go_env.cc:56-57
    const bool active = exists && !s.terminated;
    bool acc = false, done = exists && s.terminated;                  // go_env.cc:52-55
    const bool is_pass = active && (action == kPass || action == kResign);
    const bool is_stone = active && action >= 0 && action < GG::P;
    const DS<N> lg0 = dlegal<N>(s, s.a, s.b);
    const int bit = is_stone ? action_to_bit<N>(action) : 0;
    bool placing = false;
    TG_DW placing = placing || (is_stone && i == (bit >> 5) && ((lg0.w[i] >> (bit & 31)) & 1u));
    dplace<N, T>(s, s.a, s.b, bit, placing, max_step, sl, sr);        // every lane walks through (votes are warp-wide)
    if (placing) { acc = true; done = s.terminated != 0; }
    if (is_pass) {                                                    // board.cc:536-542, 554-558, 656-661; ko fields untouched
        s.player = 3 - s.player; s.last2 = s.last1; s.last1 = action; s.step += 1;
        const bool over = s.step > 1 && ((s.last1 == kPass && s.last2 == kPass) || s.last1 == kResign);
        if (over || s.step > max_step) s.terminated = 1;
        acc = true; done = over || s.terminated != 0;
    }
    // after an accepted move the roles swap: the side to move is now the old `op`
    const DS<N> me2 = dsel<N>(acc, s.b, s.a), op2 = dsel<N>(acc, s.a, s.b);
    const DS<N> lg = dlegal<N>(s, me2, op2);
    if (exists && lead) {
        const bool now_black = s.player == kBlack;
        const DS<N> black = dsel<N>(now_black, me2, op2), white = dsel<N>(now_black, op2, me2);
        if (acc) {
            p0.stones[0] = from_dense<N>(black); p0.stones[1] = from_dense<N>(white);
            p0.last1 = s.last1; p0.last2 = s.last2; p0.ko_loc = s.ko_loc; p0.ko_age = s.ko_age; p0.step = s.step; p0.player = s.player;
            p0.ko_color = s.ko_color; p0.terminated = s.terminated;
            pos_store<N>(p0, states, g, G, 1);
        }
        TG_DW rec[i] = s.l1.w[i];
        rec[D::NW] = stones_digest<N>(black, white);
        if (done_out) done_out[g] = done ? 1 : 0;
        if (acc_out) acc_out[g] = acc ? 1 : 0;
        if (legal_out) {
            uint32_t d[GG::DQ * 4];
#pragma unroll
            for (int i = 0; i < GG::DQ * 4; ++i) d[i] = 0;
            bb_to_dense<N>(from_dense<N>(lg), d);
            if (!dany<N>(lg)) d[GG::P >> 5] |= 1u << (GG::P & 31);
            uint4* o = reinterpret_cast<uint4*>(legal_out) + g * GG::DQ;
#pragma unroll
            for (int k = 0; k < GG::DQ; ++k) o[k] = make_uint4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
        }
    }
}

// Row-per-lane playout (tg_rows.cuh): 31 / (N+1) games per warp, one board row per lane.  Same move rule, same results.
template <int N>
__global__ void __launch_bounds__(kThreads)
k_playout_rows(uint4* __restrict__ states, long long G, int plies, int max_step, unsigned long long seed, long long game_id0,
               int no_eye, int16_t* __restrict__ trace, unsigned long long* __restrict__ plies_done) {
    using R = RowGeo<N>;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (warp * R::GPW >= G) return;                       // whole warps leave together
    RowLane<N> L; L.init(warp, G);
    RowPos<N> p; rows_load<N>(p, L, states, G);
    int played = 0;
    for (int t = 0; t < plies; ++t) {
        const bool act = L.rowlane && !p.terminated;
        int action = kInvalid;
        if (__any_sync(kFull, act)) {                     // warp-uniform
            const uint32_t m = act ? R::ROW : 0u;
            TG_STAT(0, 1);
            uint32_t lg = rlegal<N>(p, L, m);
            if (no_eye) {                                 // kernel-uniform
                const bool me_black = p.player == kBlack;
                lg &= ~rtrue_eyes<N>((me_black ? p.b : p.w) & m, (me_black ? p.w : p.b) & m, m, L.r);
            }
            // per-game count and k-th legal point in ascending action order = (row, bit) order: one warp-wide inclusive scan
            const int pc = __popc(lg);
            int incl = pc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(kFull, incl, d); incl += (L.lane >= d) ? v : 0; }
            const int base = __shfl_sync(kFull, incl, L.lead - 1);          // separator before the game = sum of earlier games
            const int nl = __shfl_sync(kFull, incl, L.lead + N - 1) - base;
            const uint32_t rnd = mix32(seed, (unsigned long long)(game_id0 + L.g), (unsigned long long)p.step);
            const int k = (int)(((unsigned long long)rnd * (unsigned)nl) >> 32);
            const int excl = incl - pc - base;
            const bool placing = act && nl > 0;
            const bool mine = placing && k >= excl && k < excl + pc;
            const int bit = mine ? nth_bit(lg, k - excl) : 0;
            const uint32_t pb = mine ? (1u << bit) : 0u;
            const uint32_t ob = __ballot_sync(kFull, mine) & L.gm;
            const int a = __shfl_sync(kFull, L.r * N + bit, ob ? (__ffs((int)ob) - 1) : L.lane);
            rplace<N>(p, L, m, pb, placing, a);
            if (act) {
                action = placing ? a : Geo<N>::P;
                if (!placing) {                            // pass: board.cc:536-542, 554-558, 656-661
                    p.player = 3 - p.player; p.last2 = p.last1; p.last1 = kPass; p.step += 1;
                    if (p.step > 1 && p.last2 == kPass) p.terminated = 1;
                }
                if (p.step > max_step) p.terminated = 1;   // go_env.cc:67-73
                played += 1;
            }
        }
        if (trace && L.rowlane && L.r == 0) trace[(long long)t * G + L.g] = (int16_t)action;
    }
    rows_store<N>(p, L, states, G);
    if (L.rowlane && L.r == 0 && plies_done && played) atomicAdd(plies_done, (unsigned long long)played);
}

__global__ void k_sub_encode(const float* __restrict__ enc, long long B, int n, int C, int sub, int cuts, float* __restrict__ out) {
    // board.cc:1166-1271: corner crops TL, TR, BL, BR (+ centre when cuts == 5)
    const long long total = B * cuts * C * sub * sub;
    const int iv = n - sub, ter = iv / 2;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int x = (int)(e % sub); long long r = e / sub;
        int y = (int)(r % sub); r /= sub;
        int c = (int)(r % C); r /= C;
        int i = (int)(r % cuts); long long b = r / cuts;
        int x0 = (i == 1 || i == 3) ? iv : (i == 4 ? ter : 0);
        int y0 = (i == 2 || i == 3) ? iv : (i == 4 ? ter : 0);
        out[e] = enc[((b * C + c) * n + (y + y0)) * n + (x + x0)];
    }
}

// ------------------------------------------------------------------------------------------ host side
static thread_local char g_err[256] = "ok";
int set_err(int code, const char* what) {
    if (code == 0) return 0;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, code > 0 ? cudaGetErrorString((cudaError_t)code) : "bad argument");
    return code;
}
const char* last_err() { return g_err; }

int pick_tile(int tile, long long G) {
    if (tile == 1 || tile == 2 || tile == 4 || tile == 8 || tile == 16 || tile == 32) return tile;
    // enough warps to fill 148 SMs x 64 warps with whole-warp tiles? otherwise trade latency for throughput
    if (G <= 16384) return 32;
    if (G <= 131072) return 8;
    return 1;
}

#define TG_DISPATCH_T(NN, tile, FN, ...)                                 \
    switch (tile) {                                                      \
        case 32: FN<NN, 32>(__VA_ARGS__); break;                         \
        case 8:  FN<NN, 8>(__VA_ARGS__); break;                          \
        case 1:  FN<NN, 1>(__VA_ARGS__); break;                          \
        default: return set_err(-1, "tile must be 32, 8 or 1");          \
    }
#define TG_DISPATCH_T6(NN, tile, FN, ...)                                \
    switch (tile) {                                                      \
        case 32: FN<NN, 32>(__VA_ARGS__); break;                         \
        case 16: FN<NN, 16>(__VA_ARGS__); break;                         \
        case 8:  FN<NN, 8>(__VA_ARGS__); break;                          \
        case 4:  FN<NN, 4>(__VA_ARGS__); break;                          \
        case 2:  FN<NN, 2>(__VA_ARGS__); break;                          \
        case 1:  FN<NN, 1>(__VA_ARGS__); break;                          \
        case 16 | TG_TILE_UNIFORM: FN<NN, 16 | TG_TILE_UNIFORM>(__VA_ARGS__); break; \
        case 8 | TG_TILE_UNIFORM:  FN<NN, 8 | TG_TILE_UNIFORM>(__VA_ARGS__); break;  \
        case 4 | TG_TILE_UNIFORM:  FN<NN, 4 | TG_TILE_UNIFORM>(__VA_ARGS__); break;  \
        default: return set_err(-1, "tile must be a power of two <= 32"); \
    }
#define TG_DISPATCH(n, tile, FN, ...)                                    \
    if (n == 9) { TG_DISPATCH_T(9, tile, FN, __VA_ARGS__) }              \
    else if (n == 19) { TG_DISPATCH_T(19, tile, FN, __VA_ARGS__) }       \
    else return set_err(-1, "board size must be 9 or 19");

static inline int snap_tile(int tile, long long G) {
    int t = pick_tile(tile, G);
    return t >= 32 ? 32 : (t >= 8 ? 8 : 1);
}
static inline unsigned blocks_for(long long G, int T) { return (unsigned)((G * T + kThreads - 1) / kThreads); }

template <int N, int T> void launch_step(cudaStream_t st, uint4* states, long long G, const int16_t* actions, uint32_t* legal_out,
                                         uint8_t* done_out, uint8_t* acc_out, int max_step) {
    k_step<N, T><<<blocks_for(G, T), kThreads, 0, st>>>(states, G, actions, legal_out, done_out, acc_out, max_step);
}
template <int N, int T> void launch_legal(cudaStream_t st, const uint4* states, long long G, uint32_t* legal_out, int no_eye) {
    k_legal<N, T><<<blocks_for(G, T), kThreads, 0, st>>>(states, G, legal_out, no_eye);
}
template <int N, int T> void launch_encode(cudaStream_t st, const uint4* states, long long G, float* planes) {
    constexpr int GPC = kThreads / T;
    // the CTA's bit string, plus (wide tiles) one staging area of the ten bitboards per game (planes_to_smem_staged)
    const size_t smem = planes_smem_bytes<N>(GPC) + (T >= 8 ? (size_t)GPC * planes_stage_words<N>() * 4 : 0);
    static unsigned long long attr_done = 0;                 // one bit per device: the attribute is per device, not per process
    int dev = 0; cudaGetDevice(&dev);
    if (smem > 48 * 1024 && !((attr_done >> (dev & 63)) & 1ull)) {
        cudaFuncSetAttribute(k_encode<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done |= 1ull << (dev & 63);
    }
    k_encode<N, T><<<(unsigned)((G + GPC - 1) / GPC), kThreads, smem, st>>>(states, G, planes);
}
template <int N, int T> void launch_score(cudaStream_t st, const uint4* states, long long G, float komi, float* score, float* territory) {
    k_score<N, T><<<blocks_for(G, T), kThreads, 0, st>>>(states, G, komi, score, territory);
}
template <int N> void launch_playout_rows(cudaStream_t st, uint4* states, long long G, int plies, int max_step, unsigned long long seed,
                                           long long game_id0, int no_eye, int16_t* trace, unsigned long long* plies_done) {
    const long long warps = (G + RowGeo<N>::GPW - 1) / RowGeo<N>::GPW;
    constexpr int kRowThreads = 64;       // small CTAs: 1366 warps (4096 games) spread evenly over 148 SMs
    k_playout_rows<N><<<(unsigned)((warps * 32 + kRowThreads - 1) / kRowThreads), kRowThreads, 0, st>>>(states, G, plies, max_step, seed, game_id0,
                                                                                                 no_eye, trace, plies_done);
}
template <int N, int TC> void launch_playout(cudaStream_t st, uint4* states, long long G, int plies, int max_step, unsigned long long seed,
                                             long long game_id0, int no_eye, int16_t* trace, unsigned long long* plies_done) {
    constexpr int T = TC & (TG_TILE_UNIFORM - 1);  // TC = lanes per game (+ TG_TILE_UNIFORM: the tiles of a warp run in lock step)
    if constexpr ((TC & TG_TILE_UNIFORM) != 0) {
        // a playout CTA lives for the whole game: with few warps per SM one-warp CTAs spread them evenly over the 148 SMs
        const long long warps = (G * T + 31) / 32;
        const int threads = warps < 148LL * 16 ? 32 : kThreads;
        k_playout_wu<N, T><<<(unsigned)((G * T + threads - 1) / threads), threads, 0, st>>>(states, G, plies, max_step, seed, game_id0, no_eye, trace, plies_done);
    }
    else
        k_playout<N, T><<<blocks_for(G, T), kThreads, 0, st>>>(states, G, plies, max_step, seed, game_id0, no_eye, trace, plies_done);
}

static int g_playout_games_per_warp = 0;     // tg_env_debug_games_per_warp: 0 = fill the warp

template <int N, int T> int launch_playout_inc(cudaStream_t st, uint4* states, long long G, int plies, int max_step, unsigned long long seed,
                                                long long game_id0, int no_eye, int16_t* trace, unsigned long long* plies_done,
                                                int from_reset = 0, float* score_out = nullptr, float komi = 0.0f,
                                                const uint4* states_in = nullptr) {
    // a playout CTA lives for the whole game: with few warps per SM one-warp CTAs spread them evenly over the 148 SMs
    int gw = 32 / T;
    if (T > 1 && g_playout_games_per_warp > 0 && g_playout_games_per_warp < gw) gw = g_playout_games_per_warp;
    const long long warps = T == 1 ? (G + 31) / 32 : (G + gw - 1) / gw;
    const int threads = warps < 148LL * 16 ? 32 : kThreads;
    const unsigned blocks = (unsigned)((warps * 32 + threads - 1) / threads);
    if (no_eye) k_playout_inc<N, T, true><<<blocks, threads, 0, st>>>(states, G, plies, max_step, seed, game_id0, trace, plies_done, from_reset, score_out, komi, gw, states_in);
    else k_playout_inc<N, T, false><<<blocks, threads, 0, st>>>(states, G, plies, max_step, seed, game_id0, trace, plies_done, from_reset, score_out, komi, gw, states_in);
    return 0;
}

}  // namespace tg

using namespace tg;

extern "C" {

int tg_version(void) { return 100; }
const char* tg_last_error(void) { return last_err(); }
int tg_supported_board(int n) { return n == 9 || n == 19; }
int tg_state_bytes(int n) { return n == 9 ? Geo<9>::SW * 4 : (n == 19 ? Geo<19>::SW * 4 : 0); }
int tg_legal_words(int n) { return n == 9 ? Geo<9>::DQ * 4 : (n == 19 ? Geo<19>::DQ * 4 : 0); }

int tg_env_reset(int n, void* states, int64_t G, void* stream) {
    if (!states || G <= 0) return set_err(-1, "tg_env_reset");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned nb = (unsigned)((G + 255) / 256);
    if (n == 9) k_reset<9><<<nb, 256, 0, st>>>((uint4*)states, G);
    else if (n == 19) k_reset<19><<<nb, 256, 0, st>>>((uint4*)states, G);
    else return set_err(-1, "board size must be 9 or 19");
    return set_err((int)cudaGetLastError(), "tg_env_reset");
}

int tg_env_step(int n, void* states, int64_t G, const int16_t* actions, uint32_t* legal_out, uint8_t* done_out,
                uint8_t* acc_out, int max_step, int tile, void* stream) {
    if (!states || !actions || G <= 0) return set_err(-1, "tg_env_step");
    cudaStream_t st = (cudaStream_t)stream;
    tile = snap_tile(tile, G);
    TG_DISPATCH(n, tile, launch_step, st, (uint4*)states, G, actions, legal_out, done_out, acc_out, max_step);
    return set_err((int)cudaGetLastError(), "tg_env_step");
}

int tg_env_atari_words(int n) { return n == 9 ? AtariRec<9>::Q * 4 : (n == 19 ? AtariRec<19>::Q * 4 : 0); }

int tg_env_step_inc(int n, void* states, uint32_t* atari, int64_t G, const int16_t* actions, uint32_t* legal_out, uint8_t* done_out,
                    uint8_t* accepted_out, int max_step, void* stream) {
    if (!states || !atari || !actions || G <= 0) return set_err(-1, "tg_env_step_inc");
    constexpr int T = 4;
    const unsigned blocks = (unsigned)((G * T + kThreads - 1) / kThreads);
    if (n == 9) k_step_inc<9, T><<<blocks, kThreads, 0, (cudaStream_t)stream>>>((uint4*)states, atari, G, actions, legal_out, done_out, accepted_out, max_step);
    else if (n == 19) k_step_inc<19, T><<<blocks, kThreads, 0, (cudaStream_t)stream>>>((uint4*)states, atari, G, actions, legal_out, done_out, accepted_out, max_step);
    else return set_err(-1, "board size must be 9 or 19");
    return set_err((int)cudaGetLastError(), "tg_env_step_inc");
}

int tg_env_legal(int n, const void* states, int64_t G, uint32_t* legal_out, int no_eye, int tile, void* stream) {
    if (!states || !legal_out || G <= 0) return set_err(-1, "tg_env_legal");
    cudaStream_t st = (cudaStream_t)stream;
    tile = snap_tile(tile, G);
    TG_DISPATCH(n, tile, launch_legal, st, (const uint4*)states, G, legal_out, no_eye);
    return set_err((int)cudaGetLastError(), "tg_env_legal");
}

int tg_env_encode(int n, const void* states, int64_t G, float* planes, int tile, void* stream) {
    if (!states || !planes || G <= 0) return set_err(-1, "tg_env_encode");
    if ((reinterpret_cast<uintptr_t>(planes) & 3) != 0) return set_err(-1, "tg_env_encode: planes must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    tile = snap_tile(tile, G);
    TG_DISPATCH(n, tile, launch_encode, st, (const uint4*)states, G, planes);
    return set_err((int)cudaGetLastError(), "tg_env_encode");
}

int tg_env_score(int n, const void* states, int64_t G, float komi, float* score, float* territory, int tile, void* stream) {
    if (!states || G <= 0 || (!score && !territory)) return set_err(-1, "tg_env_score");
    cudaStream_t st = (cudaStream_t)stream;
    tile = snap_tile(tile, G);
    TG_DISPATCH(n, tile, launch_score, st, (const uint4*)states, G, komi, score, territory);
    return set_err((int)cudaGetLastError(), "tg_env_score");
}

/* the layout tg_env_playout runs for (n, G, tile): TG_TILE_ROWS or lanes per game (| TG_TILE_UNIFORM | TG_TILE_INC) */
int tg_env_playout_tile(int n, int64_t G, int tile) {
    if (tile == 0) {
        // measured on B200 (profiles/r02_playout_layouts.txt): the incremental engine wins at every size; lanes per game shrink
        // as the batch grows.  4 lanes per game = 8 games per warp: at the BASELINE batch (4096 games = 512 warps) every
        // scheduler holds at most one warp and the launch lasts one warp's 120-ply dependency chain.
        if (n == 19) return (G < 768 ? 32 : G < 1536 ? 16 : G < 3072 ? 8 : G < 12288 ? 4 : 1) | TG_TILE_INC;
        return (G < 3072 ? 8 : G < 24576 ? 4 : 1) | TG_TILE_INC;
    }
    if (tile > 0 && (tile & TG_TILE_INC)) {
        const int t = tile & (TG_TILE_UNIFORM - 1);
        return (t == 32 || t == 16 || t == 8 || t == 4 || t == 1) ? (t | TG_TILE_INC) : tg_env_playout_tile(n, G, 0);
    }
    if (tile > 0 && (tile & TG_TILE_UNIFORM)) {
        const int t = tile & (TG_TILE_UNIFORM - 1);
        return (t == 16 || t == 8 || t == 4) ? tile : pick_tile(0, G);
    }
    return tile == TG_TILE_ROWS ? tile : pick_tile(tile, G);
}

static int playout_impl(int n, void* states, int64_t G, int plies, int max_step, uint64_t seed, int64_t game_id0, int no_eye,
                        int16_t* trace_actions, unsigned long long* plies_done, int tile, void* stream, int from_reset, float* score_out, float komi,
                        const void* states_in = nullptr) {
    if (!states || G <= 0 || plies < 0) return set_err(-1, "tg_env_playout");
    cudaStream_t st = (cudaStream_t)stream;
    if (states_in && !(tg_env_playout_tile(n, G, tile) > 0 && (tg_env_playout_tile(n, G, tile) & TG_TILE_INC))) {
        // the other layouts play in place: bring the input over first (either side may be pinned host memory)
        const cudaError_t e = cudaMemcpyAsync(states, states_in, (size_t)G * tg_state_bytes(n), cudaMemcpyDefault, st);
        if (e != cudaSuccess) return set_err((int)e, "tg_env_playout_io (copy)");
        states_in = nullptr;
    }
    // tile == 0, measured on B200 (profiles/r01_playout_layouts.txt): 19x19 -> rows layout at every batch size (up to 3.6x over
    // the bitboard tiles); 9x9 -> four games per warp in lock step while the batch cannot fill the SMs, one thread per game beyond
    tile = tg_env_playout_tile(n, G, tile);
    if ((from_reset || score_out) && !(tile > 0 && (tile & TG_TILE_INC))) {       // the other layouts: the same result in separate launches
        if (from_reset) { const int rc = tg_env_reset(n, states, G, stream); if (rc) return rc; }
        const int rc = playout_impl(n, states, G, plies, max_step, seed, game_id0, no_eye, trace_actions, plies_done, tile, stream, 0, nullptr, 0.0f);
        if (rc || !score_out) return rc;
        return tg_env_score(n, states, G, komi, score_out, nullptr, 0, stream);
    }
    if (tile == TG_TILE_ROWS) {
        if (n == 9) launch_playout_rows<9>(st, (uint4*)states, G, plies, max_step, seed, game_id0, no_eye, trace_actions, plies_done);
        else if (n == 19) launch_playout_rows<19>(st, (uint4*)states, G, plies, max_step, seed, game_id0, no_eye, trace_actions, plies_done);
        else return set_err(-1, "board size must be 9 or 19");
        return set_err((int)cudaGetLastError(), "tg_env_playout");
    }
    if (tile & TG_TILE_INC) {
        const int t = tile & (TG_TILE_UNIFORM - 1);
#define TG_INC_CASE(NN, TT) launch_playout_inc<NN, TT>(st, (uint4*)states, G, plies, max_step, seed, game_id0, no_eye, trace_actions, plies_done, from_reset, score_out, komi, (const uint4*)states_in)
        if (n == 9) { if (t == 32) TG_INC_CASE(9, 32); else if (t == 16) TG_INC_CASE(9, 16); else if (t == 8) TG_INC_CASE(9, 8); else if (t == 4) TG_INC_CASE(9, 4); else TG_INC_CASE(9, 1); }
        else if (n == 19) { if (t == 32) TG_INC_CASE(19, 32); else if (t == 16) TG_INC_CASE(19, 16); else if (t == 8) TG_INC_CASE(19, 8); else if (t == 4) TG_INC_CASE(19, 4); else TG_INC_CASE(19, 1); }
        else return set_err(-1, "board size must be 9 or 19");
#undef TG_INC_CASE
        return set_err((int)cudaGetLastError(), "tg_env_playout");
    }
    if (n == 9) { TG_DISPATCH_T6(9, tile, launch_playout, st, (uint4*)states, G, plies, max_step, seed, game_id0, no_eye, trace_actions, plies_done) }
    else if (n == 19) { TG_DISPATCH_T6(19, tile, launch_playout, st, (uint4*)states, G, plies, max_step, seed, game_id0, no_eye, trace_actions, plies_done) }
    else return set_err(-1, "board size must be 9 or 19");
    return set_err((int)cudaGetLastError(), "tg_env_playout");
}

int tg_env_debug_games_per_warp(int gw) { const int old = g_playout_games_per_warp; g_playout_games_per_warp = gw; return old; }

int tg_env_playout(int n, void* states, int64_t G, int plies, int max_step, uint64_t seed, int64_t game_id0, int no_eye,
                   int16_t* trace_actions, unsigned long long* plies_done, int tile, void* stream) {
    return playout_impl(n, states, G, plies, max_step, seed, game_id0, no_eye, trace_actions, plies_done, tile, stream, 0, nullptr, 0.0f);
}

int tg_env_playout_scored(int n, void* states, int64_t G, int plies, int max_step, uint64_t seed, int64_t game_id0, int no_eye,
                          int16_t* trace_actions, unsigned long long* plies_done, int from_reset, float* score_out, float komi,
                          int tile, void* stream) {
    return playout_impl(n, states, G, plies, max_step, seed, game_id0, no_eye, trace_actions, plies_done, tile, stream, from_reset, score_out, komi);
}

int tg_env_playout_io(int n, const void* states_in, void* states_out, int64_t G, int plies, int max_step, uint64_t seed, int64_t game_id0,
                      int no_eye, int16_t* trace_actions, unsigned long long* plies_done, float* score_out, float komi, int tile, void* stream) {
    if (!states_in) return set_err(-1, "tg_env_playout_io");
    return playout_impl(n, states_out, G, plies, max_step, seed, game_id0, no_eye, trace_actions, plies_done, tile, stream, 0, score_out, komi,
                        states_in == states_out ? nullptr : states_in);
}

#ifdef TG_ROWS_STATS
int tg_debug_rows_stats(unsigned long long* out8, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, g_rows_stats, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_rows_stats, z, sizeof(z)); }
    return 0;
}
#endif

int tg_sub_encode(int n, const float* enc, int64_t B, int channels, int sub, int cuts, float* out, void* stream) {
    if (!enc || !out || B <= 0 || sub <= 0 || sub > n || cuts < 1 || cuts > 5) return set_err(-1, "tg_sub_encode");
    long long total = (long long)B * cuts * channels * sub * sub;
    unsigned nb = (unsigned)((total + 255) / 256); if (nb > 148 * 16) nb = 148 * 16;
    k_sub_encode<<<nb, 256, 0, (cudaStream_t)stream>>>(enc, B, n, channels, sub, cuts, out);
    return set_err((int)cudaGetLastError(), "tg_sub_encode");
}

}  // extern "C"
