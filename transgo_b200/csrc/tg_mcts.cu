// tg_mcts.cu -- batched MCTS kernels (one warp per game) + the tg_mcts_* C ABI of include/transgo_b200.h.
// Replaces the pure-Python tree search self_play.py: Node_V :51-96 and WP_MCTS :575-875.
//
// Arithmetic contract (SURVEY.md 8a-3 dtype ledger, numpy 2.x semantics of the reference): every float operation is an
// explicit round-to-nearest intrinsic in the reference's order and width; the file is also built with -fmad=false.
#include <cmath>
#include <cstdio>
#include "../../include/transgo_b200.h"
#include "tg_board.cuh"
#include "tg_common.cuh"

namespace tg {

enum { H_BUF = 0, H_NNODES, H_NEDGES, H_ROOTN, H_ROOTONS, H_ROOTW, H_ROOTMEAN, H_ROOTVAR, H_TIE, H_N0, H_MOVE, H_FLAGS,
       H_NPATHS, H_SIMS, H_STAT_DESC, H_STAT_DEPTH, H_STAT_FAN, H_SPARE1, H_WORDS_USED, H_WORDS = 20 };   // 80 B per game
constexpr unsigned kErrAction = 256u;
constexpr int kMaxReadouts = 8;
constexpr int kParWarps = 4;        // k_select_par: one warp per descent, up to this many (the reference default is 4 readouts)

// Edge records (round 2 layout: the selection scan reads only what changes its arithmetic per wave)
//   hot  uint4 {x: cp   = f32(c1 * prior)            -- the first factor of prior_score (self_play.py:718), fixed at expansion
//               y: negq = -(W / (N + 1)) in f32      -- value_score (:724, :67-68), refreshed by every backup
//               z: vs   = f32(c2 * sqrtf(1 + clip(var, 0, 3)))  -- var_score of a VISITED child (:721-722), refreshed by every backup
//               w: N[31:16] | ons[15:9] | action[8:0]}
//   cold uint4 {x: W f32, y: var f32 (Welford M2), z: prior f32 as expanded, w: 0}   -- read by backups / root statistics only
//   kid  uint2 {x: child node (0xffffffff: none), y: edge base | edge count << 22 of the child node}
// The running mean of Node_V.value_mean_var (:84-88) is 0 before the first visit and W / (N + 1) = -negq afterwards.
__device__ __forceinline__ int pk_n(uint32_t p) { return (int)(p >> 16); }
__device__ __forceinline__ int pk_ons(uint32_t p) { return (int)((p >> 9) & 127u); }
__device__ __forceinline__ int pk_act(uint32_t p) { return (int)(p & 511u); }

struct View {
    // config
    int G, R, num_sim, wu, max_step, NC, EC, MAXD, tie_mode;
    double c1, c2;
    float komi;
    unsigned long long tie_seed;
    long long game_id0;
    // workspace
    int* hdr;               // [G][H_WORDS]
    uint4* root_state;      // [SQ][G]
    double* root_prior64;   // [G][A]
    uint4* hot;             // [2][G][EC]
    uint4* cold;            // [2][G][EC]
    uint2* kid;             // [2][G][EC]  child node, edge base | edge count << 22 of the child node (valid when the child exists)
    uint4* nodes;           // [2][G][NC][SQ+1]
    int* paths;             // [G][R][MAXD]
    int* path_len;          // [G][R]
    int* leaf_parent;       // [G][R]
    int* leaf_action;       // [G][R]
    uint4* leaf_state;      // [SQ][G*R]
    uint32_t* leaf_legal;   // [G*R][LW]
    size_t total_bytes;
};

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

template <int N>
static View make_view(const TgMctsConfig* c, void* ws) {
    using GG = Geo<N>;
    View v{};
    v.G = c->num_games; v.R = c->readouts; v.num_sim = c->num_simulation; v.wu = c->wu_loss; v.max_step = c->max_step;
    v.NC = c->node_cap; v.EC = c->edge_cap; v.MAXD = c->max_depth; v.tie_mode = c->tie_mode;
    v.c1 = c->c_puct1; v.c2 = c->c_puct2; v.komi = c->komi; v.tie_seed = c->tie_seed; v.game_id0 = c->game_id0;
    char* p = (char*)ws; size_t off = 0;
    auto take = [&](size_t bytes) { char* q = p ? p + off : nullptr; off += align_up(bytes); return q; };
    const size_t G = v.G, R = v.R;
    v.hdr = (int*)take(G * H_WORDS * 4);
    v.root_state = (uint4*)take(G * GG::SQ * 16);
    v.root_prior64 = (double*)take(G * GG::A * 8);
    v.hot = (uint4*)take(2 * G * (size_t)v.EC * 16);
    v.cold = (uint4*)take(2 * G * (size_t)v.EC * 16);
    v.kid = (uint2*)take(2 * G * (size_t)v.EC * 8);
    v.nodes = (uint4*)take(2 * G * (size_t)v.NC * (GG::SQ + 1) * 16);
    v.paths = (int*)take(G * R * (size_t)v.MAXD * 4);
    v.path_len = (int*)take(G * R * 4);
    v.leaf_parent = (int*)take(G * R * 4);
    v.leaf_action = (int*)take(G * R * 4);
    v.leaf_state = (uint4*)take(G * R * GG::SQ * 16);
    v.leaf_legal = (uint32_t*)take(G * R * GG::DQ * 16);
    v.total_bytes = off;
    return v;
}

template <int N> __device__ __forceinline__ uint4* hot_of(const View& v, int buf, int g) { return v.hot + ((size_t)buf * v.G + g) * v.EC; }
template <int N> __device__ __forceinline__ uint4* cold_of(const View& v, int buf, int g) { return v.cold + ((size_t)buf * v.G + g) * v.EC; }
template <int N> __device__ __forceinline__ uint2* kid_of(const View& v, int buf, int g) { return v.kid + ((size_t)buf * v.G + g) * v.EC; }
constexpr uint32_t kNoChild = 0xffffffffu;

// a fresh edge (Node_V.expand :70-77, complete_update :743-752): never visited, value_sum = w0
__device__ __forceinline__ void write_fresh_edge(uint4* hot, uint4* cold, uint2* kid, int e, float c1f, float prior, float w0, int action) {
    hot[e] = make_uint4(__float_as_uint(__fmul_rn(c1f, prior)), __float_as_uint(-__fdiv_rn(w0, 1.0f)), 0u, (uint32_t)action);
    cold[e] = make_uint4(__float_as_uint(w0), 0u, __float_as_uint(prior), 0u);
    kid[e] = make_uint2(kNoChild, 0u);
}
// one backup step on an edge (backpropagate :758-764 + value_mean_var :84-88); q / c are the loaded records
__device__ __forceinline__ void backup_edge(uint4& q, uint4& c, float sv, float c2f, unsigned& err) {
    const int n0 = pk_n(q.w);
    const float W = __fadd_rn(__uint_as_float(c.x), sv);
    const int n = n0 + 1;
    if (n > 65535) err |= TG_MCTS_ERR_COUNT;
    const float t = n0 == 0 ? 0.0f : -__uint_as_float(q.y);          // the running mean before this visit
    const float mean = __fdiv_rn(W, (float)(n + 1));
    const float var = __fadd_rn(__uint_as_float(c.y), __fmul_rn(__fsub_rn(sv, t), __fsub_rn(sv, mean)));
    const float cv = var < 0.0f ? 0.0f : (var > 3.0f ? 3.0f : var);  // np.clip(var, 0, 3)
    c.x = __float_as_uint(W); c.y = __float_as_uint(var);
    q.y = __float_as_uint(-mean);
    q.z = __float_as_uint(__fmul_rn(c2f, __fsqrt_rn(__fadd_rn(1.0f, cv))));
    q.w += 1u << 16;
}
__device__ __forceinline__ uint32_t pack_kid(int eb, int K) { return (uint32_t)eb | ((uint32_t)K << 22); }
template <int N> __device__ __forceinline__ uint4* nodes_of(const View& v, int buf, int g) { return v.nodes + ((size_t)buf * v.G + g) * v.NC * (Geo<N>::SQ + 1); }

__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        double y = __shfl_xor_sync(0xffffffffu, x, s);
        x = y > x ? y : x;
    }
    return x;
}

// value backup along a recorded path (self_play.py:758-764 + Node_V.value_mean_var :84-88); lane-parallel: edge i of a
// path of `len` edges receives v * (-1)^(len-1-i), the root itself v * (-1)^len.
// ons_delta (a multiple of wu_loss, usually -wu_loss) is applied to the same nodes in the same pass: revert_incomplete_update
// (:770-774) and backpropagate touch exactly the same records.
template <int N>
__device__ __forceinline__ void backup(const View& v, int* h, uint4* hot, uint4* cold, const int* path, int len, float value, unsigned& err,
                                       int ons_delta = 0) {
    const int lane = threadIdx.x & 31;
    const float c2f = (float)v.c2;
    for (int i = lane; i <= len; i += 32) {
        const float sv = ((len - 1 - i) & 1) ? -value : value;       // i == len -> (-1)^(-1) handled below
        if (i == len) {
            const float rv = (len & 1) ? -value : value;
            const float W = __fadd_rn(__int_as_float(h[H_ROOTW]), rv);
            const int n = h[H_ROOTN] + 1;
            const float t = __int_as_float(h[H_ROOTMEAN]);
            const float mean = __fdiv_rn(W, (float)(n + 1));
            const float var = __fadd_rn(__int_as_float(h[H_ROOTVAR]), __fmul_rn(__fsub_rn(rv, t), __fsub_rn(rv, mean)));
            h[H_ROOTW] = __float_as_int(W); h[H_ROOTN] = n; h[H_ROOTMEAN] = __float_as_int(mean); h[H_ROOTVAR] = __float_as_int(var);
            h[H_ROOTONS] += ons_delta;
        } else {
            const int e = path[i];
            uint4 q = hot[e];
            uint4 c = cold[e];
            backup_edge(q, c, sv, c2f, err);
            q.w = (uint32_t)((int)q.w + ons_delta * 512);
            hot[e] = q; cold[e] = c;
        }
    }
    __syncwarp();
}

// ons += d * wu_loss on every node of the path, root included (self_play.py:767-774)
__device__ __forceinline__ void add_ons(int* h, uint4* hot, const int* path, int len, int d) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i <= len; i += 32) {
        if (i == len) h[H_ROOTONS] += d;
        else {
            uint32_t* w = &reinterpret_cast<uint32_t*>(hot + path[i])[3];
            *w = (uint32_t)((int)*w + d * 512);
        }
    }
    __syncwarp();
}

// One leaf of a wave, one warp: env.step from the parent (self_play.py:629), getLegalAction (:634, :737) and encode (:798)
// fused; the planes go straight into the net's input tensor.  sm: this warp's leaf_smem_words<N>() words of shared memory.
template <int N> __host__ __device__ constexpr int leaf_smem_words() { return (Geo<N>::PLANE_BITS + 31) / 32 + 2 + planes_stage_words<N>(); }

template <int N>
__device__ __forceinline__ void leaf_work(const View& v, const uint4* nodes, size_t slot, int parent, int action, float* __restrict__ obs,
                                          uint32_t* sm, int lane) {
    using GG = Geo<N>;
    Pos<N> s; pos_load<N>(s, nodes + (size_t)parent * (GG::SQ + 1), 0, 1, 0);
    // an action the search took from the parent's legal list (trusted): only the groups touching the new stone matter for
    // captures / ko, so only those are analysed (tg_board.cuh: apply_stone_lazy)
    if (action == GG::P) apply_pass<N>(s, kPass, v.max_step);
    else apply_stone_lazy<N, 32>(s, bb_point<N>(action), action, bb_zero<N>(), bb_zero<N>(), bb_zero<N>(), v.max_step);
    BB<N> pl[10];
    Groups<N> g2;
    feature_planes_g<N, 32>(s, pl, g2);
    const BB<N> lg = legal_points(s, g2);
    constexpr int WPW = (GG::PLANE_BITS + 31) / 32 + 2;
    for (int i = lane; i < WPW; i += 32) sm[i] = 0;
    __syncwarp();
    planes_to_smem_staged<N, 32>(pl, sm, 0, sm + WPW);
    __syncwarp();
    planes_smem_to_global_part<N>(sm, obs + slot * GG::PLANE_BITS, 1, lane, 32);
    if (lane == 0) {
        pos_store<N>(s, v.leaf_state, (long long)slot, (long long)v.G * v.R, 1);
        uint32_t d[GG::DQ * 4];
#pragma unroll
        for (int i = 0; i < GG::DQ * 4; ++i) d[i] = 0;
        bb_to_dense<N>(lg, d);
        if (!any(lg)) d[GG::P >> 5] |= 1u << (GG::P & 31);
        uint4* o = reinterpret_cast<uint4*>(v.leaf_legal) + slot * GG::DQ;
#pragma unroll
        for (int k = 0; k < GG::DQ; ++k) o[k] = make_uint4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
    }
}

// ---------------------------------------------------------------------------------------------- selection
// order-preserving map double -> u64 (finite values and infinities; a NaN maps above everything and is caught by the
// equality test afterwards), so that the warp arg-max is two REDUX.MAX instead of a five-step shuffle butterfly
__device__ __forceinline__ unsigned long long ordered_bits(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// sqrt(N_parent + ons_parent) as numpy computes it (np.sqrt of a Python int: correctly rounded float64): a table for the
// counts a search meets, filled by the host's correctly rounded sqrt at library start-up; __dsqrt_rn beyond.  The index is
// warp-uniform, so the constant cache broadcasts it.
constexpr int kSqrtTab = 4096;
__constant__ double c_sqrt_tab[kSqrtTab];
__device__ __forceinline__ double sqrt_of_count(int n) { return n < kSqrtTab ? c_sqrt_tab[n] : __dsqrt_rn((double)n); }

// One node visit of select_child (self_play.py:706-725), split in two so that the loads can be in flight while the warp
// still waits for something else.  Lane l scans children l, l+32, ...
template <int N> struct ScanRegs {
    static constexpr int MAXR = (Geo<N>::A + 31) / 32;
    uint4 q[MAXR]; uint32_t ch[MAXR], kd[MAXR]; double pr64[MAXR];
};

// every load of the scan is issued before the first use (no branch in between): one memory round trip per node
template <int N>
__device__ __forceinline__ void scan_load(ScanRegs<N>& s, const uint4* hot, const uint2* kid, const double* p64,
                                          bool noisy_root, int eb, int K, int lane) {
#pragma unroll
    for (int r = 0; r < ScanRegs<N>::MAXR; ++r) {
        if (N == 9 || 32 * r < K) {                           // 19x19: skip whole rounds beyond K (warp-uniform)
            const int k = lane + 32 * r;
            const int kk = k < K ? k : K - 1;
            s.q[r] = hot[eb + kk];
            const uint2 kq = kid[eb + kk];                    // child node of the edge and where the child's edges live: loaded
            s.ch[r] = kq.x; s.kd[r] = kq.y;                   // with the scan, so that the descent needs no second trip
            s.pr64[r] = noisy_root ? p64[kk] : 0.0;
        } else { s.q[r] = make_uint4(0, 0, 0, 0); s.ch[r] = 0xffffffffu; s.kd[r] = 0; s.pr64[r] = 0.0; }
    }
}

// scores, arg-max, exact-tie rule.  xo[r] = virtual-loss units (multiples of wu_loss are added to the child's ons) that earlier
// descents of the same wave put on this lane's r-th child and that memory does not hold yet (nullptr: none).
// Returns false when no child has a finite score (NaN from the evaluator).  tie_events counts exact ties seen.
template <int N>
__device__ __forceinline__ bool scan_pick(const View& v, const ScanRegs<N>& s, bool noisy_root, int K, int Np, int onsp, const int* xo,
                                          int tie_counter, int& tie_events, int g, int lane,
                                          int& ksel, uint32_t& pk_sel, uint32_t& ch_sel, uint32_t& kd_sel) {
    constexpr int MAXR = ScanRegs<N>::MAXR;
    const double sq = sqrt_of_count(Np + onsp);
    unsigned long long ob[MAXR];
    unsigned hi_max = 0;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        const int k = lane + 32 * r;
        const int nc = pk_n(s.q[r].w), oc = pk_ons(s.q[r].w) + (xo ? xo[r] * v.wu : 0);
        // prior_score = c1 * prior * sqrt(Np + ons_p) / (Nc + ons_c + 1)          (:718-719); f32(c1 * prior) comes with the record
        double ps = noisy_root ? __dmul_rn(v.c1, s.pr64[r]) : (double)__uint_as_float(s.q[r].x);
        ps = __ddiv_rn(__dmul_rn(ps, sq), (double)(nc + oc + 1));
        // var_score = c2 * sqrt(1 + clip(var, 0, 3)) in f32 for a visited child (kept in the record by every backup), the
        // float64 constant for a never-visited one (:721-722); value_score = -child.value() in f32 (:724, :67-68), also kept
        const double vs = nc == 0 ? v.c2 : (double)__uint_as_float(s.q[r].z);
        const double qv = (double)__uint_as_float(s.q[r].y);
        const double sc = __dadd_rn(__dadd_rn(ps, vs), qv);
        ob[r] = k < K ? ordered_bits(sc) : 0ull;
        const unsigned hi = (unsigned)(ob[r] >> 32);
        hi_max = hi > hi_max ? hi : hi_max;
    }
    hi_max = __reduce_max_sync(0xffffffffu, hi_max);
    unsigned lo_max = 0;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        const unsigned lo = (unsigned)(ob[r] >> 32) == hi_max ? (unsigned)ob[r] : 0u;
        lo_max = lo > lo_max ? lo : lo_max;
    }
    lo_max = __reduce_max_sync(0xffffffffu, lo_max);
    const unsigned long long best = ((unsigned long long)hi_max << 32) | lo_max;
    // exact ties, ascending action order (:708-713).  A NaN score (ordered above +inf or below -inf, never equal to a real
    // one) or an all-NaN node leaves no candidate: fail loudly instead of walking a bogus edge.
    if (best <= 0x000fffffffffffffull || best >= 0xfff0000000000000ull) return false;      // the maximum is not a finite number
    unsigned bal[MAXR]; int total = 0;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) { bal[r] = __ballot_sync(0xffffffffu, ob[r] == best); total += __popc(bal[r]); }
    int j = 0;
    if (total > 1) {
        if (v.tie_mode == 1) j = (int)(((unsigned long long)mix32(v.tie_seed, (unsigned long long)(v.game_id0 + g), (unsigned long long)tie_counter) * (unsigned)total) >> 32);
        tie_events += 1;
    }
    ksel = -1; ch_sel = 0xffffffffu; kd_sel = 0;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        const int c = __popc(bal[r]);
        if (ksel < 0 && j < c) {
            const int src = nth_bit(bal[r], j);
            ksel = src + 32 * r;
            pk_sel = __shfl_sync(0xffffffffu, s.q[r].w, src);
            ch_sel = __shfl_sync(0xffffffffu, s.ch[r], src);
            kd_sel = __shfl_sync(0xffffffffu, s.kd[r], src);
        } else if (ksel < 0) j -= c;
    }
    return true;
}

// terminal leaf (self_play.py:638-642): step, score, back the result up at once
template <int N>
__device__ __forceinline__ void terminal_backup(const View& v, int* h, uint4* hot, uint4* cold, const uint4* nodes, int node, int action,
                                                const int* path, int depth, unsigned& err) {
    using GG = Geo<N>;
    Pos<N> s; pos_load<N>(s, nodes + (size_t)node * (GG::SQ + 1), 0, 1, 0);
    Groups<N> gr;
    if (action != GG::P) analyse_groups<N, 32, false>(s.stones[0], s.stones[1], s.player, gr);
    else { gr.l1[0] = gr.l1[1] = gr.l2[0] = gr.l2[1] = gr.lives = bb_zero<N>(); }
    bool acc;
    apply_move<N, 32>(s, gr, action, v.max_step, acc, true);
    BB<N> bt, wt;
    const int raw = area_score<N, 32>(s.stones[0], s.stones[1], bt, wt);
    const int winner = ((float)raw - v.komi) > 0.0f ? kBlack : kWhite;          // environment.py:118-119
    const float value = (s.player == winner) ? 1.0f : -1.0f;
    backup<N>(v, h, hot, cold, path, depth, value, err);
}

// The wave of one game, one descent after the other, exactly as WP_MCTS.run does it (self_play.py:616-646).  One warp.
template <int N>
__device__ __forceinline__ void select_sequential(const View& v, int g, int lane) {
    using GG = Geo<N>;
    int* h = v.hdr + (size_t)g * H_WORDS;
    const unsigned flags = (unsigned)h[H_FLAGS];
    const int buf = h[H_BUF];
    uint4* hot = hot_of<N>(v, buf, g);
    uint4* cold = cold_of<N>(v, buf, g);
    const uint2* kid = kid_of<N>(v, buf, g);
    const uint4* nodes = nodes_of<N>(v, buf, g);
    const uint4 root_info = nodes[GG::SQ];                            // the root's edges: read once per wave
    const uint4 root_meta = nodes[GG::SQ - 1];                        // last1 | last2, ko, step | player | ko_color, terminated
    const uint32_t root_kid = pack_kid((int)root_info.x, (int)root_info.y);
    const bool root_after_pass = (int16_t)(root_meta.x & 0xffffu) == (int16_t)kPass;
    const int root_step = (int)(root_meta.z & 0xffffu);
    const double* p64 = v.root_prior64 + (size_t)g * GG::A;
    const bool noisy = (flags & TG_MCTS_NOISY) != 0;
    int npaths = 0, tie = h[H_TIE], sims = 0;
    int st_desc = 0, st_depth = 0, st_fan = 0;                        // descents, edges walked, children scanned (bench: d, K)
    unsigned err = 0;
    for (int attempt = 0; attempt < 2 * v.R && npaths < v.R && !err; ++attempt) {      // self_play.py:616-619
        int* path = v.paths + ((size_t)g * v.R + npaths) * v.MAXD;
        int node = 0, depth = 0;
        int Np = h[H_ROOTN], onsp = h[H_ROOTONS];
        uint32_t pk_sel = 0, kid_cur = root_kid;
        bool after_pass = root_after_pass;                            // the move that led to `node` was a pass
        while (true) {                                                                // :623-627
            // where the node's edges live came with the edge that led here: no dependent load of the node record
            const int eb = (int)(kid_cur & 0x3fffffu), K = (int)(kid_cur >> 22);
            st_fan += K;
            ScanRegs<N> sr;
            scan_load<N>(sr, hot, kid, p64, noisy && node == 0, eb, K, lane);
            int ksel; uint32_t ch_sel, kd_sel; int ties = 0;
            if (!scan_pick<N>(v, sr, noisy && node == 0, K, Np, onsp, nullptr, tie, ties, g, lane, ksel, pk_sel, ch_sel, kd_sel)) { err |= TG_MCTS_ERR_NAN; break; }
            tie += ties;
            const int e = eb + ksel;
            if (depth >= v.MAXD) { err |= TG_MCTS_ERR_DEPTH; break; }
            if (lane == 0) path[depth] = e;
            depth += 1;
            const int child = (int)ch_sel;
            if (child < 0) break;
            after_pass = pk_act(pk_sel) == GG::P;
            Np = pk_n(pk_sel); onsp = pk_ons(pk_sel); node = child; kid_cur = kd_sel;
        }
        if (err) break;
        st_desc += 1; st_depth += depth;
        __syncwarp();
        // the leaf: step from its parent's state (:629).  `done` follows from what the descent already knows: every move
        // (a pass too) advances step_count by one, and the parent's last move is the edge that led to it.
        const int action = pk_act(pk_sel);
        const int step2 = root_step + depth;
        const bool done = (action == GG::P && after_pass && step2 > 1) || step2 > v.max_step;
        if (done) {                                                                   // :638-642
            terminal_backup<N>(v, h, hot, cold, nodes, node, action, path, depth, err);
            sims += 1;
        } else {
            if (lane == 0) {
                const size_t slot = (size_t)g * v.R + npaths;
                v.leaf_parent[slot] = node; v.leaf_action[slot] = action; v.path_len[slot] = depth;
            }
            add_ons(h, hot, path, depth, v.wu);                                        // :644
            npaths += 1;
        }
    }
    if (lane == 0) {
        h[H_TIE] = tie; h[H_NPATHS] = npaths; h[H_SIMS] += sims;
        h[H_STAT_DESC] += st_desc; h[H_STAT_DEPTH] += st_depth; h[H_STAT_FAN] += st_fan;
        if (err) h[H_FLAGS] = (int)(flags | err);
    }
}

// out-of-line copy for the (rare) replay inside k_select_par: keeps that kernel's hot path small
template <int N> __device__ __noinline__ void select_sequential_cold(const View& v, int g, int lane) { select_sequential<N>(v, g, lane); }

template <int N> __device__ __forceinline__ bool game_runs(const View& v, const int* h) {
    const unsigned flags = (unsigned)h[H_FLAGS];
    return (flags & TG_MCTS_ACTIVE) && !(flags & (TG_MCTS_NEED_EXPAND | TG_MCTS_ERR_ANY | kErrAction)) && h[H_ROOTN] < h[H_N0] + v.num_sim;
}

// one warp per game, descents one after the other (reference order); also the fallback of k_select_par
template <int N>
__global__ void __launch_bounds__(kThreads) k_select(View v) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= v.G) return;
    int* h = v.hdr + (size_t)g * H_WORDS;
    if (!game_runs<N>(v, h)) { if (lane == 0) h[H_NPATHS] = 0; return; }
    select_sequential<N>(v, g, lane);
}

// One CTA per game, one WARP PER DESCENT of the wave (R = parallel_readouts warps).  The descents of a wave only interact
// through the virtual loss (`ons`) earlier descents leave on the nodes they pass (self_play.py:644, :767-769): the choice of
// descent j at tree level L depends on the choices descents 0..j-1 made AT LEVEL L and nothing deeper.  So the warps run as a
// pipeline -- warp j handles level L as soon as warp j-1 has published its level-L choice in shared memory -- and the wave
// costs (depth + R - 1) node visits of latency instead of depth * R, with R times as many loads in flight.  Nothing is written
// to the tree until all descents are known.  Two cases need the reference order and are replayed by select_sequential:
// a descent that ends in a terminal leaf (its backup changes W / N before the next descent starts, :638-642), and more than one
// exact tie in the wave (the tie-break draws are numbered in descent-major order).  Both are rare outside the last moves.
// Launch bounds: a CTA lives for the whole wave of its game (latency bound), so ALL games must be resident at once or the
// kernel takes two rounds: 8 CTAs of 4 warps per SM (<= 64 registers) hold 1184 games of 9x9, 4 per SM 592 games of 19x19.
// FUSE: the leaves (step + legal mask + feature planes) follow in the same launch.  Off at 19x19, where the leaf code's register
// needs (ten 7-word bitboards) would spill inside the residency bound above.
__device__ int g_prefetch_guess = 0;      // tg_mcts_debug_prefetch: speculative prefetch of the most-visited child's edges (results unaffected)

// diagnostics (tg_mcts_debug_times): when set, every game's CTA of k_select_par records globaltimer at its start, after the
// selection pipeline and at its end, plus its deepest path, as 4 x u64 per game
__device__ unsigned long long* d_dbg_times = nullptr;
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int N, bool FUSE>
__global__ void __launch_bounds__(32 * kParWarps, N == 9 ? 8 : 4) k_select_par(View v, float* __restrict__ obs) {
    using GG = Geo<N>;
    constexpr int MAXR = ScanRegs<N>::MAXR;
    cudaGridDependencySynchronize();                                  // launched with programmatic stream serialization (launch_pdl)
    extern __shared__ int sm_sel[];                                   // [MAXD][R] node, then [MAXD][R] child index
    __shared__ int progress[kMaxReadouts], plen[kMaxReadouts];
    __shared__ int s_ties[kMaxReadouts], s_flag[kMaxReadouts], s_fan[kMaxReadouts];
    const int g = blockIdx.x, R = v.R;
    const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* h = v.hdr + (size_t)g * H_WORDS;
    unsigned long long* dbg_t = d_dbg_times;
    if (dbg_t && threadIdx.x == 0) dbg_t[4 * g] = gtimer();
    if (!game_runs<N>(v, h)) { if (threadIdx.x == 0) h[H_NPATHS] = 0; return; }   // CTA-uniform
    volatile int* sel_node = sm_sel;
    volatile int* sel_k = sm_sel + v.MAXD * R;
    volatile int* vprog = progress; volatile int* vlen = plen;
    if (threadIdx.x < kMaxReadouts) { progress[threadIdx.x] = 0; plen[threadIdx.x] = 1 << 30; }
    __syncthreads();
    const unsigned flags = (unsigned)h[H_FLAGS];
    const int buf = h[H_BUF];
    uint4* hot = hot_of<N>(v, buf, g);
    uint4* cold = cold_of<N>(v, buf, g);
    const uint2* kid = kid_of<N>(v, buf, g);
    const uint4* nodes = nodes_of<N>(v, buf, g);
    const uint4 root_info = nodes[GG::SQ];
    const uint4 root_meta = nodes[GG::SQ - 1];
    const bool root_after_pass = (int16_t)(root_meta.x & 0xffffu) == (int16_t)kPass;
    const int root_step = (int)(root_meta.z & 0xffffu);
    const double* p64 = v.root_prior64 + (size_t)g * GG::A;
    const bool noisy = (flags & TG_MCTS_NOISY) != 0;
    const int tie0 = h[H_TIE], rootN = h[H_ROOTN], rootons = h[H_ROOTONS];
    int* path = v.paths + ((size_t)g * R + j) * v.MAXD;
    // ---- descent j, speculative: registers, shared memory and this descent's private path row only
    int node = 0, depth = 0, Np = rootN, ons_base = rootons;
    uint32_t pk_sel = 0, kid_cur = pack_kid((int)root_info.x, (int)root_info.y);
    bool after_pass = root_after_pass;
    int ties = 0, fan = 0;
    int flag = 0;                                                     // 1: needs the sequential replay, 2: depth error, 4: NaN
    for (int L = 0;; ++L) {
        const int eb = (int)(kid_cur & 0x3fffffu), K = (int)(kid_cur >> 22);
        fan += K;
        ScanRegs<N> sr;
        scan_load<N>(sr, hot, kid, p64, noisy && node == 0, eb, K, lane);    // in flight while waiting for warp j-1
        // every earlier descent must have published its level-L choice (or have ended): a short descent j-1 says nothing
        // about how far a long descent j-2 has come
        for (int jp = 0; jp < j; ++jp) { while (vprog[jp] <= L) __nanosleep(40); }
        __syncwarp();
        // speculation: the most-visited child is the likely pick.  Pull ITS edge records towards L2 while the scores are worked
        // out (a pick is ~1 us, the round trip it hides ~0.5 us); a wrong guess costs a few idle prefetches, never a result.
        if (g_prefetch_guess) {
            unsigned key = 0;
#pragma unroll
            for (int r = 0; r < MAXR; ++r) {
                const int k = lane + 32 * r;
                const unsigned kk = k < K && sr.ch[r] != kNoChild ? (((unsigned)pk_n(sr.q[r].w) << 12) | ((unsigned)r << 5) | (unsigned)lane) + 1u : 0u;
                key = kk > key ? kk : key;
            }
            key = __reduce_max_sync(0xffffffffu, key);
            if (key) {
                const int rs = (int)(((key - 1u) >> 5) & 127u), ls = (int)((key - 1u) & 31u);
                uint32_t kdg = 0;
#pragma unroll
                for (int r = 0; r < MAXR; ++r) { const uint32_t t = __shfl_sync(0xffffffffu, sr.kd[r], ls); kdg = r == rs ? t : kdg; }
                const int ebg = (int)(kdg & 0x3fffffu), Kg = (int)(kdg >> 22);
                const char* ph = reinterpret_cast<const char*>(hot + ebg);
                const char* pk = reinterpret_cast<const char*>(kid + ebg);
                if (lane * 128 < Kg * 16) asm volatile("prefetch.global.L2 [%0];" :: "l"(ph + lane * 128));
                if (lane * 128 < Kg * 8) asm volatile("prefetch.global.L2 [%0];" :: "l"(pk + lane * 128));
            }
        }
        // virtual loss of the earlier descents that stand on this node at this level
        int same = 0, xo[MAXR];
#pragma unroll
        for (int r = 0; r < MAXR; ++r) xo[r] = 0;
        for (int jp = 0; jp < j; ++jp) {
            if (L < vlen[jp] && sel_node[L * R + jp] == node) {
                same += 1;
                const int kx = sel_k[L * R + jp];
#pragma unroll
                for (int r = 0; r < MAXR; ++r) xo[r] += (kx == lane + 32 * r) ? 1 : 0;
            }
        }
        int ksel; uint32_t ch_sel, kd_sel;
        // exact ties are numbered in the reference's order (all of descent 0's, then descent 1's, ...).  Ties are rare: pick
        // assuming the earlier descents have none; on a tie wait until they have finished, and pick again if they had any.
        const int ties_before = ties;
        if (!scan_pick<N>(v, sr, noisy && node == 0, K, Np, ons_base + same * v.wu, xo, tie0 + ties_before, ties, g, lane, ksel, pk_sel, ch_sel, kd_sel)) { flag |= 4; break; }
        if (ties != ties_before && j > 0) {
            int earlier = 0;
            for (int jp = 0; jp < j; ++jp) { while (vprog[jp] < (1 << 30)) __nanosleep(40); earlier += s_ties[jp]; }
            if (earlier) {
                ties = ties_before;
                scan_pick<N>(v, sr, noisy && node == 0, K, Np, ons_base + same * v.wu, xo, tie0 + earlier + ties_before, ties, g, lane, ksel, pk_sel, ch_sel, kd_sel);
            }
        }
        if (depth >= v.MAXD) { flag |= 2; break; }
        if (lane == 0) { path[depth] = eb + ksel; sel_node[L * R + j] = node; sel_k[L * R + j] = ksel; }
        depth += 1;
        __syncwarp();
        __threadfence_block();
        if (lane == 0) vprog[j] = L + 1;
        const int child = (int)ch_sel;
        if (child < 0) break;
        after_pass = pk_act(pk_sel) == GG::P;
        Np = pk_n(pk_sel); ons_base = pk_ons(pk_sel); node = child; kid_cur = kd_sel;
    }
    const int action = pk_act(pk_sel);
    const int step2 = root_step + depth;
    if (!flag && ((action == GG::P && after_pass && step2 > 1) || step2 > v.max_step)) flag |= 1;   // terminal leaf
    if (lane == 0) {
        vlen[j] = depth; s_ties[j] = ties; s_flag[j] = flag; s_fan[j] = fan;
        __threadfence_block();
        vprog[j] = 1 << 30;
    }
    __syncthreads();
    if (dbg_t && threadIdx.x == 0) { int md = 0; for (int jp = 0; jp < R; ++jp) md = plen[jp] > md ? plen[jp] : md; dbg_t[4 * g + 1] = gtimer(); dbg_t[4 * g + 3] = (unsigned long long)md; }
    int any_flag = 0, all_ties = 0;
    for (int jp = 0; jp < R; ++jp) { any_flag |= s_flag[jp]; all_ties += s_ties[jp]; }
    uint32_t* sm_planes = reinterpret_cast<uint32_t*>(sm_sel) + j * leaf_smem_words<N>();   // the pipeline state is dead by now
    if (any_flag) {                                                   // CTA-uniform: replay in the reference order
        if (threadIdx.x == 0) h[H_SPARE1] += 1;                       // replay counter (diagnostics)
        if (j == 0) select_sequential_cold<N>(v, g, lane);
        __syncthreads();
        if (FUSE && j < h[H_NPATHS]) {
            const size_t slot = (size_t)g * R + j;
            leaf_work<N>(v, nodes, slot, v.leaf_parent[slot], v.leaf_action[slot], obs, sm_planes, lane);
        }
        return;
    }
    // ---- commit: every descent found a leaf to evaluate
    __syncwarp();
    for (int i = lane; i <= depth; i += 32) {                         // incomplete_update (:767-769): +wu on every node of the path
        if (i == depth) atomicAdd(&h[H_ROOTONS], v.wu);
        else atomicAdd(&reinterpret_cast<uint32_t*>(hot + path[i])[3], (uint32_t)(v.wu * 512));
    }
    const size_t slot = (size_t)g * R + j;
    if (lane == 0) { v.leaf_parent[slot] = node; v.leaf_action[slot] = action; v.path_len[slot] = depth; }
    if (threadIdx.x == 0) {
        int sd = 0, sf = 0;
        for (int jp = 0; jp < R; ++jp) { sd += plen[jp]; sf += s_fan[jp]; }
        h[H_TIE] = tie0 + all_ties; h[H_NPATHS] = R;
        h[H_STAT_DESC] += R; h[H_STAT_DEPTH] += sd; h[H_STAT_FAN] += sf;
    }
    // ---- the leaf of this descent: step + legal mask + feature planes, straight into the evaluator's input (k_leaves fused:
    // a game goes on to its leaves as soon as ITS descents are done instead of waiting for the slowest game of the batch)
    if (FUSE) leaf_work<N>(v, nodes, slot, node, action, obs, sm_planes, lane);
    if (dbg_t) { __syncthreads(); if (threadIdx.x == 0) dbg_t[4 * g + 2] = gtimer(); }
}

// ---------------------------------------------------------------------------------------------- leaves
// One warp per leaf slot: env.step from the parent (:629), getLegalAction (:634, :737) and encode (:798) fused; the planes
// go straight into the net's input tensor.
template <int N>
__global__ void __launch_bounds__(kThreads) k_leaves(View v, float* __restrict__ obs) {
    using GG = Geo<N>;
    extern __shared__ uint32_t sm_bits[];
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (slot >= v.G * v.R) return;
    const int g = slot / v.R, jj = slot - g * v.R;
    const int* h = v.hdr + (size_t)g * H_WORDS;
    if (jj >= h[H_NPATHS]) return;
    leaf_work<N>(v, nodes_of<N>(v, h[H_BUF], g), (size_t)slot, v.leaf_parent[slot], v.leaf_action[slot], obs, sm_bits + wib * leaf_smem_words<N>(), lane);
}

// write one node's edges from a dense legal mask: prior_k = pol[a_k] * (renorm ? 1/scale : raw), W = w0
// returns K; sp/sa are per-warp shared scratch of A entries.
template <int N>
__device__ __forceinline__ int compact_legal(const uint32_t* __restrict__ legal, const float* __restrict__ pol, float* sp, int* sa) {
    using GG = Geo<N>;
    constexpr int MAXR = (GG::A + 31) / 32;
    const int lane = threadIdx.x & 31;
    int K = 0;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        const int a = lane + 32 * r;
        const uint32_t word = legal[r];                           // word r holds actions 32r .. 32r+31
        const bool on = a < GG::A && ((word >> lane) & 1u);
        const int pos = K + __popc(word & ((1u << lane) - 1u));
        if (on) { sp[pos] = pol[a]; sa[pos] = a; }
        K += __popc(word);
    }
    __syncwarp();
    return K;
}

// ---------------------------------------------------------------------------------------------- expansion + backup
template <int N>
__global__ void __launch_bounds__(kThreads) k_expand(View v, const float* __restrict__ policy, const float* __restrict__ value, int* __restrict__ pending) {
    using GG = Geo<N>;
    extern __shared__ uint32_t sm_raw[];
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (g >= v.G) return;
    float* sp = reinterpret_cast<float*>(sm_raw) + wib * 2 * GG::A;
    int* sa = reinterpret_cast<int*>(sp + GG::A);
    int* h = v.hdr + (size_t)g * H_WORDS;
    unsigned flags = (unsigned)h[H_FLAGS];
    const int npaths = h[H_NPATHS];
    const int buf = h[H_BUF];
    uint4* hot = hot_of<N>(v, buf, g);
    uint4* cold = cold_of<N>(v, buf, g);
    uint2* kid = kid_of<N>(v, buf, g);
    uint4* nodes = nodes_of<N>(v, buf, g);
    unsigned err = 0;
    int sims = 0;
    int nn = h[H_NNODES], ne = h[H_NEDGES];
    for (int j = 0; j < npaths; ++j) {                                                // self_play.py:651-654
        const size_t slot = (size_t)g * v.R + j;
        const int* path = v.paths + slot * v.MAXD;
        const int len = v.path_len[slot];
        const int e = path[len - 1];
        const float val = value[slot];                                                // (issued with the loads below, used later)
        if (kid[e].x != kNoChild || err) { add_ons(h, hot, path, len, -v.wu); continue; }   // :653, :733-734 duplicate leaf: no backup
        const int K = compact_legal<N>(v.leaf_legal + slot * GG::DQ * 4, policy + slot * GG::A, sp, sa);
        const int node = nn, eb = ne;
        if (node + 1 > v.NC) err |= TG_MCTS_ERR_NODES;
        if (eb + K > v.EC) err |= TG_MCTS_ERR_EDGES;
        if (err) { add_ons(h, hot, path, len, -v.wu); continue; }
        float scale = 0.0f;                                                           // :739 python sum(): sequential float32
        if (lane == 0) for (int k = 0; k < K; ++k) scale = __fadd_rn(scale, sp[k]);
        scale = __shfl_sync(0xffffffffu, scale, 0);
        // a NaN / inf from the evaluator would poison W along the whole path and every later selection: fail loudly here
        if (!isfinite(scale) || !isfinite(val)) { err |= TG_MCTS_ERR_NAN; add_ons(h, hot, path, len, -v.wu); continue; }
        const bool ok = scale > 0.0f;                                                 // :740
        for (int k = lane; k < K; k += 32) {
            const float prior = ok ? __fdiv_rn(sp[k], scale) : 0.0f;                   // :743-744
            const float w0 = ok ? -val : -0.0f;                                        // :752 / :76
            write_fresh_edge(hot, cold, kid, eb + k, (float)v.c1, prior, w0, sa[k]);
        }
        uint4* rec = nodes + (size_t)node * (GG::SQ + 1);
        if (lane < GG::SQ) rec[lane] = v.leaf_state[(size_t)lane * v.G * v.R + slot];
        if (lane == 0) {
            rec[GG::SQ] = make_uint4((uint32_t)eb, (uint32_t)K, 0u, 0u);
            kid[e] = make_uint2((uint32_t)node, pack_kid(eb, K));
        }
        nn = node + 1; ne = eb + K;
        __syncwarp();
        backup<N>(v, h, hot, cold, path, len, val, err, -v.wu);                        // :653 + :755 in one pass over the path
        sims += 1;
        __syncwarp();
    }
    if (lane == 0) {
        h[H_NPATHS] = 0; h[H_SIMS] += sims; h[H_NNODES] = nn; h[H_NEDGES] = ne;
        if (err) { flags |= err; h[H_FLAGS] = (int)flags; }
        if (pending && (flags & TG_MCTS_ACTIVE) && !(flags & (TG_MCTS_ERR_ANY | kErrAction)) && h[H_ROOTN] < h[H_N0] + v.num_sim)
            atomicAdd(pending, 1);
    }
}

// One CTA per game, one warp per leaf of the wave (the expansion counterpart of k_select_par).  complete_update
// (self_play.py:727-755) of the R leaves of a wave is independent work -- legal list, prior renormalisation, R x K new edge
// records, the node record -- except for three things that follow the reference's leaf order j = 0..R-1 and are resolved
// between the phases: duplicate leaves (:733-734), arena allocation, and the value backups, whose float32 sums / Welford
// updates on shared ancestors must be applied in leaf order.
//   phase A  warp j: loads of leaf j, legal list compaction, sequential float32 prior sum (:739)
//   phase B  thread 0: the sequential decisions (duplicate / error / node and edge base) for j = 0..R-1
//   phase C  warp j: writes the node and its edges
//   phase D  warp 0: all backups and virtual-loss reverts, lane = tree level; each lane loads the <= R records of its level
//            once, applies the leaves' updates in order with forwarding between leaves that share the edge, stores once
template <int N>
__global__ void __launch_bounds__(32 * kParWarps) k_expand_par(View v, const float* __restrict__ policy, const float* __restrict__ value, int* __restrict__ pending) {
    using GG = Geo<N>;
    cudaGridDependencySynchronize();                                  // launched with programmatic stream serialization (launch_pdl)
    extern __shared__ uint32_t sm_raw[];
    __shared__ int s_e[kParWarps], s_len[kParWarps], s_K[kParWarps], s_pre[kParWarps], s_bad[kParWarps], s_exp[kParWarps], s_node[kParWarps], s_eb[kParWarps];
    __shared__ float s_val[kParWarps], s_scale[kParWarps];
    __shared__ unsigned s_err;
    __shared__ int s_nn, s_ne;
    const int g = blockIdx.x, R = v.R;
    const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sp = reinterpret_cast<float*>(sm_raw) + j * 2 * GG::A;
    int* sa = reinterpret_cast<int*>(sp + GG::A);
    int* h = v.hdr + (size_t)g * H_WORDS;
    unsigned flags = (unsigned)h[H_FLAGS];
    const int npaths = h[H_NPATHS];
    const int buf = h[H_BUF];
    uint4* hot = hot_of<N>(v, buf, g);
    uint4* cold = cold_of<N>(v, buf, g);
    uint2* kid = kid_of<N>(v, buf, g);
    uint4* nodes = nodes_of<N>(v, buf, g);
    const size_t slot = (size_t)g * R + j;
    const int* path = v.paths + slot * v.MAXD;
    // ---- phase A
    if (j < npaths) {
        const int len = v.path_len[slot];
        const int e = path[len - 1];
        const float val = value[slot];
        const int pre = kid[e].x != kNoChild;
        const int K = compact_legal<N>(v.leaf_legal + slot * GG::DQ * 4, policy + slot * GG::A, sp, sa);
        float scale = 0.0f;                                                           // :739 python sum(): sequential float32
        if (lane == 0) for (int k = 0; k < K; ++k) scale = __fadd_rn(scale, sp[k]);
        if (lane == 0) {
            s_e[j] = e; s_len[j] = len; s_K[j] = K; s_pre[j] = pre; s_val[j] = val; s_scale[j] = scale;
            s_bad[j] = (!isfinite(scale) || !isfinite(val)) ? 1 : 0;
        }
    }
    __syncthreads();
    // ---- phase B: the reference's leaf order
    if (threadIdx.x == 0) {
        unsigned err = 0;
        int nn = h[H_NNODES], ne = h[H_NEDGES];
        for (int jj = 0; jj < npaths; ++jj) {
            s_exp[jj] = 0;
            bool dup = s_pre[jj] != 0;
            for (int jp = 0; jp < jj; ++jp) dup = dup || (s_exp[jp] && s_e[jp] == s_e[jj]);
            if (dup || err) continue;                                                 // :653, :733-734 duplicate leaf: no backup
            if (nn + 1 > v.NC) err |= TG_MCTS_ERR_NODES;
            if (ne + s_K[jj] > v.EC) err |= TG_MCTS_ERR_EDGES;
            if (err) continue;
            if (s_bad[jj]) { err |= TG_MCTS_ERR_NAN; continue; }                      // NaN / inf from the evaluator: fail loudly
            s_exp[jj] = 1; s_node[jj] = nn; s_eb[jj] = ne;
            nn += 1; ne += s_K[jj];
        }
        s_err = err; s_nn = nn; s_ne = ne;
    }
    __syncthreads();
    // ---- phase C
    if (j < npaths && s_exp[j]) {
        const int K = s_K[j], eb = s_eb[j], node = s_node[j], e = s_e[j];
        const float scale = s_scale[j], val = s_val[j];
        const bool ok = scale > 0.0f;                                                 // :740
        for (int k = lane; k < K; k += 32) {
            const float prior = ok ? __fdiv_rn(sp[k], scale) : 0.0f;                   // :743-744
            const float w0 = ok ? -val : -0.0f;                                        // :752 / :76
            write_fresh_edge(hot, cold, kid, eb + k, (float)v.c1, prior, w0, sa[k]);
        }
        uint4* rec = nodes + (size_t)node * (GG::SQ + 1);
        if (lane < GG::SQ) rec[lane] = v.leaf_state[(size_t)lane * v.G * R + slot];
        if (lane == 0) {
            rec[GG::SQ] = make_uint4((uint32_t)eb, (uint32_t)K, 0u, 0u);
            kid[e] = make_uint2((uint32_t)node, pack_kid(eb, K));
        }
    }
    __syncthreads();
    // ---- phase D: revert_incomplete_update (:770-774) of every leaf + backpropagate (:758-764) of the expanded ones, leaf order
    if (j == 0) {
        unsigned err = s_err;
        int maxlen = 0;
        for (int jj = 0; jj < npaths; ++jj) maxlen = s_len[jj] > maxlen ? s_len[jj] : maxlen;
        for (int i = lane; i < maxlen; i += 32) {
            int e[kParWarps]; uint4 q[kParWarps]; uint4 c[kParWarps];
#pragma unroll
            for (int jj = 0; jj < kParWarps; ++jj) {
                e[jj] = (jj < npaths && i < s_len[jj]) ? v.paths[((size_t)g * R + jj) * v.MAXD + i] : -1;
                if (e[jj] >= 0) { q[jj] = hot[e[jj]]; c[jj] = cold[e[jj]]; }
            }
#pragma unroll
            for (int jj = 0; jj < kParWarps; ++jj) {
                if (e[jj] < 0) continue;
#pragma unroll
                for (int jp = 0; jp < jj; ++jp) if (e[jp] == e[jj]) { q[jj] = q[jp]; c[jj] = c[jp]; }    // the latest earlier leaf on this edge
                q[jj].w = (uint32_t)((int)q[jj].w - v.wu * 512);
                if (s_exp[jj]) {
                    const float val = s_val[jj];
                    backup_edge(q[jj], c[jj], ((s_len[jj] - 1 - i) & 1) ? -val : val, (float)v.c2, err);
                }
            }
#pragma unroll
            for (int jj = 0; jj < kParWarps; ++jj) {
                if (e[jj] < 0) continue;
                bool last = true;
#pragma unroll
                for (int jq = jj + 1; jq < kParWarps; ++jq) last = last && e[jq] != e[jj];
                if (last) { hot[e[jj]] = q[jj]; cold[e[jj]] = c[jj]; }
            }
        }
        err |= __reduce_or_sync(0xffffffffu, err);
        if (lane == 0) {                                                               // the root's own statistics, leaf order
            float W = __int_as_float(h[H_ROOTW]), mean = __int_as_float(h[H_ROOTMEAN]), var = __int_as_float(h[H_ROOTVAR]);
            int n = h[H_ROOTN], ons = h[H_ROOTONS], sims = 0;
            for (int jj = 0; jj < npaths; ++jj) {
                ons -= v.wu;
                if (!s_exp[jj]) continue;
                const float val = s_val[jj];
                const float rv = (s_len[jj] & 1) ? -val : val;
                W = __fadd_rn(W, rv); n += 1;
                const float t = mean;
                mean = __fdiv_rn(W, (float)(n + 1));
                var = __fadd_rn(var, __fmul_rn(__fsub_rn(rv, t), __fsub_rn(rv, mean)));
                sims += 1;
            }
            h[H_ROOTW] = __float_as_int(W); h[H_ROOTN] = n; h[H_ROOTMEAN] = __float_as_int(mean); h[H_ROOTVAR] = __float_as_int(var);
            h[H_ROOTONS] = ons;
            h[H_NPATHS] = 0; h[H_SIMS] += sims; h[H_NNODES] = s_nn; h[H_NEDGES] = s_ne;
            if (err) { flags |= err; h[H_FLAGS] = (int)flags; }
            if (pending && (flags & TG_MCTS_ACTIVE) && !(flags & (TG_MCTS_ERR_ANY | kErrAction)) && n < h[H_N0] + v.num_sim)
                atomicAdd(pending, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------- roots
template <int N>
__global__ void k_init(View v, const uint8_t* __restrict__ mask) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= v.G) return;
    if (mask && !mask[g]) return;
    int* h = v.hdr + (size_t)g * H_WORDS;
    const int sims = mask ? h[H_SIMS] : 0;
    for (int i = 0; i < H_WORDS; ++i) h[i] = 0;
    h[H_SIMS] = sims;
    h[H_FLAGS] = (int)(TG_MCTS_ACTIVE | TG_MCTS_NEED_EXPAND);
    Pos<N> s; pos_reset(s);
    pos_store<N>(s, v.root_state, g, v.G, 1);
}

template <int N>
__global__ void __launch_bounds__(kThreads) k_root_expand(View v, const float* __restrict__ policy_root, const float* __restrict__ value_root) {
    using GG = Geo<N>;
    extern __shared__ uint32_t sm_raw[];
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (g >= v.G) return;
    int* h = v.hdr + (size_t)g * H_WORDS;
    unsigned flags = (unsigned)h[H_FLAGS];
    if (!(flags & TG_MCTS_NEED_EXPAND) || !(flags & TG_MCTS_ACTIVE)) return;
    float* sp = reinterpret_cast<float*>(sm_raw) + wib * 2 * GG::A;
    int* sa = reinterpret_cast<int*>(sp + GG::A);
    __shared__ uint32_t dense[kThreads / 32][GG::DQ * 4];
    const int buf = h[H_BUF];
    uint4* hot = hot_of<N>(v, buf, g);
    uint4* cold = cold_of<N>(v, buf, g);
    uint2* kid = kid_of<N>(v, buf, g);
    uint4* nodes = nodes_of<N>(v, buf, g);
    Pos<N> s; pos_load<N>(s, v.root_state, g, v.G, 1);
    Groups<N> gr;
    analyse_groups<N, 32, false>(s.stones[0], s.stones[1], s.player, gr);
    const BB<N> lg = legal_points(s, gr);
    if (lane == 0) {
        uint32_t d[GG::DQ * 4];
#pragma unroll
        for (int i = 0; i < GG::DQ * 4; ++i) d[i] = 0;
        bb_to_dense<N>(lg, d);
        if (!any(lg)) d[GG::P >> 5] |= 1u << (GG::P & 31);
#pragma unroll
        for (int i = 0; i < GG::DQ * 4; ++i) dense[wib][i] = d[i];
    }
    __syncwarp();
    const int K = compact_legal<N>(dense[wib], policy_root + (size_t)g * GG::A, sp, sa);
    unsigned err = 0;
    if (K > v.EC) err |= TG_MCTS_ERR_EDGES;
    if (v.NC < 1) err |= TG_MCTS_ERR_NODES;
    bool bad = !isfinite(value_root[g]);
    for (int k = lane; k < K; k += 32) bad = bad || !isfinite(sp[k]);
    if (__any_sync(0xffffffffu, bad)) err |= TG_MCTS_ERR_NAN;                          // evaluator returned NaN / inf
    if (!err) {
        const float w0 = -value_root[g];                                                // :76, value from :599
        for (int k = lane; k < K; k += 32) write_fresh_edge(hot, cold, kid, k, (float)v.c1, sp[k], w0, sa[k]);   // RAW prior (:602)
        uint4* rec = nodes;
        if (lane < GG::SQ) rec[lane] = v.root_state[(size_t)lane * v.G + g];
        if (lane == 0) rec[GG::SQ] = make_uint4(0u, (uint32_t)K, 0u, 0u);
    }
    if (lane == 0) {
        h[H_NNODES] = 1; h[H_NEDGES] = K;
        flags &= ~(TG_MCTS_NEED_EXPAND | TG_MCTS_NOISY);
        h[H_FLAGS] = (int)(flags | err);
    }
}

template <int N>
__global__ void k_begin_move(View v, const double* __restrict__ noise) {
    using GG = Geo<N>;
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= v.G) return;
    int* h = v.hdr + (size_t)g * H_WORDS;
    unsigned flags = (unsigned)h[H_FLAGS];
    if (!(flags & TG_MCTS_ACTIVE) || (flags & TG_MCTS_NEED_EXPAND)) return;
    const int K = h[H_NEDGES] > 0 ? (int)nodes_of<N>(v, h[H_BUF], g)[GG::SQ].y : 0;
    if (noise) {                                                                     // Node_V.dirichlet_prior :90-95
        const uint4* cold = cold_of<N>(v, h[H_BUF], g);
        for (int k = lane; k < K; k += 32) {
            const float p = __fmul_rn(__uint_as_float(cold[k].z), 0.75f);
            v.root_prior64[(size_t)g * GG::A + k] = __dadd_rn((double)p, __dmul_rn(noise[(size_t)g * GG::A + k], 0.25));
        }
        flags |= TG_MCTS_NOISY;
    } else flags &= ~TG_MCTS_NOISY;
    if (lane == 0) { h[H_N0] = h[H_ROOTN]; h[H_FLAGS] = (int)flags; }
}

template <int N>
__global__ void k_root_stats(View v, int* __restrict__ counts, float* __restrict__ W, float* __restrict__ var, float* __restrict__ prior,
                             int* __restrict__ root_n, int* __restrict__ root_k, int* __restrict__ step) {
    using GG = Geo<N>;
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= v.G) return;
    const int* h = v.hdr + (size_t)g * H_WORDS;
    const size_t row = (size_t)g * GG::A;
    for (int a = lane; a < GG::A; a += 32) {
        if (counts) counts[row + a] = 0;
        if (W) W[row + a] = 0.0f;
        if (var) var[row + a] = 0.0f;
        if (prior) prior[row + a] = 0.0f;
    }
    __syncwarp();
    const bool has = !((unsigned)h[H_FLAGS] & TG_MCTS_NEED_EXPAND) && h[H_NNODES] > 0;
    const int K = has ? (int)nodes_of<N>(v, h[H_BUF], g)[GG::SQ].y : 0;
    const uint4* hot = hot_of<N>(v, h[H_BUF], g);
    const uint4* cold = cold_of<N>(v, h[H_BUF], g);
    for (int k = lane; k < K; k += 32) {
        const uint32_t pk = hot[k].w;
        const int a = pk_act(pk);
        if (counts) counts[row + a] = pk_n(pk);
        if (W || var || prior) {
            const uint4 c = cold[k];
            if (W) W[row + a] = __uint_as_float(c.x);
            if (var) var[row + a] = __uint_as_float(c.y);
            if (prior) prior[row + a] = __uint_as_float(c.z);
        }
    }
    if (lane == 0) {
        if (root_n) root_n[g] = h[H_ROOTN];
        if (root_k) root_k[g] = K;
        if (step) step[g] = (int)(v.root_state[(size_t)(GG::SQ - 1) * v.G + g].z & 0xffffu);
    }
}

// update_with_action (:857-872): env.step on the root state, then keep the chosen child's subtree by copying it
// breadth-first into the other half of the arena.
template <int N>
__global__ void __launch_bounds__(kThreads) k_advance(View v, const int16_t* __restrict__ actions, uint8_t* __restrict__ done_out) {
    using GG = Geo<N>;
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= v.G) return;
    int* h = v.hdr + (size_t)g * H_WORDS;
    unsigned flags = (unsigned)h[H_FLAGS];
    if (!(flags & TG_MCTS_ACTIVE) || (flags & (TG_MCTS_NEED_EXPAND | TG_MCTS_ERR_ANY | kErrAction))) {
        if (lane == 0 && done_out) done_out[g] = (flags & TG_MCTS_ACTIVE) ? 0 : 1;
        return;
    }
    const int action = (int)actions[g];
    const int buf = h[H_BUF];
    const uint4* shot = hot_of<N>(v, buf, g);
    const uint4* scold = cold_of<N>(v, buf, g);
    const uint2* skid = kid_of<N>(v, buf, g);
    const uint4* snodes = nodes_of<N>(v, buf, g);
    const int K0 = (int)snodes[GG::SQ].y;
    // the root edge of `action`
    int ksel = -1;
    for (int base = 0; base < K0; base += 32) {
        const int k = base + lane;
        const bool hit = k < K0 && pk_act(shot[k].w) == action;
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (b && ksel < 0) ksel = base + __ffs((int)b) - 1;
    }
    if (ksel < 0) { if (lane == 0) { h[H_FLAGS] = (int)(flags | kErrAction); if (done_out) done_out[g] = 0; } return; }
    // env.step(root.state, action)  (:859)
    Pos<N> s; pos_load<N>(s, v.root_state, g, v.G, 1);
    Groups<N> gr;
    analyse_groups<N, 32, false>(s.stones[0], s.stones[1], s.player, gr);
    bool acc;
    const bool done = apply_move<N, 32>(s, gr, action, v.max_step, acc, true);
    if (lane == 0) pos_store<N>(s, v.root_state, g, v.G, 1);
    const uint4 qe = shot[ksel];
    const uint4 ce = scold[ksel];
    const int child = (int)skid[ksel].x;
    int nd = 0, ne = 0;
    if (child >= 0) {
        const int dbuf = 1 - buf;
        uint4* dhot = hot_of<N>(v, dbuf, g);
        uint4* dcold = cold_of<N>(v, dbuf, g);
        uint2* dkid = kid_of<N>(v, dbuf, g);
        uint4* dnodes = nodes_of<N>(v, dbuf, g);
        if (lane == 0) { dnodes[GG::SQ].z = (uint32_t)child; dnodes[GG::SQ].w = 0xffffffffu; }   // .w: the copied edge that leads here
        __syncwarp();
        nd = 1;
        for (int i = 0; i < nd; ++i) {
            uint4* drec = dnodes + (size_t)i * (GG::SQ + 1);
            const int src = (int)drec[GG::SQ].z;
            const uint4* srec = snodes + (size_t)src * (GG::SQ + 1);
            const uint4 sinfo = srec[GG::SQ];
            const int seb = (int)sinfo.x, K = (int)sinfo.y;
            const uint32_t pe = drec[GG::SQ].w;                                        // set when this node was enqueued
            if (lane == 0 && pe != 0xffffffffu) dkid[pe].y = pack_kid(ne, K);           // its edges now start at ne
            if (lane < GG::SQ) drec[lane] = srec[lane];
            for (int base = 0; base < K; base += 32) {
                const int k = base + lane;
                uint4 q = make_uint4(0, 0, 0, 0), c = make_uint4(0, 0, 0, 0); uint2 kq = make_uint2(kNoChild, 0u);
                if (k < K) { q = shot[seb + k]; c = scold[seb + k]; kq = skid[seb + k]; }
                const bool has = k < K && kq.x != kNoChild;
                const unsigned b = __ballot_sync(0xffffffffu, has);
                if (has) {
                    const int my = nd + __popc(b & ((1u << lane) - 1u));
                    dnodes[(size_t)my * (GG::SQ + 1) + GG::SQ].z = kq.x;
                    dnodes[(size_t)my * (GG::SQ + 1) + GG::SQ].w = (uint32_t)(ne + k);
                    kq.x = (uint32_t)my;                                               // .y is patched when the child is dequeued
                }
                if (k < K) { dhot[ne + k] = q; dcold[ne + k] = c; dkid[ne + k] = kq; }
                nd += __popc(b);
            }
            __syncwarp();
            if (lane == 0) drec[GG::SQ] = make_uint4((uint32_t)ne, (uint32_t)K, (uint32_t)src, 0u);
            ne += K;
            __syncwarp();
        }
    }
    if (lane == 0) {
        h[H_BUF] = child >= 0 ? 1 - buf : buf;
        h[H_NNODES] = nd; h[H_NEDGES] = ne;
        h[H_ROOTN] = pk_n(qe.w); h[H_ROOTONS] = pk_ons(qe.w);
        // the new root's own statistics = those of the edge that led to it (value_mean: 0 before its first visit, else W / (N + 1))
        h[H_ROOTW] = (int)ce.x; h[H_ROOTVAR] = (int)ce.y;
        h[H_ROOTMEAN] = pk_n(qe.w) == 0 ? 0 : __float_as_int(-__uint_as_float(qe.y));
        h[H_MOVE] += 1;
        flags &= ~TG_MCTS_NOISY;
        if (child < 0) flags |= TG_MCTS_NEED_EXPAND;                                   // :861-870
        if (done) flags &= ~TG_MCTS_ACTIVE;
        h[H_FLAGS] = (int)flags;
        if (done_out) done_out[g] = done ? 1 : 0;
    }
}

// selection statistics since the previous call, summed over the games: descents, edges walked, children scanned
__global__ void k_search_stats(View v, unsigned long long* __restrict__ out3) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= v.G) return;
    int* h = v.hdr + (size_t)g * H_WORDS;
    if (h[H_STAT_DESC]) {
        atomicAdd(&out3[0], (unsigned long long)h[H_STAT_DESC]);
        atomicAdd(&out3[1], (unsigned long long)h[H_STAT_DEPTH]);
        atomicAdd(&out3[2], (unsigned long long)h[H_STAT_FAN]);
    }
    h[H_STAT_DESC] = 0; h[H_STAT_DEPTH] = 0; h[H_STAT_FAN] = 0;
}

__global__ void k_status(View v, uint32_t* __restrict__ status, unsigned long long* __restrict__ sims) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0;
    if (g < v.G) {
        int* h = v.hdr + (size_t)g * H_WORDS;
        if (status) status[g] = (uint32_t)h[H_FLAGS];
        if (sims) { mine = (unsigned long long)(unsigned)h[H_SIMS]; h[H_SIMS] = 0; }
    }
    if (sims) {
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, s);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(sims, mine);
    }
}

__global__ void k_winner(const float* __restrict__ score, uint8_t* __restrict__ winner, int G) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) winner[g] = score[g] > 0.0f ? kBlack : kWhite;                          // environment.py:118-119
}

__global__ void k_bump(unsigned long long* c) { *c += 1ull; }

// number of games that still need simulations for the current move (the loop condition of self_play.py:660-664); one block
template <int N>
__global__ void k_pending(View v, int* __restrict__ out) {
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    int mine = 0;
    for (int g = threadIdx.x; g < v.G; g += blockDim.x) {
        const int* h = v.hdr + (size_t)g * H_WORDS;
        const unsigned flags = (unsigned)h[H_FLAGS];
        if ((flags & TG_MCTS_ACTIVE) && !(flags & (TG_MCTS_NEED_EXPAND | TG_MCTS_ERR_ANY | kErrAction)) && h[H_ROOTN] < h[H_N0] + v.num_sim) mine += 1;
    }
    if (mine) atomicAdd(&total, mine);
    __syncthreads();
    if (threadIdx.x == 0) *out = total;
}

// synthetic provider (bench only): heavy-tailed integer weights -> one f32 division per entry
__global__ void k_synth(int A, long long rows, unsigned long long seed, unsigned long long wave, const unsigned long long* __restrict__ wave_dev,
                        float* __restrict__ policy, float* __restrict__ value) {
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    if (wave_dev) wave += *wave_dev;            // the wave index lives on the device when the wave is replayed from a CUDA graph
    // each lane owns actions lane, lane + 32, ...: hash once, keep the weights in registers (A <= 12 * 32)
    unsigned long long sum = 0;
    unsigned wgt[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int a = lane + 32 * i;
        wgt[i] = 0;
        if (a < A) {
            const unsigned long long r = mix32(seed + wave * 0x632BE59BD9B4E019ull, (unsigned long long)row, (unsigned long long)a) & 0xFFFFu;
            wgt[i] = (unsigned)(1ull + ((r * r * r) >> 32));
            sum += wgt[i];
        }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    const float total = (float)sum;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int a = lane + 32 * i;
        if (a < A) policy[row * A + a] = __fdiv_rn((float)wgt[i], total);
    }
    if (lane == 0) {
        const int vv = (int)(mix32(seed ^ 0x5555ull, (unsigned long long)row, wave) & 0xFFFFu) - 32768;
        value[row] = __fdiv_rn((float)vv, 32832.0f);
    }
}

// device-side wave counter: every block reads it when it starts; the block that finishes LAST (ticket) advances it, so the
// increment needs no launch of its own.  ticket lives next to the counter (wave_dev[1]).
__global__ void k_synth_counted(int A, long long rows, unsigned long long seed, unsigned long long* __restrict__ wave_dev,
                                float* __restrict__ policy, float* __restrict__ value) {
    __shared__ unsigned long long s_wave;
    cudaGridDependencySynchronize();                                  // launched with programmatic stream serialization (launch_pdl)
    if (threadIdx.x == 0) s_wave = *reinterpret_cast<volatile unsigned long long*>(wave_dev);
    __syncthreads();
    const unsigned long long wave = s_wave;
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row < rows) {
        unsigned long long sum = 0;
        unsigned wgt[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int a = lane + 32 * i;
            wgt[i] = 0;
            if (a < A) {
                const unsigned long long r = mix32(seed + wave * 0x632BE59BD9B4E019ull, (unsigned long long)row, (unsigned long long)a) & 0xFFFFu;
                wgt[i] = (unsigned)(1ull + ((r * r * r) >> 32));
                sum += wgt[i];
            }
        }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
        const float total = (float)sum;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int a = lane + 32 * i;
            if (a < A) policy[row * A + a] = __fdiv_rn((float)wgt[i], total);
        }
        if (lane == 0) {
            const int vv = (int)(mix32(seed ^ 0x5555ull, (unsigned long long)row, wave) & 0xFFFFu) - 32768;
            value[row] = __fdiv_rn((float)vv, 32832.0f);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long t = atomicAdd(&wave_dev[1], 1ull);
        if (t == (unsigned long long)gridDim.x - 1ull) { wave_dev[1] = 0ull; wave_dev[0] = wave + 1ull; }
    }
}

// ------------------------------------------------------------------------------------------ host side
// the sqrt table of sqrt_of_count, once per device (std::sqrt of an exactly representable integer is correctly rounded)
static int ensure_tables() {
    static unsigned long long done = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if ((done >> (dev & 63)) & 1ull) return 0;
    static double tab[kSqrtTab];
    for (int i = 0; i < kSqrtTab; ++i) tab[i] = std::sqrt((double)i);
    const cudaError_t rc = cudaMemcpyToSymbol(c_sqrt_tab, tab, sizeof(tab));
    if (rc != cudaSuccess) return (int)rc;
    done |= 1ull << (dev & 63);
    return 0;
}

// Launch with programmatic dependent launch allowed: the grid may be scheduled while the previous kernel of the stream is still
// draining; the kernel itself waits (cudaGridDependencySynchronize) before it touches anything that kernel produced.  Inside a
// replayed CUDA graph this hides most of the ~3 us kernel-node dispatch gap of the three tree kernels of a wave.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static int g_force_sequential_select = 0;     // tg_mcts_debug_sequential_select: tests compare the two selection kernels

static int check_cfg(const TgMctsConfig* c) {
    if (!c) return set_err(-1, "tg_mcts: null config");
    if (c->n != 9 && c->n != 19) return set_err(-1, "tg_mcts: board size must be 9 or 19");
    if (c->num_games <= 0 || c->readouts < 1 || c->readouts > kMaxReadouts) return set_err(-1, "tg_mcts: bad num_games / readouts");
    if (c->wu_loss < 0 || c->wu_loss * c->readouts > 127) return set_err(-1, "tg_mcts: wu_loss * readouts must be <= 127");
    if (c->node_cap < 1 || c->edge_cap < c->n * c->n + 1 || c->max_depth < 2) return set_err(-1, "tg_mcts: capacities too small");
    // per-edge child info packs (edge base : 22 | edge count : 10) into one word (pack_kid): larger arenas would alias
    if (c->edge_cap > (1 << 22)) return set_err(-1, "tg_mcts: edge_cap must be <= 4194304 (edge base is a 22-bit field)");
    if (c->n * c->n + 1 > 1023) return set_err(-1, "tg_mcts: more than 1023 actions per node");
    return 0;
}
static inline unsigned warp_blocks(long long warps) { return (unsigned)((warps * 32 + kThreads - 1) / kThreads); }

#define TG_MCTS_N(c, ...)                        \
    if ((c)->n == 9) { constexpr int N = 9; __VA_ARGS__; } else { constexpr int N = 19; __VA_ARGS__; }

}  // namespace tg

using namespace tg;

extern "C" {

int64_t tg_mcts_workspace_bytes(const TgMctsConfig* cfg) {
    if (check_cfg(cfg)) return -1;
    TG_MCTS_N(cfg, return (int64_t)make_view<N>(cfg, nullptr).total_bytes);
}

int tg_mcts_init(const TgMctsConfig* cfg, void* ws, void* stream) {
    if (check_cfg(cfg) || !ws) return set_err(-1, "tg_mcts_init");
    if (ensure_tables()) return set_err(-1, "tg_mcts_init: constant tables");
    TG_MCTS_N(cfg, { View v = make_view<N>(cfg, ws); k_init<N><<<(v.G + 127) / 128, 128, 0, (cudaStream_t)stream>>>(v, nullptr); });
    return set_err((int)cudaGetLastError(), "tg_mcts_init");
}

int tg_mcts_reset_games(const TgMctsConfig* cfg, void* ws, const uint8_t* mask, void* stream) {
    if (check_cfg(cfg) || !ws || !mask) return set_err(-1, "tg_mcts_reset_games");
    TG_MCTS_N(cfg, { View v = make_view<N>(cfg, ws); k_init<N><<<(v.G + 127) / 128, 128, 0, (cudaStream_t)stream>>>(v, mask); });
    return set_err((int)cudaGetLastError(), "tg_mcts_reset_games");
}

void* tg_mcts_root_states(const TgMctsConfig* cfg, void* ws) {
    if (check_cfg(cfg) || !ws) return nullptr;
    TG_MCTS_N(cfg, return (void*)make_view<N>(cfg, ws).root_state);
}

int tg_mcts_root_obs(const TgMctsConfig* cfg, void* ws, float* obs_root, void* stream) {
    if (check_cfg(cfg) || !ws || !obs_root) return set_err(-1, "tg_mcts_root_obs");
    return tg_env_encode(cfg->n, tg_mcts_root_states(cfg, ws), cfg->num_games, obs_root, 32, stream);
}

int tg_mcts_root_expand(const TgMctsConfig* cfg, void* ws, const float* policy_root, const float* value_root, void* stream) {
    if (check_cfg(cfg) || !ws || !policy_root || !value_root) return set_err(-1, "tg_mcts_root_expand");
    TG_MCTS_N(cfg, {
        View v = make_view<N>(cfg, ws);
        const size_t smem = (size_t)(kThreads / 32) * 2 * Geo<N>::A * 4;
        k_root_expand<N><<<warp_blocks(v.G), kThreads, smem, (cudaStream_t)stream>>>(v, policy_root, value_root);
    });
    return set_err((int)cudaGetLastError(), "tg_mcts_root_expand");
}

int tg_mcts_begin_move(const TgMctsConfig* cfg, void* ws, const double* noise, void* stream) {
    if (check_cfg(cfg) || !ws) return set_err(-1, "tg_mcts_begin_move");
    TG_MCTS_N(cfg, { View v = make_view<N>(cfg, ws); k_begin_move<N><<<warp_blocks(v.G), kThreads, 0, (cudaStream_t)stream>>>(v, noise); });
    return set_err((int)cudaGetLastError(), "tg_mcts_begin_move");
}

int tg_mcts_wave_begin(const TgMctsConfig* cfg, void* ws, float* obs, void* stream) {
    if (check_cfg(cfg) || !ws || !obs) return set_err(-1, "tg_mcts_wave_begin");
    if ((reinterpret_cast<uintptr_t>(obs) & 3) != 0) return set_err(-1, "tg_mcts_wave_begin: obs must be 4-byte aligned");
    TG_MCTS_N(cfg, {
        View v = make_view<N>(cfg, ws);
        // one CTA per game with a warp per descent (pipelined through shared memory); the one-warp-per-game kernel remains for
        // configurations whose pipeline state does not fit the default shared-memory window
        size_t sel_smem = (size_t)2 * v.MAXD * v.R * sizeof(int);
        const size_t smem = (size_t)(kThreads / 32) * leaf_smem_words<N>() * 4;
        if (v.R > 1 && v.R <= kParWarps && sel_smem <= 40 * 1024 && !g_force_sequential_select) {
            // select + leaves in one launch: a game moves on to its leaves as soon as its own descents are done
            const size_t planes = (size_t)v.R * leaf_smem_words<N>() * 4;
            if (planes > sel_smem) sel_smem = planes;
            if (N == 9) launch_pdl(k_select_par<N, true>, dim3((unsigned)v.G), dim3(32 * v.R), sel_smem, (cudaStream_t)stream, v, obs);
            else {
                launch_pdl(k_select_par<N, false>, dim3((unsigned)v.G), dim3(32 * v.R), sel_smem, (cudaStream_t)stream, v, obs);
                k_leaves<N><<<warp_blocks((long long)v.G * v.R), kThreads, smem, (cudaStream_t)stream>>>(v, obs);
            }
        } else {
            k_select<N><<<warp_blocks(v.G), kThreads, 0, (cudaStream_t)stream>>>(v);
            k_leaves<N><<<warp_blocks((long long)v.G * v.R), kThreads, smem, (cudaStream_t)stream>>>(v, obs);
        }
    });
    return set_err((int)cudaGetLastError(), "tg_mcts_wave_begin");
}

int tg_mcts_pending(const TgMctsConfig* cfg, void* ws, int32_t* out, void* stream) {
    if (check_cfg(cfg) || !ws || !out) return set_err(-1, "tg_mcts_pending");
    TG_MCTS_N(cfg, { View v = make_view<N>(cfg, ws); k_pending<N><<<1, 256, 0, (cudaStream_t)stream>>>(v, out); });
    return set_err((int)cudaGetLastError(), "tg_mcts_pending");
}

int tg_mcts_debug_prefetch(int on) {
    return set_err((int)cudaMemcpyToSymbol(tg::g_prefetch_guess, &on, sizeof(on)), "tg_mcts_debug_prefetch");
}

int tg_mcts_debug_times(unsigned long long* times /* device, 4 x u64 per game, or NULL to stop */) {
    return set_err((int)cudaMemcpyToSymbol(tg::d_dbg_times, &times, sizeof(times)), "tg_mcts_debug_times");
}

int tg_mcts_debug_sequential_select(int on) { const int old = g_force_sequential_select; g_force_sequential_select = on; return old; }

int tg_mcts_wave_end(const TgMctsConfig* cfg, void* ws, const float* policy, const float* value, int32_t* pending, void* stream) {
    if (check_cfg(cfg) || !ws || !policy || !value) return set_err(-1, "tg_mcts_wave_end");
    TG_MCTS_N(cfg, {
        View v = make_view<N>(cfg, ws);
        const size_t smem = (size_t)(kThreads / 32) * 2 * Geo<N>::A * 4;
        if (v.R > 1 && v.R <= kParWarps && !g_force_sequential_select)
            launch_pdl(k_expand_par<N>, dim3((unsigned)v.G), dim3(32 * v.R), (size_t)v.R * 2 * Geo<N>::A * 4, (cudaStream_t)stream, v, policy, value, pending);
        else
            k_expand<N><<<warp_blocks(v.G), kThreads, smem, (cudaStream_t)stream>>>(v, policy, value, pending);
    });
    return set_err((int)cudaGetLastError(), "tg_mcts_wave_end");
}

int tg_mcts_root_stats(const TgMctsConfig* cfg, void* ws, int32_t* counts, float* W, float* var, float* prior, int32_t* root_n,
                       int32_t* root_k, int32_t* step, void* stream) {
    if (check_cfg(cfg) || !ws) return set_err(-1, "tg_mcts_root_stats");
    TG_MCTS_N(cfg, { View v = make_view<N>(cfg, ws); k_root_stats<N><<<warp_blocks(v.G), kThreads, 0, (cudaStream_t)stream>>>(v, counts, W, var, prior, root_n, root_k, step); });
    return set_err((int)cudaGetLastError(), "tg_mcts_root_stats");
}

int tg_mcts_advance(const TgMctsConfig* cfg, void* ws, const int16_t* actions, uint8_t* done_out, void* stream) {
    if (check_cfg(cfg) || !ws || !actions) return set_err(-1, "tg_mcts_advance");
    TG_MCTS_N(cfg, { View v = make_view<N>(cfg, ws); k_advance<N><<<warp_blocks(v.G), kThreads, 0, (cudaStream_t)stream>>>(v, actions, done_out); });
    return set_err((int)cudaGetLastError(), "tg_mcts_advance");
}

int tg_mcts_status(const TgMctsConfig* cfg, void* ws, uint32_t* status, unsigned long long* sims, void* stream) {
    if (check_cfg(cfg) || !ws) return set_err(-1, "tg_mcts_status");
    TG_MCTS_N(cfg, { View v = make_view<N>(cfg, ws); k_status<<<(v.G + 127) / 128, 128, 0, (cudaStream_t)stream>>>(v, status, sims); });
    return set_err((int)cudaGetLastError(), "tg_mcts_status");
}

int tg_mcts_search_stats(const TgMctsConfig* cfg, void* ws, unsigned long long* out3, void* stream) {
    if (check_cfg(cfg) || !ws || !out3) return set_err(-1, "tg_mcts_search_stats");
    TG_MCTS_N(cfg, { View v = make_view<N>(cfg, ws); k_search_stats<<<(v.G + 127) / 128, 128, 0, (cudaStream_t)stream>>>(v, out3); });
    return set_err((int)cudaGetLastError(), "tg_mcts_search_stats");
}

int tg_mcts_outcome(const TgMctsConfig* cfg, void* ws, uint8_t* winner, float* score, float* territory, void* stream) {
    if (check_cfg(cfg) || !ws || !score) return set_err(-1, "tg_mcts_outcome (score is required)");
    int rc = tg_env_score(cfg->n, tg_mcts_root_states(cfg, ws), cfg->num_games, cfg->komi, score, territory, 32, stream);
    if (rc) return rc;
    if (winner) k_winner<<<(cfg->num_games + 127) / 128, 128, 0, (cudaStream_t)stream>>>(score, winner, cfg->num_games);
    return set_err((int)cudaGetLastError(), "tg_mcts_outcome");
}

int tg_synth_provider(int n, int64_t rows, uint64_t seed, uint64_t wave, float* policy, float* value, void* stream) {
    if (!policy || !value || rows <= 0 || (n != 9 && n != 19)) return set_err(-1, "tg_synth_provider");
    k_synth<<<warp_blocks(rows), kThreads, 0, (cudaStream_t)stream>>>(n * n + 1, rows, seed, wave, nullptr, policy, value);
    return set_err((int)cudaGetLastError(), "tg_synth_provider");
}

int tg_synth_provider_dev(int n, int64_t rows, uint64_t seed, unsigned long long* wave_counter, float* policy, float* value, void* stream) {
    if (!policy || !value || !wave_counter || rows <= 0 || (n != 9 && n != 19)) return set_err(-1, "tg_synth_provider_dev");
    launch_pdl(k_synth_counted, dim3(warp_blocks(rows)), dim3(kThreads), 0, (cudaStream_t)stream, n * n + 1, (long long)rows, (unsigned long long)seed, wave_counter, policy, value);
    return set_err((int)cudaGetLastError(), "tg_synth_provider_dev");
}

}  // extern "C"
