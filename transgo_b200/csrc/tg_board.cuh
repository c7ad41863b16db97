// tg_board.cuh -- bitboard Go rules for sm_100a, tile-cooperative (T lanes per game, T in {1,..,32}).
//
// Replaces the reference rules engine GoEnv/cpp_src/{board,board_feature,go_env}.cc.  The reference keeps
// per-point linked lists and per-block liberty counters; here a position is two bitboards plus 16 B of meta,
// every rule is a whole-board bit operation, and the only per-group work (exact liberty counts) is spread
// over the lanes of a tile: each lane flood-fills one group from its "leader" stone, all lanes in lock step.
//
// Board layout in a word: rows are padded to N+1 bits (one always-zero guard bit) so that E/W shifts never
// wrap across rows; RPW rows share a word.  9x9: 3 rows per 32-bit word, 3 words.  19x19: 3 rows per 64-bit
// word, 7 words.  Point p = y*N + x (go_comm.h:54-56) lives in word y/RPW, bit (y%RPW)*(N+1)+x, so ascending
// (word, bit) order is ascending action order.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tg {

enum : int { kEmpty = 0, kBlack = 1, kWhite = 2 };          // go_comm.h:36-39
enum : int { kPass = -1, kResign = -2, kInvalid = -3 };     // go_comm.h:24-27

template <int N> struct WordSel { using type = uint64_t; };
template <> struct WordSel<9> { using type = uint32_t; };

template <int N>
struct Geo {
    using word_t = typename WordSel<N>::type;
    static constexpr int P = N * N;
    static constexpr int A = P + 1;                          // action P == PASS (go_env.cc:56-57)
    static constexpr int BITS = sizeof(word_t) * 8;
    static constexpr int STRIDE = N + 1;
    static constexpr int RPW = BITS / STRIDE;
    static constexpr int W = (N + RPW - 1) / RPW;
    static constexpr int STONE_U32 = 2 * W * (int)(sizeof(word_t) / 4);
    static constexpr int SW = ((STONE_U32 + 3) / 4) * 4 + 4; // state size in u32 (multiple of 4 -> uint4 planes)
    static constexpr int SQ = SW / 4;                        // state size in uint4
    static constexpr int DW = (A + 31) / 32;                 // dense action mask, u32 words (bit a == action a)
    static constexpr int DQ = (DW + 3) / 4;                  // ... in uint4
    static constexpr int PLANE_BITS = 10 * P;
    static constexpr int PLANE_W = (PLANE_BITS + 31) / 32;
    __host__ __device__ static constexpr word_t row() { return (word_t(1) << N) - 1; }
    __host__ __device__ static constexpr int rows_in(int i) { return (N - i * RPW) < RPW ? (N - i * RPW) : RPW; }
    __host__ __device__ static constexpr word_t wmask(int i) {
        word_t m = 0;
        for (int r = 0; r < rows_in(i); ++r) m |= row() << (r * STRIDE);
        return m;
    }
    // edge points (have at least one off-board diagonal)
    __host__ __device__ static constexpr word_t emask(int i) {
        word_t m = 0;
        for (int r = 0; r < rows_in(i); ++r) {
            int y = i * RPW + r;
            word_t rowm = (y == 0 || y == N - 1) ? row() : ((word_t(1) << (N - 1)) | word_t(1));
            m |= rowm << (r * STRIDE);
        }
        return m;
    }
};

__device__ __forceinline__ int popc(uint32_t x) { return __popc(x); }
__device__ __forceinline__ int popc(uint64_t x) { return __popcll(x); }
__device__ __forceinline__ int ffs0(uint32_t x) { return __ffs((int)x) - 1; }
__device__ __forceinline__ int ffs0(uint64_t x) { return __ffsll((long long)x) - 1; }

// index of the k-th (0-based) set bit of x; caller guarantees k < popc(x)
__device__ __forceinline__ int nth_bit(uint32_t x, int k) {
    int pos = 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        int c = __popc(x & ((1u << s) - 1));
        bool hi = k >= c;
        k -= hi ? c : 0;
        x = hi ? (x >> s) : x;
        pos += hi ? s : 0;
    }
    return pos;
}
__device__ __forceinline__ int nth_bit(uint64_t x, int k) {
    uint32_t lo = (uint32_t)x;
    int c = __popc(lo);
    return k < c ? nth_bit(lo, k) : 32 + nth_bit((uint32_t)(x >> 32), k - c);
}

template <int N>
struct BB {
    using G = Geo<N>;
    using word_t = typename G::word_t;
    word_t w[G::W];
};

#define TG_FOR_W _Pragma("unroll") for (int i = 0; i < Geo<N>::W; ++i)

template <int N> __device__ __forceinline__ BB<N> bb_zero() { BB<N> r; TG_FOR_W r.w[i] = 0; return r; }
template <int N> __device__ __forceinline__ BB<N> bb_full() { BB<N> r; TG_FOR_W r.w[i] = Geo<N>::wmask(i); return r; }
template <int N> __device__ __forceinline__ BB<N> operator&(const BB<N>& a, const BB<N>& b) { BB<N> r; TG_FOR_W r.w[i] = a.w[i] & b.w[i]; return r; }
template <int N> __device__ __forceinline__ BB<N> operator|(const BB<N>& a, const BB<N>& b) { BB<N> r; TG_FOR_W r.w[i] = a.w[i] | b.w[i]; return r; }
template <int N> __device__ __forceinline__ BB<N> andn(const BB<N>& a, const BB<N>& b) { BB<N> r; TG_FOR_W r.w[i] = a.w[i] & ~b.w[i]; return r; }   // a & ~b
template <int N> __device__ __forceinline__ BB<N> bb_not(const BB<N>& a) { BB<N> r; TG_FOR_W r.w[i] = ~a.w[i] & Geo<N>::wmask(i); return r; }
template <int N> __device__ __forceinline__ bool any(const BB<N>& a) { typename Geo<N>::word_t o = 0; TG_FOR_W o |= a.w[i]; return o != 0; }
template <int N> __device__ __forceinline__ bool differs(const BB<N>& a, const BB<N>& b) { typename Geo<N>::word_t o = 0; TG_FOR_W o |= a.w[i] ^ b.w[i]; return o != 0; }
template <int N> __device__ __forceinline__ int count(const BB<N>& a) { int c = 0; TG_FOR_W c += popc(a.w[i]); return c; }
template <int N> __device__ __forceinline__ BB<N> bb_select(bool p, const BB<N>& a, const BB<N>& b) { BB<N> r; TG_FOR_W r.w[i] = p ? a.w[i] : b.w[i]; return r; }

template <int N> __device__ __forceinline__ BB<N> shift_e(const BB<N>& a) { BB<N> r; TG_FOR_W r.w[i] = (a.w[i] << 1) & Geo<N>::wmask(i); return r; }   // x -> x+1
template <int N> __device__ __forceinline__ BB<N> shift_w(const BB<N>& a) { BB<N> r; TG_FOR_W r.w[i] = (a.w[i] >> 1) & Geo<N>::wmask(i); return r; }   // x -> x-1
template <int N> __device__ __forceinline__ BB<N> shift_s(const BB<N>& a) {                                                                       // y -> y+1
    using G = Geo<N>; BB<N> r;
    TG_FOR_W {
        typename G::word_t v = a.w[i] << G::STRIDE;
        if (i > 0) v |= a.w[i - 1] >> (G::STRIDE * (G::RPW - 1));
        r.w[i] = v & G::wmask(i);
    }
    return r;
}
template <int N> __device__ __forceinline__ BB<N> shift_n(const BB<N>& a) {                                                                       // y -> y-1
    using G = Geo<N>; BB<N> r;
    TG_FOR_W {
        typename G::word_t v = a.w[i] >> G::STRIDE;
        if (i < G::W - 1) v |= a.w[i + 1] << (G::STRIDE * (G::RPW - 1));
        r.w[i] = v & G::wmask(i);
    }
    return r;
}
// the 4-neighbourhood of a set (self excluded unless reached from a neighbour); go_comm.h:43-44
template <int N> __device__ __forceinline__ BB<N> nbrs(const BB<N>& a) {
    using G = Geo<N>; BB<N> r;
    TG_FOR_W {
        typename G::word_t v = (a.w[i] << 1) | (a.w[i] >> 1) | (a.w[i] << G::STRIDE) | (a.w[i] >> G::STRIDE);
        if (i > 0) v |= a.w[i - 1] >> (G::STRIDE * (G::RPW - 1));
        if (i < G::W - 1) v |= a.w[i + 1] << (G::STRIDE * (G::RPW - 1));
        r.w[i] = v & G::wmask(i);
    }
    return r;
}

template <int N> __device__ __forceinline__ void pt_to_wb(int p, int& wi, int& b) {
    using G = Geo<N>;
    int y = p / N, x = p - y * N;
    wi = y / G::RPW;
    b = (y - wi * G::RPW) * G::STRIDE + x;
}
template <int N> __device__ __forceinline__ int wb_to_pt(int wi, int b) {
    using G = Geo<N>;
    int r = b / G::STRIDE;
    return (wi * G::RPW + r) * N + (b - r * G::STRIDE);
}
template <int N> __device__ __forceinline__ BB<N> bb_point(int p) {     // p must be on board
    using word_t = typename Geo<N>::word_t;
    int wi, b; pt_to_wb<N>(p, wi, b);
    BB<N> r; TG_FOR_W r.w[i] = (i == wi) ? (word_t(1) << b) : word_t(0);
    return r;
}
template <int N> __device__ __forceinline__ bool bb_test(const BB<N>& a, int p) {
    using word_t = typename Geo<N>::word_t;
    int wi, b; pt_to_wb<N>(p, wi, b);
    word_t v = 0; TG_FOR_W v = (i == wi) ? a.w[i] : v;
    return (v >> b) & 1;
}
template <int N> __device__ __forceinline__ int bb_lowest(const BB<N>& a) {   // lowest point, a must be non-empty
    int p = -1;
#pragma unroll
    for (int i = Geo<N>::W - 1; i >= 0; --i) if (a.w[i]) p = wb_to_pt<N>(i, ffs0(a.w[i]));
    return p;
}
// word index + bit index of the k-th (0-based, ascending action order) point of a set; wi = -1 when k is out of range.
template <int N> __device__ __forceinline__ void bb_locate(const BB<N>& a, int k, int& wi, int& bit) {
    using word_t = typename Geo<N>::word_t;
    wi = -1;
    word_t sel = 0;
    int kk = 0;
    TG_FOR_W {
        const int c = popc(a.w[i]);
        const bool here = wi < 0 && k >= 0 && k < c;
        sel = here ? a.w[i] : sel;
        kk = here ? k : kk;
        wi = here ? i : wi;
        k -= c;
    }
    bit = nth_bit(sel, kk);
}
// k-th point of a set; k < count(a)
template <int N> __device__ __forceinline__ int bb_nth(const BB<N>& a, int k) {
    int wi, bit; bb_locate<N>(a, k, wi, bit);
    return wb_to_pt<N>(wi, bit);
}
// single-bit set holding the k-th point of a (empty set when k is out of range)
template <int N> __device__ __forceinline__ BB<N> bb_nth_bit(const BB<N>& a, int k) {
    using word_t = typename Geo<N>::word_t;
    int wi, bit; bb_locate<N>(a, k, wi, bit);
    BB<N> r; TG_FOR_W r.w[i] = (i == wi) ? (word_t(1) << bit) : word_t(0);
    return r;
}

// dense mask: bit a of a little-endian u32 string <-> point a.  d must hold Geo<N>::DW words, pre-zeroed by caller.
template <int N> __device__ __forceinline__ void bb_to_dense(const BB<N>& a, uint32_t* d) {
    using G = Geo<N>;
#pragma unroll
    for (int y = 0; y < N; ++y) {
        const int wi = y / G::RPW, sh = (y % G::RPW) * G::STRIDE;
        uint32_t rowbits = (uint32_t)((a.w[wi] >> sh) & G::row());
        const int off = y * N, dw = off >> 5, ds = off & 31;
        d[dw] |= rowbits << ds;
        if (ds + N > 32) d[dw + 1] |= rowbits >> (32 - ds);
    }
}
template <int N> __device__ __forceinline__ BB<N> dense_to_bb(const uint32_t* d) {
    using G = Geo<N>; using word_t = typename G::word_t;
    BB<N> r = bb_zero<N>();
#pragma unroll
    for (int y = 0; y < N; ++y) {
        const int off = y * N, dw = off >> 5, ds = off & 31;
        uint32_t rowbits = d[dw] >> ds;
        if (ds + N > 32) rowbits |= d[dw + 1] << (32 - ds);
        rowbits &= (uint32_t)G::row();
        const int wi = y / G::RPW, sh = (y % G::RPW) * G::STRIDE;
        r.w[wi] |= word_t(rowbits) << sh;
    }
    return r;
}

// ------------------------------------------------------------------------------------------------------------
// Position = the observable content of the reference Board + GoState (board.h:26-58, go_env.h:14-17).
template <int N>
struct Pos {
    BB<N> stones[2];       // [0] black, [1] white
    int last1, last2;      // board.h:43-44
    int ko_loc, ko_age;    // board.h:47,49
    int step;              // board.h:42 (1-based)
    int player;            // board.h:41 (1 black, 2 white)
    int ko_color;          // board.h:48
    int terminated;        // go_env.h:16
};

// colour-indexed access without dynamic register indexing (which would spill the struct to local memory)
template <int N> __device__ __forceinline__ BB<N> stones_of(const Pos<N>& s, int colour /*1|2*/) { return bb_select(colour == kBlack, s.stones[0], s.stones[1]); }
template <int N> __device__ __forceinline__ void set_stones(Pos<N>& s, int colour, const BB<N>& v) {
    TG_FOR_W { s.stones[0].w[i] = (colour == kBlack) ? v.w[i] : s.stones[0].w[i]; s.stones[1].w[i] = (colour == kBlack) ? s.stones[1].w[i] : v.w[i]; }
}

template <int N> __device__ __forceinline__ void pos_reset(Pos<N>& s) {      // board.cc:13-26
    s.stones[0] = bb_zero<N>(); s.stones[1] = bb_zero<N>();
    s.last1 = s.last2 = s.ko_loc = kInvalid;
    s.ko_age = 0; s.step = 1; s.player = kBlack; s.ko_color = 0; s.terminated = 0;
}

// State <-> u32[SW] serialisation (SW/4 uint4).  q(k) yields the k-th uint4 of the state.
template <int N> __device__ __forceinline__ void pos_pack(const Pos<N>& s, uint32_t* u) {
    using G = Geo<N>;
#pragma unroll
    for (int i = 0; i < G::SW; ++i) u[i] = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c)
        TG_FOR_W {
            if constexpr (sizeof(typename G::word_t) == 4) u[c * G::W + i] = (uint32_t)s.stones[c].w[i];
            else { u[(c * G::W + i) * 2] = (uint32_t)s.stones[c].w[i]; u[(c * G::W + i) * 2 + 1] = (uint32_t)((uint64_t)s.stones[c].w[i] >> 32); }
        }
    const int m = G::SW - 4;
    u[m + 0] = ((uint32_t)(uint16_t)(int16_t)s.last1) | ((uint32_t)(uint16_t)(int16_t)s.last2 << 16);
    u[m + 1] = ((uint32_t)(uint16_t)(int16_t)s.ko_loc) | ((uint32_t)(uint16_t)(int16_t)s.ko_age << 16);
    u[m + 2] = ((uint32_t)(uint16_t)s.step) | ((uint32_t)(s.player & 0xff) << 16) | ((uint32_t)(s.ko_color & 0xff) << 24);
    u[m + 3] = (uint32_t)(s.terminated & 1);
}
template <int N> __device__ __forceinline__ void pos_unpack(Pos<N>& s, const uint32_t* u) {
    using G = Geo<N>; using word_t = typename G::word_t;
#pragma unroll
    for (int c = 0; c < 2; ++c)
        TG_FOR_W {
            if constexpr (sizeof(word_t) == 4) s.stones[c].w[i] = (word_t)u[c * G::W + i];
            else s.stones[c].w[i] = (word_t)((uint64_t)u[(c * G::W + i) * 2] | ((uint64_t)u[(c * G::W + i) * 2 + 1] << 32));
        }
    const int m = G::SW - 4;
    s.last1 = (int16_t)(u[m] & 0xffff); s.last2 = (int16_t)(u[m] >> 16);
    s.ko_loc = (int16_t)(u[m + 1] & 0xffff); s.ko_age = (int16_t)(u[m + 1] >> 16);
    s.step = (int)(u[m + 2] & 0xffff); s.player = (int)((u[m + 2] >> 16) & 0xff); s.ko_color = (int)(u[m + 2] >> 24);
    s.terminated = (int)(u[m + 3] & 1);
}
// strided access: element k of state `idx` sits at base[k * kstride + idx * istride]
//   SoA planes over G games : kstride = G, istride = 1      AoS records : kstride = 1, istride = SQ
template <int N> __device__ __forceinline__ void pos_load(Pos<N>& s, const uint4* __restrict__ base, long long idx, long long kstride, long long istride) {
    using G = Geo<N>;
    uint32_t u[G::SW];
#pragma unroll
    for (int k = 0; k < G::SQ; ++k) {
        uint4 q = base[k * kstride + idx * istride];
        u[4 * k] = q.x; u[4 * k + 1] = q.y; u[4 * k + 2] = q.z; u[4 * k + 3] = q.w;
    }
    pos_unpack<N>(s, u);
}
template <int N> __device__ __forceinline__ void pos_store(const Pos<N>& s, uint4* __restrict__ base, long long idx, long long kstride, long long istride) {
    using G = Geo<N>;
    uint32_t u[G::SW];
    pos_pack<N>(s, u);
#pragma unroll
    for (int k = 0; k < G::SQ; ++k) base[k * kstride + idx * istride] = make_uint4(u[4 * k], u[4 * k + 1], u[4 * k + 2], u[4 * k + 3]);
}

// ------------------------------------------------------------------------------------------------------------
// Tile helpers.  A tile is T consecutive lanes of a warp working on one game (T a power of two <= 32).
template <int T> struct Tile {
    static_assert(T >= 1 && T <= 32 && (T & (T - 1)) == 0, "tile width");
    __device__ static __forceinline__ int lane() { return (int)(threadIdx.x & (T - 1)); }
    __device__ static __forceinline__ unsigned mask() {
        if constexpr (T == 32) return 0xffffffffu;
        else return ((1u << T) - 1u) << ((threadIdx.x & 31) & ~(T - 1));
    }
    __device__ static __forceinline__ bool any(bool p) {
        if (T == 1) return p;
        if (T == 32) return __any_sync(0xffffffffu, p);
        return (__ballot_sync(mask(), p)) != 0;   // ballot with a member mask only reports member lanes
    }
    template <typename W> __device__ static __forceinline__ W reduce_or(W v) {
        if (T == 1) return v;
        if (sizeof(W) == 4) {
            if (T == 32) return (W)__reduce_or_sync(0xffffffffu, (unsigned)v);
            unsigned x = (unsigned)v;
#pragma unroll
            for (int s = T / 2; s >= 1; s >>= 1) x |= __shfl_xor_sync(mask(), x, s);
            return (W)x;
        } else {
            unsigned long long x = (unsigned long long)v;
            unsigned lo = (unsigned)x, hi = (unsigned)(x >> 32);
            lo = reduce_or<unsigned>(lo); hi = reduce_or<unsigned>(hi);
            return (W)((unsigned long long)lo | ((unsigned long long)hi << 32));
        }
    }
    __device__ static __forceinline__ void sync() { if (T > 1) __syncwarp(mask()); }
};
template <int N, int T> __device__ __forceinline__ BB<N> tile_or(const BB<N>& a) {
    BB<N> r; TG_FOR_W r.w[i] = Tile<T>::template reduce_or<typename Geo<N>::word_t>(a.w[i]); return r;
}
// the same when the whole warp is known to be converged (warp-uniform kernels): butterflies under the constant full mask
// stay inside the aligned T-lane tile and spare the variable-mask synchronisation protocol
template <int N, int T> __device__ __forceinline__ BB<N> tile_or_converged(const BB<N>& a) {
    using word_t = typename Geo<N>::word_t;
    if constexpr (T == 1) return a;
    else if constexpr (T == 32) return tile_or<N, 32>(a);
    else {
        BB<N> r = a;
#pragma unroll
        for (int s = T / 2; s >= 1; s >>= 1)
            TG_FOR_W {
                if constexpr (sizeof(word_t) == 4) r.w[i] |= __shfl_xor_sync(0xffffffffu, r.w[i], s);
                else r.w[i] |= (word_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)r.w[i], s);
            }
        return r;
    }
}

// Loop / branch votes.  WU ("warp-uniform"): the tiles that share a warp vote together, so every lane of the warp follows
// the same control flow -- a game whose own condition is false runs harmless extra iterations instead of forcing the
// warp to execute the two games' paths one after the other.
template <int T, bool WU> __device__ __forceinline__ bool vote_any(bool p) {
    if (WU && T < 32) return __any_sync(0xffffffffu, p);
    return Tile<T>::any(p);
}

// a condition that is already the same in every lane of a tile (it depends on the position only): WU widens it to the warp
template <int T, bool WU> __device__ __forceinline__ bool uniform_cond(bool p) {
    if (WU && T < 32) return __any_sync(0xffffffffu, p);
    return p;
}

// Eastward fill of the seeds g (a subset of mk) to the end of their rows' runs in mk: adding the seeds to the mask lets the
// carry ripple up each run of ones; the always-zero guard bit between rows stops it.
template <int N> __device__ __forceinline__ BB<N> efill(const BB<N>& g, const BB<N>& mk) {
    using word_t = typename Geo<N>::word_t;
    BB<N> r;
    TG_FOR_W { const word_t t = mk.w[i] + g.w[i]; r.w[i] = ((t ^ mk.w[i]) & mk.w[i]) | g.w[i]; }
    return r;
}

// flood `seed` through `mask` (4-connectivity) until no lane of the tile changes any more.
template <int N, int T, bool WU = false> __device__ __forceinline__ BB<N> flood(BB<N> g, const BB<N>& mask) {
    bool ch = any(g);
    while (vote_any<T, WU>(ch)) {
        BB<N> n1 = efill<N>((g | nbrs(g)) & mask, mask);
        BB<N> n2 = efill<N>((n1 | nbrs(n1)) & mask, mask);
        ch = differs(n2, g);
        g = n2;
    }
    return g;
}

// Flood the group of `seed` inside `col` and count its liberties, but stop caring once `lim` liberties are known:
// callers only distinguish 1 / 2 / >=3 liberties (board_feature.cc:17-41, board.cc:130-158), and big groups reach the
// limit in the first step.  On return `complete` tells whether g is the whole group (then libs is exact); otherwise
// libs >= lim.  The loop is tile-uniform: it runs until no lane still needs another step.
template <int N, int T, bool WU = false>
__device__ __forceinline__ void flood_group(BB<N>& g, const BB<N>& col, const BB<N>& empty, int lim, int& libs, bool& complete) {
    bool ch = any(g);
    libs = 0;
    complete = true;
    while (true) {
        const BB<N> nb = nbrs(g);
        // one step in every direction, then east to the end of each row's run: floods start at a group's north-west corner
        // (its leader stone), so most of the growth is eastward and southward
        const BB<N> n1 = efill<N>((g | nb) & col, col);
        libs = count(nb & empty);
        ch = differs(n1, g);
        g = n1;
        complete = !ch;
        if (!vote_any<T, WU>(ch && libs < lim)) break;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Group analysis: exact liberty classes of every group (what Block::liberties holds, board.h:22), plus, on
// request, the "lives" set of board_feature.cc:164-182 / board.cc:731-817 for colour `me`.
template <int N>
struct Groups {
    BB<N> l1[2];     // stones of groups with exactly 1 liberty, per colour
    BB<N> l2[2];     // ... exactly 2 liberties
    BB<N> lives;     // stones of `me` groups that satisfy GivenBlockLives (only when WANT_LIVES)
};
template <int N> __device__ __forceinline__ BB<N> l1_of(const Groups<N>& g, int colour) { return bb_select(colour == kBlack, g.l1[0], g.l1[1]); }
template <int N> __device__ __forceinline__ BB<N> l2_of(const Groups<N>& g, int colour) { return bb_select(colour == kBlack, g.l2[0], g.l2[1]); }

// true eyes of colour `own` against `opp` (board.cc:665-714)
template <int N> __device__ __forceinline__ BB<N> true_eyes(const BB<N>& own, const BB<N>& opp) {
    using G = Geo<N>;
    BB<N> empty = bb_not(own | opp);
    BB<N> eye = andn(empty, nbrs(bb_not(own)));                 // every on-board neighbour is own (walls are free)
    BB<N> on = shift_n(opp), os = shift_s(opp);
    BB<N> d1 = shift_e(on), d2 = shift_w(on), d3 = shift_e(os), d4 = shift_w(os);
    BB<N> anyd = d1 | d2 | d3 | d4;
    BB<N> twod = (d1 & d2) | ((d1 | d2) & d3) | ((d1 | d2 | d3) & d4);
    BB<N> fake;
    TG_FOR_W fake.w[i] = (G::emask(i) & anyd.w[i]) | (~G::emask(i) & twod.w[i]);
    return andn(eye, fake);
}

// WANT_L2: also classify "exactly two liberties" (feature planes); without it only l1 is produced (rules).
template <int N, int T, bool WANT_LIVES, bool WANT_L2 = WANT_LIVES>
__device__ __forceinline__ void analyse_groups(const BB<N>& black, const BB<N>& white, int me /*1|2*/, Groups<N>& out) {
    using G = Geo<N>;
    const BB<N> empty = bb_not(black | white);
    // leaders: stones with no same-colour neighbour to the west or north; every group owns at least one
    const BB<N> lb = andn(andn(black, shift_e(black)), shift_s(black));
    const BB<N> lw = andn(andn(white, shift_e(white)), shift_s(white));
    const int nb = count(lb), total = nb + count(lw);
    BB<N> eyes_me = bb_zero<N>();
    bool has_eyes = false;
    if (WANT_LIVES) {
        eyes_me = (me == kBlack) ? true_eyes(black, white) : true_eyes(white, black);
        has_eyes = any(eyes_me);
    }
    BB<N> p1b = bb_zero<N>(), p1w = p1b, p2b = p1b, p2w = p1b, plv = p1b;
    for (int base = 0; base < total; base += T) {
        const int j = base + Tile<T>::lane();
        const bool valid = j < total;
        const bool is_w = j >= nb;
        BB<N> g = bb_nth_bit<N>(is_w ? lw : lb, valid ? (is_w ? j - nb : j) : -1);
        const BB<N> col = bb_select(is_w, white, black);
        const bool mine = WANT_LIVES && has_eyes && valid && ((me == kWhite) == is_w);
        int c; bool complete;
        flood_group<N, T>(g, col, empty, mine ? 1 << 20 : (WANT_L2 ? 3 : 2), c, complete);
        const bool is1 = valid && complete && c == 1, is2 = WANT_L2 && valid && complete && c == 2;
        TG_FOR_W {
            const typename G::word_t m1 = is1 ? g.w[i] : 0, m2 = is2 ? g.w[i] : 0;
            p1b.w[i] |= is_w ? 0 : m1; p1w.w[i] |= is_w ? m1 : 0;
            if (WANT_L2) { p2b.w[i] |= is_w ? 0 : m2; p2w.w[i] |= is_w ? m2 : 0; }
        }
        if (WANT_LIVES && has_eyes) {                               // tile-uniform
            // candidate eyes of this block, then the diagonal test of board.cc:775-815
            BB<N> E = nbrs(g) & eyes_me;
            BB<N> bad = bb_not(col | E);                            // on-board points that are neither own stones nor candidate eyes
            BB<N> bn = shift_n(bad), bs = shift_s(bad);
            BB<N> d1 = shift_e(bn), d2 = shift_w(bn), d3 = shift_e(bs), d4 = shift_w(bs);
            BB<N> anyd = d1 | d2 | d3 | d4;
            BB<N> twod = (d1 & d2) | ((d1 | d2) & d3) | ((d1 | d2 | d3) & d4);
            BB<N> good;
            TG_FOR_W good.w[i] = E.w[i] & ((G::emask(i) & ~anyd.w[i]) | (~G::emask(i) & ~twod.w[i]));
            const bool alive = mine && c >= 2 && count(good) >= 2;
            TG_FOR_W plv.w[i] |= alive ? g.w[i] : 0;
        }
    }
    out.l1[0] = tile_or<N, T>(p1b); out.l1[1] = tile_or<N, T>(p1w);
    if (WANT_L2) { out.l2[0] = tile_or<N, T>(p2b); out.l2[1] = tile_or<N, T>(p2w); }
    else { out.l2[0] = bb_zero<N>(); out.l2[1] = bb_zero<N>(); }
    if (WANT_LIVES && has_eyes) out.lives = tile_or<N, T>(plv); else out.lives = bb_zero<N>();
}

// Seeded analysis: atari classes of ONLY the groups that hold a stone of `seeds`.  The rules need liberty counts in two
// places -- StoneLibertyAnalysis of a point with no empty neighbour (board.cc:90-158) and the capture test around the
// stone just played (board.cc:606-630) -- so a playout ply floods the handful of groups next to those points instead of
// every group on the board.  l1b / l1w are OR-ed into (callers pass zero or an earlier partial result).
template <int N, int T, bool WU = false>
__device__ __forceinline__ void analyse_seeded(const BB<N>& black, const BB<N>& white, const BB<N>& seeds, BB<N>& l1b, BB<N>& l1w) {
    using G = Geo<N>;
    const BB<N> empty = bb_not(black | white);
    BB<N> sb = seeds & black, sw = seeds & white;
    if (T == 1) {
        BB<N> p1b = bb_zero<N>(), p1w = p1b;
        // one thread: take the lowest remaining seed, flood its group, strike the flooded stones from the seed sets
        while (any(sb | sw)) {
            const bool is_w = !any(sb);
            const BB<N> src = is_w ? sw : sb;
            BB<N> g; bool taken = false;
            TG_FOR_W { g.w[i] = taken ? 0 : (src.w[i] & (0 - src.w[i])); taken = taken || src.w[i] != 0; }
            const BB<N> col = bb_select(is_w, white, black);
            int c; bool complete;
            flood_group<N, 1>(g, col, empty, 2, c, complete);
            const bool is1 = complete && c == 1;
            TG_FOR_W {
                const typename G::word_t m1 = is1 ? g.w[i] : 0;
                p1b.w[i] |= is_w ? 0 : m1; p1w.w[i] |= is_w ? m1 : 0;
                sb.w[i] &= ~g.w[i]; sw.w[i] &= ~g.w[i];
            }
        }
        l1b = l1b | p1b; l1w = l1w | p1w;
        return;
    }
    // seeds whose west / north neighbour is a same-colour seed belong to that seed's group: drop them
    sb = andn(andn(sb, shift_e(sb)), shift_s(sb));
    sw = andn(andn(sw, shift_e(sw)), shift_s(sw));
    // Each pass gives the first T seeds to the T lanes.  When seeds are left over, everything the floods of this pass reached
    // is struck from them first: a group usually owns several seeds (every "north-west corner" of its shape is one), so the
    // second pass rarely has much to do.  Atari groups of both colours are collected in ONE set (a group has one colour, so
    // the set is split by colour after the reduction: half the reduction traffic).
    BB<N> p1 = bb_zero<N>();
    bool any_pass = false;
    while (true) {
        const int nb = count(sb), total = nb + count(sw);
        if (!uniform_cond<T, WU>(total > 0)) break;                 // tile- (WU: warp-) uniform
        any_pass = true;
        const int j = Tile<T>::lane();
        const bool valid = j < total;
        const bool is_w = j >= nb;
        BB<N> g = bb_nth_bit<N>(is_w ? sw : sb, valid ? (is_w ? j - nb : j) : -1);
        const BB<N> col = bb_select(is_w, white, black);
        int c; bool complete;
        flood_group<N, T, WU>(g, col, empty, 2, c, complete);
        const bool is1 = valid && complete && c == 1;
        TG_FOR_W p1.w[i] |= is1 ? g.w[i] : 0;
        if (!uniform_cond<T, WU>(total > T)) break;
        const BB<N> covered = WU ? tile_or_converged<N, T>(g) : tile_or<N, T>(g);     // includes the T seeds just used
        sb = andn(sb, covered); sw = andn(sw, covered);
    }
    if (WU || any_pass) {
        const BB<N> all1 = WU ? tile_or_converged<N, T>(p1) : tile_or<N, T>(p1);
        l1b = l1b | (all1 & black); l1w = l1w | (all1 & white);
    }
}

// Empty points with no empty neighbour: the only points whose legality depends on group liberties.
template <int N> __device__ __forceinline__ BB<N> enclosed_points(const BB<N>& black, const BB<N>& white) {
    const BB<N> empty = bb_not(black | white);
    return andn(empty, nbrs(empty));
}

// Legal board points of the player to move.  Group liberties matter only next to enclosed points, so nothing is flooded
// when there is none.  Otherwise a single thread (T == 1) floods just the groups next to them (its cost is the SUM over
// groups), while a wider tile floods every group at once, one per lane in lock step (its cost is the MAX over groups, and
// the result then also serves the capture test of the move).  On return l1b / l1w hold the atari classes of the analysed
// groups and `known` = the points whose neighbour groups are all analysed (apply_stone_lazy wants both).
template <int N, int T, bool WU = false>
__device__ __forceinline__ BB<N> legal_points_lazy(const Pos<N>& s, BB<N>& l1b, BB<N>& l1w, BB<N>& known) {
    const BB<N>& black = s.stones[0]; const BB<N>& white = s.stones[1];
    const BB<N> stones = black | white;
    const BB<N> empty = bb_not(stones);
    const BB<N> encl = andn(empty, nbrs(empty));
    l1b = bb_zero<N>(); l1w = bb_zero<N>();
    known = encl;
    if (uniform_cond<T, WU>(any(encl))) {                                                       // tile- (WU: warp-) uniform branch
        if (T == 1) analyse_seeded<N, T>(black, white, nbrs(encl) & stones, l1b, l1w);
        else { analyse_seeded<N, T, WU>(black, white, stones, l1b, l1w); known = bb_full<N>(); }
    }
    const bool me_black = s.player == kBlack;
    const BB<N> own = bb_select(me_black, black, white);
    const BB<N> l1own = bb_select(me_black, l1b, l1w), l1opp = bb_select(me_black, l1w, l1b);
    // suicide (board.cc:130-158): enclosed, no own neighbour group with >1 liberty, no enemy neighbour group in atari
    BB<N> lg = andn(empty, andn(encl, nbrs(andn(own, l1own) | l1opp)));
    if (s.ko_age == 0 && s.ko_color == s.player && s.ko_loc >= 0 && s.ko_loc < Geo<N>::P)   // board.cc:198-200
        lg = andn(lg, bb_point<N>(s.ko_loc));
    return lg;
}

// One TRUSTED board move (the caller took pb from legal_points_lazy of this very position), given the partial atari sets
// that call produced.  Same state transition as apply_move (go_env.cc:51-80, board.cc:546-661).
template <int N, int T, bool WU = false>
__device__ __forceinline__ bool apply_stone_lazy(Pos<N>& s, const BB<N>& pb, int action, BB<N> l1b, BB<N> l1w, const BB<N>& known, int max_step,
                                                 bool placing = true) {
    const bool me_black = s.player == kBlack;
    const BB<N> own = bb_select(me_black, s.stones[0], s.stones[1]), opp = bb_select(me_black, s.stones[1], s.stones[0]);
    const BB<N> around = nbrs(pb);
    const BB<N> touch = around & opp;
    if (uniform_cond<T, WU>(any(touch) && !any(pb & known)))               // groups next to the stone not analysed yet
        analyse_seeded<N, T, WU>(s.stones[0], s.stones[1], touch, l1b, l1w);
    const BB<N> l1opp = bb_select(me_black, l1w, l1b);
    const BB<N> hit = around & l1opp;                                 // enemy neighbours in atari: they die
    bool ko = false;
    if (uniform_cond<T, WU>(any(hit))) {                                  // tile- (WU: warp-) uniform
        const BB<N> empty = bb_not(own | opp);
        if (any(hit) && !any(around & (empty | own))) {               // board.cc:163-194, 561-570
            const BB<N> cand = andn(hit, nbrs(opp));
            ko = count(cand) == 1;
            if (ko) { s.ko_loc = bb_lowest(cand); s.ko_color = 3 - s.player; s.ko_age = 0; }
        }
        const BB<N> dead = flood<N, T, WU>(hit, l1opp);               // board.cc:606-630
        const BB<N> left = andn(opp, dead);
        TG_FOR_W { s.stones[0].w[i] = me_black ? s.stones[0].w[i] : left.w[i]; s.stones[1].w[i] = me_black ? left.w[i] : s.stones[1].w[i]; }
    }
    if (!placing) return false;
    if (!ko) s.ko_age += 1;
    TG_FOR_W { s.stones[0].w[i] |= me_black ? pb.w[i] : 0; s.stones[1].w[i] |= me_black ? 0 : pb.w[i]; }
    s.player = 3 - s.player; s.last2 = s.last1; s.last1 = action; s.step += 1;
    if (s.step > max_step) { s.terminated = 1; return true; }           // go_env.cc:67-73
    return false;
}
// Pass (action P or -1) / resign (-2): board.cc:536-542, 554-558, 656-661; ko fields untouched.
template <int N>
__device__ __forceinline__ bool apply_pass(Pos<N>& s, int action, int max_step) {
    s.player = 3 - s.player; s.last2 = s.last1; s.last1 = action; s.step += 1;
    const bool done = s.step > 1 && ((s.last1 == kPass && s.last2 == kPass) || s.last1 == kResign);
    if (done || s.step > max_step) { s.terminated = 1; return true; }
    return done;
}

// Points that are suicide for `me` (board.cc:130-158 over every empty point, ko ignored): no empty neighbour,
// no own neighbour group with >1 liberty, no enemy neighbour group with exactly 1 liberty.
template <int N> __device__ __forceinline__ BB<N> suicide_points(const BB<N>& own, const BB<N>& opp, const BB<N>& l1own, const BB<N>& l1opp) {
    BB<N> empty = bb_not(own | opp);
    return andn(empty, nbrs(empty | andn(own, l1own) | l1opp));
}

// Legal board points for the player to move (board.cc:467-489 / TryPlay :437-464).
template <int N> __device__ __forceinline__ BB<N> legal_points(const Pos<N>& s, const Groups<N>& g) {
    const int me = s.player, op = 3 - s.player;
    BB<N> empty = bb_not(s.stones[0] | s.stones[1]);
    BB<N> lg = andn(empty, suicide_points(stones_of(s, me), stones_of(s, op), l1_of(g, me), l1_of(g, op)));
    if (s.ko_age == 0 && s.ko_color == s.player && s.ko_loc >= 0 && s.ko_loc < Geo<N>::P)   // board.cc:198-200
        lg = andn(lg, bb_point<N>(s.ko_loc));
    return lg;
}

// One move.  `action` in [0,P) places a stone, P or -1 passes, -2 resigns (go_env.cc:51-80, board.cc:546-661).
// Returns done; *accepted = false when the reference would print "invalid action" and keep the state.
// g must be the analysis of the position BEFORE the move (only l1 is used).
// trusted = the caller took `action` from legal_points() of this very position (playouts, MCTS): skip the re-check.
template <int N, int T>
__device__ __forceinline__ bool apply_move(Pos<N>& s, const Groups<N>& g, int action, int max_step, bool& accepted, bool trusted = false) {
    using GG = Geo<N>;
    accepted = false;
    if (s.terminated) return true;                                   // go_env.cc:52-55
    if (action == GG::P) action = kPass;                             // go_env.cc:56-57
    const int me = s.player, op = 3 - s.player;
    bool done = false;
    if (action == kPass || action == kResign) {
        s.player = 3 - s.player; s.last2 = s.last1; s.last1 = action; s.step += 1;          // board.cc:536-542
        done = s.step > 1 && ((s.last1 == kPass && s.last2 == kPass) || s.last1 == kResign);   // board.cc:656-661
    } else {
        if (action < 0 || action >= GG::P) return false;
        const BB<N> pb = bb_point<N>(action);
        if (!trusted && !any(legal_points(s, g) & pb)) return false; // occupied / ko / suicide (board.cc:448-462)
        const BB<N> own = stones_of(s, me), opp = stones_of(s, op);
        const BB<N> l1opp = l1_of(g, op);
        const BB<N> empty = bb_not(own | opp);
        const BB<N> around = nbrs(pb);
        // simple-ko bookkeeping, decided before captures (board.cc:163-194, 561-570): no empty and no own neighbour, and
        // exactly one enemy neighbour block that is a single stone in atari
        const BB<N> hit = around & l1opp;                            // enemy neighbours in atari: they die
        bool ko = false;
        if (any(hit) && !any(around & (empty | own))) {              // rare, tile-uniform
            const BB<N> cand = andn(hit, nbrs(opp));                 // ... that are single stones
            ko = count(cand) == 1;
            if (ko) { s.ko_loc = bb_lowest(cand); s.ko_color = 3 - s.player; s.ko_age = 0; }
        }
        if (!ko) s.ko_age += 1;
        // captures: every enemy neighbour group whose single liberty was this point (board.cc:606-630).
        // Distinct groups inside l1[op] never touch, so flooding inside l1[op] lifts exactly the seeded groups.
        BB<N> dead = flood<N, T>(hit, l1opp);
        set_stones(s, op, andn(opp, dead));
        set_stones(s, me, own | pb);
        s.player = 3 - s.player; s.last2 = s.last1; s.last1 = action; s.step += 1;
    }
    accepted = true;
    if (done || s.step > max_step) { s.terminated = 1; return true; }   // go_env.cc:67-73
    return done;
}

// Tromp-Taylor area (board.cc:822-958).  Returns black - white (no komi); bt/wt = points owned by each side
// (stones + empty regions that reach only that colour).  A stone-less board scores 0 with no owner (board.cc:932-935).
template <int N, int T>
__device__ __forceinline__ int area_score(const BB<N>& black, const BB<N>& white, BB<N>& bt, BB<N>& wt) {
    const BB<N> empty = bb_not(black | white);
    // two independent floods: run them in different lanes' registers only when T==1; otherwise sequentially
    BB<N> rb = flood<N, T>(nbrs(black) & empty, empty);
    BB<N> rw = flood<N, T>(nbrs(white) & empty, empty);
    bt = black | andn(rb, rw);
    wt = white | andn(rw, rb);
    return count(bt) - count(wt);
}

// The ten feature planes of encode10 (board_feature.cc:213-223) as bitboards, from the view of the player to move.
template <int N, int T>
__device__ __forceinline__ void feature_planes_g(const Pos<N>& s, BB<N>* pl /*[10]*/, Groups<N>& g) {
    const int me = s.player, op = 3 - s.player;
    analyse_groups<N, T, true>(s.stones[0], s.stones[1], s.player, g);
    const BB<N> own = stones_of(s, me), opp = stones_of(s, op);
    const BB<N> l1me = l1_of(g, me), l1op = l1_of(g, op), l2me = l2_of(g, me), l2op = l2_of(g, op);
    pl[0] = l1me; pl[1] = l2me; pl[2] = andn(andn(own, l1me), l2me);       // getLibertyMap3 :17-41
    pl[3] = l1op; pl[4] = l2op; pl[5] = andn(andn(opp, l1op), l2op);
    pl[6] = (s.last1 >= 0 && s.last1 < Geo<N>::P) ? bb_point<N>(s.last1) : bb_zero<N>();  // getHistoryMap1 :92-100
    pl[7] = suicide_points(own, opp, l1me, l1op);                                           // getSimpleKoAndSuicide :69-89
    if (s.ko_age == 0 && s.ko_loc >= 0 && s.ko_loc < Geo<N>::P) pl[7] = pl[7] | bb_point<N>(s.ko_loc);   // colour ignored (board.cc:205-213)
    pl[8] = true_eyes(own, opp);                                                            // getTrueEyeMap :142-161
    pl[9] = g.lives;                                                                        // getBlockLives :164-182
}
template <int N, int T>
__device__ __forceinline__ void feature_planes(const Pos<N>& s, BB<N>* pl /*[10]*/) {
    Groups<N> g;
    feature_planes_g<N, T>(s, pl, g);
}

// deterministic per-(seed, game, ply) 32-bit draw shared with the CPU checkers (tests mirror it in numpy)
__host__ __device__ __forceinline__ uint32_t mix32(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t z = seed + a * 0x9E3779B97F4A7C15ull + b * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}
// the same draw with the (seed, a) part hoisted out of a loop: mix32(seed, a, b) == mix32_tail(mix32_head(seed, a), b)
__host__ __device__ __forceinline__ uint64_t mix32_head(uint64_t seed, uint64_t a) { return seed + a * 0x9E3779B97F4A7C15ull; }
__host__ __device__ __forceinline__ uint32_t mix32_tail(uint64_t head, uint64_t b) {
    uint64_t z = head + b * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

}  // namespace tg
