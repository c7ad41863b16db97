// tg_env.cu -- batched board kernels + the C ABI of include/transgo_b200.h (environment half).
// Replaces GoEnv/cpp_src/go_env.cc (reference C ABI) and the loops of board.cc / board_feature.cc behind it.
#include <cstdio>
#include <cstring>
#include <mutex>
#include "../../include/transgo_b200.h"
#include "tg_board.cuh"
#include "tg_common.cuh"

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
    Groups<N> gr;
    analyse_groups<N, T, false>(s.stones[0], s.stones[1], s.player, gr);
    bool acc;
    bool done = apply_move<N, T>(s, gr, (int)actions[g], max_step, acc);
    BB<N> lg = bb_zero<N>();
    if (legal_out) {                          // uniform
        analyse_groups<N, T, false>(s.stones[0], s.stones[1], s.player, gr);
        lg = legal_points(s, gr);
    }
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
    Groups<N> gr;
    analyse_groups<N, T, false>(s.stones[0], s.stones[1], s.player, gr);
    BB<N> lg = legal_points(s, gr);
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
        planes_to_smem<N, T>(pl, sm_bits, local);
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
            Groups<N> gr;
            analyse_groups<N, T, false>(s.stones[0], s.stones[1], s.player, gr);
            BB<N> lg = legal_points(s, gr);
            if (no_eye) lg = andn(lg, true_eyes(stones_of(s, s.player), stones_of(s, 3 - s.player)));
            const int nl = count(lg);
            if (nl == 0) action = GG::P;
            else {
                uint32_t r = mix32(seed, (unsigned long long)(game_id0 + g), (unsigned long long)s.step);
                action = bb_nth<N>(lg, (int)(((unsigned long long)r * (unsigned)nl) >> 32));
            }
            bool acc;
            apply_move<N, T>(s, gr, action, max_step, acc, true);
            played += 1;
        }
        if (trace && Tile<T>::lane() == 0) trace[(long long)t * G + g] = (int16_t)action;
    }
    if (Tile<T>::lane() == 0) {
        pos_store<N>(s, states, g, G, 1);
        if (plies_done && played) atomicAdd(plies_done, (unsigned long long)played);
    }
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
    const size_t smem = planes_smem_bytes<N>(GPC);
    static bool attr_done = false;
    if (!attr_done && smem > 48 * 1024) {
        cudaFuncSetAttribute(k_encode<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    k_encode<N, T><<<(unsigned)((G + GPC - 1) / GPC), kThreads, smem, st>>>(states, G, planes);
}
template <int N, int T> void launch_score(cudaStream_t st, const uint4* states, long long G, float komi, float* score, float* territory) {
    k_score<N, T><<<blocks_for(G, T), kThreads, 0, st>>>(states, G, komi, score, territory);
}
template <int N, int T> void launch_playout(cudaStream_t st, uint4* states, long long G, int plies, int max_step, unsigned long long seed,
                                            long long game_id0, int no_eye, int16_t* trace, unsigned long long* plies_done) {
    k_playout<N, T><<<blocks_for(G, T), kThreads, 0, st>>>(states, G, plies, max_step, seed, game_id0, no_eye, trace, plies_done);
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

int tg_env_playout(int n, void* states, int64_t G, int plies, int max_step, uint64_t seed, int64_t game_id0, int no_eye,
                   int16_t* trace_actions, unsigned long long* plies_done, int tile, void* stream) {
    if (!states || G <= 0 || plies < 0) return set_err(-1, "tg_env_playout");
    cudaStream_t st = (cudaStream_t)stream;
    tile = snap_tile(tile, G);
    TG_DISPATCH(n, tile, launch_playout, st, (uint4*)states, G, plies, max_step, seed, game_id0, no_eye, trace_actions, plies_done);
    return set_err((int)cudaGetLastError(), "tg_env_playout");
}

int tg_sub_encode(int n, const float* enc, int64_t B, int channels, int sub, int cuts, float* out, void* stream) {
    if (!enc || !out || B <= 0 || sub <= 0 || sub > n || cuts < 1 || cuts > 5) return set_err(-1, "tg_sub_encode");
    long long total = (long long)B * cuts * channels * sub * sub;
    unsigned nb = (unsigned)((total + 255) / 256); if (nb > 148 * 16) nb = 148 * 16;
    k_sub_encode<<<nb, 256, 0, (cudaStream_t)stream>>>(enc, B, n, channels, sub, cuts, out);
    return set_err((int)cudaGetLastError(), "tg_sub_encode");
}

}  // extern "C"
