// tg_traj.cu -- the output side of the hot path: trajectory assembly and root noise, on the device.
//
// Replaces the per-game Python loop at the end of SelfPlay.continuous_self_play (self_play.py:931-965: label every
// position of a finished game with the winner z and the ownership map from the mover's view, expand it with the eight
// board symmetries, mem.append each) and np.random.dirichlet (Node_V.dirichlet_prior, self_play.py:90-95).
#include <cstdint>
#include "../../include/transgo_b200.h"
#include "tg_common.cuh"

namespace tg {

// source point of output point (y, x) under symmetry s = 2 * (i - 1) + f  (i = 1..4 quarter turns of np.rot90, f = fliplr of
// that rotation; self_play.py:943-955):   rot90(m, 1)[y][x] = m[x][n-1-y],  fliplr(m)[y][x] = m[y][n-1-x]
__device__ __forceinline__ int sym_source(int s, int y, int x, int n) {
    const int f = s & 1, i = (s >> 1) + 1;
    if (f) x = n - 1 - x;
    int sy, sx;
    switch (i & 3) {
        case 1: sy = x; sx = n - 1 - y; break;
        case 2: sy = n - 1 - y; sx = n - 1 - x; break;
        case 3: sy = n - 1 - x; sx = y; break;
        default: sy = y; sx = x; break;          // four quarter turns
    }
    return sy * n + sx;
}

// One CTA per game.  Finished games (fin[g] != 0, all games when fin == nullptr) hand out length[g] positions each; rows are
// numbered game-major, position-minor, symmetry-innermost -- the order of the reference loop over a batch of games -- and
// written at (ring_index + row) % capacity.  sym = 8: the eight symmetries, sym = 1: the position as played.
__global__ void __launch_bounds__(256)
k_traj_assemble(int n, int C, int G, int T, const float* __restrict__ h_obs, const float* __restrict__ h_pi, const int* __restrict__ h_player,
                const long long* __restrict__ length, const uint8_t* __restrict__ fin, const uint8_t* __restrict__ winner,
                const float* __restrict__ terr, float* __restrict__ o_obs, float* __restrict__ o_pi, float* __restrict__ o_z,
                float* __restrict__ o_own, long long capacity, long long ring_index, int sym, long long* __restrict__ rows_out) {
    const int g = blockIdx.x;
    const int P = n * n, A = P + 1;
    __shared__ long long s_part[256];
    // rows of the finished games before this one (exclusive prefix), by a block reduction
    long long mine = 0, total = 0;
    if (!length && !fin) { if (threadIdx.x == 0) { mine = g; total = G; } }          // one position per game: no scan needed
    else for (int q = threadIdx.x; q < G; q += blockDim.x) {
        const long long len = (!fin || fin[q]) ? (length ? length[q] : 1) : 0;
        total += len;
        if (q < g) mine += len;
    }
    s_part[threadIdx.x] = mine;
    __syncthreads();
    for (int s = blockDim.x / 2; s >= 1; s >>= 1) { if ((int)threadIdx.x < s) s_part[threadIdx.x] += s_part[threadIdx.x + s]; __syncthreads(); }
    const long long before = s_part[0];
    __syncthreads();
    if (g == 0 && rows_out) {                                          // total row count, for the host's ring bookkeeping
        s_part[threadIdx.x] = total;
        __syncthreads();
        for (int s = blockDim.x / 2; s >= 1; s >>= 1) { if ((int)threadIdx.x < s) s_part[threadIdx.x] += s_part[threadIdx.x + s]; __syncthreads(); }
        if (threadIdx.x == 0) *rows_out = s_part[0] * sym;
    }
    if (fin && !fin[g]) return;
    const int len = length ? (int)length[g] : 1;
    const int win = (int)winner[g];
    const int per_row = C * P + A + 1 + P;
    for (int t = 0; t < len; ++t) {
        const float* so = h_obs + ((size_t)g * T + t) * C * P;
        const float* sp = h_pi + ((size_t)g * T + t) * A;
        const int pl = h_player[(size_t)g * T + t];
        const float z = pl == win ? 1.0f : -1.0f;                      // self_play.py:939
        const bool black = pl == kBlack;                               // ownership from the mover's view (:940-941)
        for (int s = 0; s < sym; ++s) {
            const long long row = ((before + t) * sym + s + ring_index) % capacity;
            float* oo = o_obs + (size_t)row * C * P;
            float* op = o_pi + (size_t)row * A;
            float* ow = o_own + (size_t)row * P;
            const int smap = s;
            for (int e = threadIdx.x; e < per_row; e += blockDim.x) {
                if (e < C * P) {
                    const int c = e / P, p = e - c * P, y = p / n, x = p - y * n;
                    const int src = sym == 8 ? sym_source(smap, y, x, n) : p;
                    oo[e] = so[c * P + src];
                } else if (e < C * P + A) {
                    const int a = e - C * P;
                    if (a == P) op[a] = sp[a];                         // the pass probability is appended unchanged (:950)
                    else { const int y = a / n, x = a - y * n; op[a] = sp[sym == 8 ? sym_source(smap, y, x, n) : a]; }
                } else if (e == C * P + A) {
                    o_z[row] = z;
                } else {
                    const int p = e - C * P - A - 1, y = p / n, x = p - y * n;
                    const float tv = terr[(size_t)g * P + (sym == 8 ? sym_source(smap, y, x, n) : p)];
                    ow[p] = black ? tv : -tv;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- Dirichlet root noise
// Philox4x32-10 (Salmon et al.), counter = (k, game, move, draw), key = seed
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {      // uniform in (0, 1), 53 bits
    return ((double)(((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
}

// Gamma(alpha, 1) for alpha < 1: Marsaglia & Tsang (2000) for alpha + 1, times U^(1/alpha)
__device__ double gamma_small(double alpha, uint32_t k, uint32_t game, uint32_t move, uint64_t seed) {
    const double a = alpha + 1.0, d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    uint32_t r[4];
    for (uint32_t draw = 0;; ++draw) {
        philox4x32(k, game, move, draw, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        const double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
        const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);      // Box-Muller
        const double v1 = 1.0 + c * x;
        if (v1 <= 0.0) continue;
        const double v = v1 * v1 * v1;
        philox4x32(k, game, move, draw | 0x80000000u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        const double u = u01(r[0], r[1]), ub = u01(r[2], r[3]);
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v * pow(ub, 1.0 / alpha);
    }
}

// noise[g][0..K_g) ~ Dirichlet(alpha, ..., alpha), zeros beyond K_g; one warp per game
__global__ void k_dirichlet(double* __restrict__ noise, const int* __restrict__ root_k, int G, int A, double alpha, unsigned long long seed,
                            unsigned move, long long game_id0) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= G) return;
    const int K = root_k[g] < A ? root_k[g] : A;
    double sum = 0.0;
    for (int k = lane; k < A; k += 32) {
        double x = 0.0;
        if (k < K) x = gamma_small(alpha, (uint32_t)k, (uint32_t)(game_id0 + g), move, seed);
        noise[(size_t)g * A + k] = x;
        sum += x;
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    const double inv = 1.0 / (sum > 1e-300 ? sum : 1e-300);
    __syncwarp();
    for (int k = lane; k < K; k += 32) noise[(size_t)g * A + k] *= inv;
}

}  // namespace tg

using namespace tg;

extern "C" {

int tg_traj_assemble(int n, int channels, int64_t G, int64_t T, const float* hist_obs, const float* hist_pi, const int32_t* hist_player,
                     const int64_t* length, const uint8_t* fin, const uint8_t* winner, const float* territory, float* out_obs,
                     float* out_pi, float* out_z, float* out_own, int64_t capacity, int64_t ring_index, int symmetries, int64_t* rows_out,
                     void* stream) {
    if (!hist_obs || !hist_pi || !hist_player || !winner || !territory || !out_obs || !out_pi || !out_z || !out_own ||
        G <= 0 || T <= 0 || capacity <= 0 || (symmetries != 1 && symmetries != 8) || n < 2)
        return set_err(-1, "tg_traj_assemble");
    k_traj_assemble<<<(unsigned)G, 256, 0, (cudaStream_t)stream>>>(n, channels, (int)G, (int)T, hist_obs, hist_pi, hist_player,
                                                                   (const long long*)length, fin, winner, territory, out_obs, out_pi,
                                                                   out_z, out_own, capacity, ring_index % capacity, symmetries,
                                                                   (long long*)rows_out);
    return set_err((int)cudaGetLastError(), "tg_traj_assemble");
}

int tg_dirichlet(double* noise, const int32_t* root_k, int64_t G, int A, double alpha, uint64_t seed, uint32_t move, int64_t game_id0,
                 void* stream) {
    if (!noise || !root_k || G <= 0 || A <= 0 || !(alpha > 0.0) || alpha >= 1.0) return set_err(-1, "tg_dirichlet (0 < alpha < 1)");
    k_dirichlet<<<(unsigned)((G * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(noise, root_k, (int)G, A, alpha, seed, move, game_id0);
    return set_err((int)cudaGetLastError(), "tg_dirichlet");
}

}  // extern "C"
