// tg_common.cuh -- launch constants and the shared-memory staging used to turn feature bitboards into the
// float32 [10][n][n] planes of board_feature.cc:213-223 with coalesced 128-bit stores.
#pragma once
#include "tg_board.cuh"

namespace tg {

constexpr int kThreads = 128;     // threads per CTA for every tile-cooperative kernel

int set_err(int code, const char* what);
const char* last_err();
int pick_tile(int tile, long long G);

// The CTA keeps ONE contiguous bit string: game `local` of the CTA owns bits [local*10P, (local+1)*10P), plane k
// of that game bits [k*P, (k+1)*P) inside it, in action order -- i.e. exactly the order of the output floats.
template <int N> __host__ __device__ constexpr size_t planes_smem_bytes(int games_per_cta) {
    return (size_t)((games_per_cta * Geo<N>::PLANE_BITS + 31) / 32 + 2) * 4;
}

template <int N> __device__ __forceinline__ void planes_smem_clear(uint32_t* sm, int games_per_cta) {
    const int words = (games_per_cta * Geo<N>::PLANE_BITS + 31) / 32 + 2;
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = 0;
}

template <int N, int T> __device__ __forceinline__ void planes_to_smem(const BB<N>* pl, uint32_t* sm, int local) {
    using G = Geo<N>;
    const int lane = Tile<T>::lane();
    const int base = local * G::PLANE_BITS;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
#pragma unroll
        for (int y = 0; y < N; ++y) {
            if (((k * N + y) & (T - 1)) != lane) continue;
            const int wi = y / G::RPW, sh = (y % G::RPW) * G::STRIDE;
            const uint32_t rowbits = (uint32_t)((pl[k].w[wi] >> sh) & G::row());
            if (!rowbits) continue;
            const int off = base + k * G::P + y * N;
            const int ds = off & 31;
            atomicOr(&sm[off >> 5], rowbits << ds);
            if (ds + N > 32) atomicOr(&sm[(off >> 5) + 1], rowbits >> (32 - ds));
        }
    }
}

// The same for wide tiles (T >= 8), where the loop above spends most of its instructions on the 10 N lane tests: the tile
// first parks its (replicated) bitboards in `stage` (10 * W words of word_t as u32 pairs), then lane l assembles whole 32-bit
// words l, l + T, ... of the game's output string from the <= 32 / N + 2 rows each one overlaps and ORs them into place.
template <int N> __host__ __device__ constexpr int planes_stage_words() { return 10 * Geo<N>::W * (int)(sizeof(typename Geo<N>::word_t) / 4); }

template <int N, int T> __device__ __forceinline__ void planes_to_smem_staged(const BB<N>* pl, uint32_t* sm, int local, uint32_t* stage) {
    using G = Geo<N>;
    constexpr int U = (int)(sizeof(typename G::word_t) / 4);      // u32 per board word
    constexpr int SW = 10 * G::W * U;
    const int lane = Tile<T>::lane();
    if (lane == 0) {                                   // every lane holds the same bitboards: one lane parks them (SW plain stores)
#pragma unroll
        for (int i = 0; i < SW; ++i) {
            const typename G::word_t wv = pl[i / (G::W * U)].w[(i / U) % G::W];
            stage[i] = (uint32_t)(U == 1 ? wv : (wv >> (32 * (i % U))));
        }
    }
    Tile<T>::sync();
    constexpr int NW = (G::PLANE_BITS + 31) / 32;                  // words of one game's string
    const int base = local * G::PLANE_BITS;
    for (int w = lane; w < NW; w += T) {
        uint32_t out = 0;
        const int r0 = (32 * w) / N;
        int r1 = (32 * w + 31) / N;
        r1 = r1 < 10 * N ? r1 : 10 * N - 1;
        for (int r = r0; r <= r1; ++r) {
            const int k = r / N, y = r - k * N;
            const int wi = y / G::RPW, sh = (y - wi * G::RPW) * G::STRIDE;          // bit offset of the row inside its board word
            const uint32_t* bw = stage + (k * G::W + wi) * U;
            uint32_t row;
            if (U == 1) row = bw[0] >> sh;
            else row = sh < 32 ? __funnelshift_r(bw[0], bw[1], sh) : (bw[1] >> (sh - 32));
            row &= (uint32_t)G::row();
            const int off = r * N - 32 * w;                                         // position of the row's bit 0 in this word
            out |= off >= 0 ? (row << off) : (row >> (-off));
        }
        if (out) {
            const int bit = base + 32 * w, ds = bit & 31;
            atomicOr(&sm[bit >> 5], out << ds);
            if (ds) atomicOr(&sm[(bit >> 5) + 1], out >> (32 - ds));
        }
    }
}

// ngames * PLANE_BITS floats are written by `nthreads` cooperating threads (a CTA or one warp), thread `tid` of them.
// `out` is 4-byte aligned at least; up to three leading floats are peeled so that the bulk goes out as aligned
// 128-bit stores.
template <int N> __device__ __forceinline__ void planes_smem_to_global_part(const uint32_t* sm, float* __restrict__ out, int ngames, int tid, int nthreads) {
    const int total = ngames * Geo<N>::PLANE_BITS;
    int head = (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(out) & 15u)) & 15u) >> 2);
    head = head < total ? head : total;
    const int quads = (total - head) >> 2;
    float4* o4 = reinterpret_cast<float4*>(out + head);
    for (int f = tid; f < quads; f += nthreads) {
        const int bit = head + 4 * f;
        const uint32_t nib = __funnelshift_r(sm[bit >> 5], sm[(bit >> 5) + 1], bit & 31) & 15u;
        o4[f] = make_float4((float)(nib & 1u), (float)((nib >> 1) & 1u), (float)((nib >> 2) & 1u), (float)((nib >> 3) & 1u));
    }
    if (tid < head) out[tid] = (float)((sm[0] >> tid) & 1u);
    const int e = head + (quads << 2) + tid;
    if (e < total) out[e] = (float)((sm[e >> 5] >> (e & 31)) & 1u);
}
template <int N> __device__ __forceinline__ void planes_smem_to_global(const uint32_t* sm, float* __restrict__ out, int ngames) {
    planes_smem_to_global_part<N>(sm, out, ngames, (int)threadIdx.x, (int)blockDim.x);
}

}  // namespace tg
