// tg_net_tc.cu -- the self-attention block of the TransGo trunk (model.py:286-314) as ONE tcgen05 kernel for 9x9 boards:
// the q | k | v projection (three 1x1 convolutions = one [P,128] x [128,192] GEMM), the scores, the softmax, att^T . v and the
// residual junction (gamma * o + x, BatchNorm, ReLU) for one position per pass, with every contraction on the 5th-generation
// tensor cores (tcgen05.mma, operands in shared memory, accumulators in tensor memory) -- q, k, v and the scores never touch HBM.
// (tg_net.cu keeps the mma.sync kernels: they are the cross-check, and the path for boards with more than 96 points.)
//
// One CTA = 256 threads: the 128 TMEM lanes = 128 operand rows (81 points padded), TWO threads per row that split the columns
// (warps w and w + 4 see the same 32 lanes); two CTAs per SM (2 x 256 TMEM columns, 2 x 109 KB shared memory), each looping over
// positions with the next position's x tile fetched into registers a position ahead.  Per position:
//   1. x tile  [128 x 128] bf16 -> shared memory (K-major panels)              D1[i][n] = sum_c x[i][c] w[n][c]     8 MMAs, N = 192
//   2. the two threads of row i read it from TMEM (tcgen05.ld 32x32b), add the bias, round to bf16 and park q / k (K-major);
//      the score MMA is issued at once                                          S[i][j]  = sum_d q[i][d] k[j][d]     2 MMAs, N = 96
//      and runs while the same threads park v (MN-major)
//   3. softmax over the keys j of row i: each thread of the pair owns 48 columns, one exchange of (max, sum) through shared
//      memory; p[i][:] -> shared memory as the MN-major A operand p^T           O[j][c]  = sum_i p[i][j] v[i][c]     6 MMAs, N = 128
//   4. row j of O -> bf16 -> shared memory (over the dead v panels); then the junction as a COALESCED pass over the tile: 16 threads
//      per row, x from global memory (an L2 hit, requested before the MMA wait), relu(bn(gamma * o + x)), 16-byte stores -- run while
//      the tensor core already projects the NEXT position; optionally a second output relu(bn1(out)) for the residual block that follows.
// TMEM columns: D1 = v [0,128) | q [128,160) | k [160,192) (the weight rows are permuted on the way into shared memory),
// S [128,224) over the parked q | k, O [0,128) over the parked v.
// Shared-memory operand layout ("panels"): the canonical no-swizzle layout of the UMMA descriptors with all rows of a
// 16-byte column chunk contiguous --  byte(row r, chunk q) = q * PANEL + r * 16.  For a K-major operand (chunks along K) that
// is SBO = 128 (8-row core-matrix groups), LBO = PANEL; for an MN-major operand (chunks along M / N, rows = K index) it is
// SBO = PANEL, LBO = 128.  A thread that owns one row writes 16-byte chunks, a warp writes 512 contiguous bytes per chunk.
// PANEL = rows * 16 + 16 so that consecutive chunks start 4 banks apart (the global -> shared copy of x walks the chunks).
#include <cuda_bf16.h>
#include <cstdint>
#include "../../include/transgo_b200.h"
#include "tg_common.cuh"

namespace tg {
namespace tc {

constexpr int kM = 128;                 // MMA M = TMEM lanes = threads
constexpr int kPts = 96;                // points padded to a multiple of 16 (N of the scores, K of att^T . v)
constexpr int kC = 128, kDq = 32, kNqkv = kC + 2 * kDq;
constexpr int kTcThreads = 256;
constexpr int PX = kM * 16 + 16;        // x:    16 channel chunks x 128 rows
constexpr int PW = kNqkv * 16 + 16;     // w:    16 channel chunks x 192 rows (n)
constexpr int PQ = kM * 16 + 16;        // q:     4 chunks x 128 rows (i)
constexpr int PK = kPts * 16 + 16;      // k:     4 chunks x 96 rows (j)
constexpr int PV = kPts * 16 + 16;      // v:    16 channel chunks x 96 rows (i)          MN-major B
constexpr int PP = kPts * 16 + 16;      // p^T:  16 j chunks (12 written) x 96 rows (i)    MN-major A
constexpr int OFF_W = 0;
constexpr int OFF_BIAS = OFF_W + 16 * PW;
constexpr int OFF_SCALE = OFF_BIAS + kNqkv * 4;
constexpr int OFF_SHIFT = OFF_SCALE + kC * 4;
constexpr int OFF_R1 = OFF_SHIFT + kC * 4;          // x, then q | k, then p^T (each dead before the next is written)
constexpr int R1_BYTES = 16 * PX;
constexpr int OFF_Q = OFF_R1, OFF_K = OFF_R1 + 4 * PQ, OFF_P = OFF_R1;
constexpr int OFF_V = OFF_R1 + R1_BYTES;
constexpr int OFF_SM = OFF_V + 16 * PV;             // softmax exchange: float2 (max, sum) [2][128]
constexpr int OFF_BAR = OFF_SM + 2 * kM * 8;
constexpr int SMEM_BYTES = OFF_BAR + 16;
static_assert(4 * PQ + 4 * PK <= R1_BYTES && 16 * PP <= R1_BYTES, "q | k and p^T live in the x tile");
static_assert(SMEM_BYTES * 2 + 2048 <= 227 * 1024, "two CTAs per SM");
constexpr int kTmemCols = 256;
constexpr int kColV = 0, kColQ = 128, kColK = 160, kColS = 128, kColO = 0;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// UMMA shared-memory descriptor, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor of kind::f16 (cute::UMMA::InstrDescriptor): f32 accumulate, bf16 x bf16, M = 128
__host__ __device__ constexpr uint32_t instr_desc(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(kM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// a bounded wait: a descriptor mistake must end in a trap (an error the host sees), never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (true) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }

__device__ __forceinline__ void load_x_tile(const __nv_bfloat16* __restrict__ xp, int P, int tid, uint4* v) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int i = (tid >> 4) + 16 * it, q = tid & 15;
        v[it] = i < P ? __ldg(reinterpret_cast<const uint4*>(xp + (size_t)i * kC + q * 8)) : make_uint4(0, 0, 0, 0);
    }
}
// 32 accumulator columns + bias -> four 16-byte bf16 chunks in a panel (chunk stride `pstride`); zero = write zeros instead
__device__ __forceinline__ void park32(const uint32_t* r, const float* __restrict__ bias32, unsigned char* dst, int pstride, bool zero,
                                       float* __restrict__ dump) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias32 + 8 * q), b1 = *reinterpret_cast<const float4*>(bias32 + 8 * q + 4);
        float f[8];
        f[0] = __uint_as_float(r[8 * q]) + b0.x; f[1] = __uint_as_float(r[8 * q + 1]) + b0.y; f[2] = __uint_as_float(r[8 * q + 2]) + b0.z;
        f[3] = __uint_as_float(r[8 * q + 3]) + b0.w; f[4] = __uint_as_float(r[8 * q + 4]) + b1.x; f[5] = __uint_as_float(r[8 * q + 5]) + b1.y;
        f[6] = __uint_as_float(r[8 * q + 6]) + b1.z; f[7] = __uint_as_float(r[8 * q + 7]) + b1.w;
        if (dump) {
#pragma unroll
            for (int e = 0; e < 8; ++e) dump[8 * q + e] = f[e];
        }
        uint4 o = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        if (zero) o = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dst + q * pstride) = o;
    }
}

// softmax of one row, columns [48 H, 48 H + 48) (H is a template parameter so that the key mask j < P folds at compile time when
// P is a constant); the two halves of a row meet through `sm` (one __syncthreads inside: every thread of the CTA calls this)
template <int H>
__device__ __forceinline__ void softmax_half(uint32_t tlane, int row, int P, float2* __restrict__ sm, unsigned char* __restrict__ smem,
                                             float* __restrict__ dump) {
    constexpr int HW = kPts / 2, j0 = HW * H;
    float s[HW];
    {
        uint32_t r[32];
        tmem_ld32(tlane + kColS + j0, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) s[e] = __uint_as_float(r[e]);
        tmem_ld16(tlane + kColS + j0 + 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) s[32 + e] = __uint_as_float(r[e]);
    }
    if (dump) {
#pragma unroll
        for (int e = 0; e < HW; ++e) dump[row * kPts + j0 + e] = s[e];
    }
    float m4[4] = {-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};           // four independent chains
#pragma unroll
    for (int e = 0; e < HW; ++e) m4[e & 3] = (j0 + e < P) ? fmaxf(m4[e & 3], s[e]) : m4[e & 3];
    const float mloc = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    const float mb = mloc * kLog2e;
    float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int e = 0; e < HW; ++e) { s[e] = (j0 + e < P) ? ex2_fast(fmaf(s[e], kLog2e, -mb)) : 0.0f; z4[e & 3] += s[e]; }
    const float zloc = (z4[0] + z4[1]) + (z4[2] + z4[3]);
    sm[H * kM + row] = make_float2(mloc, zloc);
    __syncthreads();
    const float2 other = sm[(H ^ 1) * kM + row];
    const float mall = fmaxf(mloc, other.x);                          // exp(s - m) = exp(s - m_half) * exp(m_half - m)
    const float fme = ex2_fast((mloc - mall) * kLog2e), fot = ex2_fast((other.x - mall) * kLog2e);
    const float z = zloc * fme + other.y * fot;
    const float inv = (row < P) ? fme / z : 0.0f;                     // rows >= P do not exist: p = 0
    if (row < kPts) {
#pragma unroll
        for (int q = 0; q < HW / 8; ++q) {
            uint4 o;
            o.x = pack_bf16(s[8 * q] * inv, s[8 * q + 1] * inv); o.y = pack_bf16(s[8 * q + 2] * inv, s[8 * q + 3] * inv);
            o.z = pack_bf16(s[8 * q + 4] * inv, s[8 * q + 5] * inv); o.w = pack_bf16(s[8 * q + 6] * inv, s[8 * q + 7] * inv);
            *reinterpret_cast<uint4*>(smem + OFF_P + (HW / 8 * H + q) * PP + row * 16) = o;
        }
    }
}

// the junction of one position, coalesced: relu(bn(gamma * o + x)) (model.py:312-313) with o (bf16) from the staging panels;
// thread -> chunk tid & 15 of rows (tid >> 4) + 16 it
// og2 (may be null): a second output relu(out * scale2 + shift2) -- the BatchNorm + ReLU that opens the NEXT residual block
// (model.py:216-217), applied to the bf16-rounded out like the separate junction kernel would; its scale / shift come from shared memory
__device__ __forceinline__ void junction_pass(const unsigned char* __restrict__ smem, const uint4* xj, uint4* __restrict__ og, int tile_chunks,
                                              int tid, float gamma, const float* sc, const float* sh, uint4* __restrict__ og2) {
#pragma unroll
    for (int it = 0; it < 6; ++it) {
        const int idx = tid + kTcThreads * it;
        if (idx < tile_chunks) {
            const uint4 ov = *reinterpret_cast<const uint4*>(smem + OFF_V + (tid & 15) * PV + (idx >> 4) * 16);
            const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, xw[4] = {xj[it].x, xj[it].y, xj[it].z, xj[it].w};
            uint32_t res[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 of = unpack_bf16(ow[k]), xf = unpack_bf16(xw[k]);
                const float s0 = fmaf(gamma, of.x, xf.x), s1 = fmaf(gamma, of.y, xf.y);
                res[k] = pack_bf16(fmaxf(fmaf(s0, sc[2 * k], sh[2 * k]), 0.0f), fmaxf(fmaf(s1, sc[2 * k + 1], sh[2 * k + 1]), 0.0f));
            }
            og[idx] = make_uint4(res[0], res[1], res[2], res[3]);
            if (og2) {
                const float* s2 = reinterpret_cast<const float*>(smem + OFF_SCALE) + (tid & 15) * 8;
                const float* t2 = reinterpret_cast<const float*>(smem + OFF_SHIFT) + (tid & 15) * 8;
                const float4 sa = *reinterpret_cast<const float4*>(s2), sb = *reinterpret_cast<const float4*>(s2 + 4);
                const float4 ta = *reinterpret_cast<const float4*>(t2), tb = *reinterpret_cast<const float4*>(t2 + 4);
                const float s2v[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w}, t2v[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
                uint32_t r2[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 of = unpack_bf16(res[k]);
                    // multiply and add rounded separately, exactly like tg_net_add_bn_relu (its file is built with -fmad=false): the fused
                    // and the two-kernel route give the same bits
                    r2[k] = pack_bf16(fmaxf(__fadd_rn(__fmul_rn(of.x, s2v[2 * k]), t2v[2 * k]), 0.0f),
                                      fmaxf(__fadd_rn(__fmul_rn(of.y, s2v[2 * k + 1]), t2v[2 * k + 1]), 0.0f));
                }
                og2[idx] = make_uint4(r2[0], r2[1], r2[2], r2[3]);
            }
        }
    }
}

// dbg (may be null): position 0 dumps D1 + bias [128][192] (q | k | v order), S [128][96], O [128][128] as float32 (tests check
// every stage).  PC: the point count as a compile-time constant (81: the 9x9 board), 0 = run-time P
template <int PC>
__global__ void __launch_bounds__(kTcThreads, 2)
k_attention_tc(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ bias,
               float gamma, const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
               int B, int Prt, float* __restrict__ dbg, const float* __restrict__ scale2, const float* __restrict__ shift2,
               __nv_bfloat16* __restrict__ out2) {
    const int P = PC ? PC : Prt;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (kM - 1), h = tid >> 7;                     // this thread's operand row / TMEM lane and its column half
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + OFF_BAR;
    float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);        // in TMEM column order: v | q | k
    float2* s_sm = reinterpret_cast<float2*>(smem + OFF_SM);

    // ---- one-time setup: TMEM columns, the barrier, the weights as K-major panels (rows permuted to v | q | k), the bias
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint4 xt[8];                                                      // the NEXT position's x tile, fetched a position ahead
    if ((int)blockIdx.x < B) load_x_tile(x + (size_t)blockIdx.x * P * kC, P, tid, xt);      // (the first one: under the weight staging)
    {
        constexpr int WIT = kNqkv * 16 / kTcThreads;                  // 12 chunks per thread, all loads in flight before the first store
        uint4 wv[WIT];
#pragma unroll
        for (int it = 0; it < WIT; ++it) wv[it] = __ldg(reinterpret_cast<const uint4*>(w) + tid + kTcThreads * it);
#pragma unroll
        for (int it = 0; it < WIT; ++it) {
            const int idx = tid + kTcThreads * it, n = idx >> 4, q = idx & 15;
            const int np = n < 2 * kDq ? n + kC : n - 2 * kDq;        // q | k | v  ->  v | q | k
            *reinterpret_cast<uint4*>(smem + OFF_W + q * PW + np * 16) = wv[it];
        }
    }
    for (int n = tid; n < kNqkv; n += kTcThreads) s_bias[n < 2 * kDq ? n + kC : n - 2 * kDq] = __bfloat162float(bias[n]);
    if (out2) {
        for (int c = tid; c < kC; c += kTcThreads) {
            reinterpret_cast<float*>(smem + OFF_SCALE)[c] = scale2[c];
            reinterpret_cast<float*>(smem + OFF_SHIFT)[c] = shift2[c];
        }
    }
    // the junction pass gives a thread the same 8 channels of every row it touches: their scale / shift live in registers
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = scale[(tid & 15) * 8 + e]; sh[e] = shift[(tid & 15) * 8 + e]; }
    fence_async_smem();
    tc_before_sync();
    __syncthreads();
    tc_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t phase = 0;
    constexpr uint32_t ID_PROJ = instr_desc(kNqkv, 0, 0), ID_S = instr_desc(kPts, 0, 0), ID_O = instr_desc(kC, 1, 1);
    const int tile_chunks = P * 16;                                   // 16-byte chunks of one position's [P][128] bf16 tile

    uint4 xj[6];                                                      // x chunks of the pending junction pass
    uint4* og_prev = nullptr;                                         // where the pending junction pass writes (null: none pending)
    uint4* og2_prev = nullptr;
    for (int pos = blockIdx.x; pos < B; pos += gridDim.x) {
        const bool dump = dbg != nullptr && pos == 0;
        // ---- 1. the x tile (rows >= P zero) from the registers that were filled during the previous position
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int i = (tid >> 4) + 16 * it, q = tid & 15;
            *reinterpret_cast<uint4*>(smem + OFF_R1 + q * PX + i * 16) = xt[it];
        }
        fence_async_smem();
        tc_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_after_sync();
#pragma unroll
            for (int k = 0; k < kC / 16; ++k)
                umma(tmem, smem_desc(sbase + OFF_R1 + 2 * k * PX, PX, 128), smem_desc(sbase + OFF_W + 2 * k * PW, PW, 128), ID_PROJ, k > 0);
            umma_commit(bar);
        }
        // ---- 4b (of the PREVIOUS position). its junction pass runs while the tensor core projects this one; the barrier above
        // also published its O staging panels
        if (og_prev) junction_pass(smem, xj, og_prev, tile_chunks, tid, gamma, sc, sh, og2_prev);
        mbar_wait(bar, phase); phase ^= 1;
        tc_after_sync();
        // ---- 2a. q (h = 0) / k (h = 1) of this row: + bias, bf16, K-major panels over the dead x tile; then the score MMA at once
        {
            uint32_t r[32];
            tmem_ld32(tlane + (h ? kColK : kColQ), r);
            tmem_ld_wait();
            if (h == 0) park32(r, s_bias + kColQ, smem + OFF_Q + row * 16, PQ, false, dump ? dbg + row * kNqkv : nullptr);
            else if (row < kPts) park32(r, s_bias + kColK, smem + OFF_K + row * 16, PK, false, dump ? dbg + row * kNqkv + kDq : nullptr);
        }
        fence_async_smem();
        tc_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_after_sync();
#pragma unroll
            for (int k = 0; k < kDq / 16; ++k)
                umma(tmem + kColS, smem_desc(sbase + OFF_Q + 2 * k * PQ, PQ, 128), smem_desc(sbase + OFF_K + 2 * k * PK, PK, 128), ID_S, k > 0);
            umma_commit(bar);
        }
        // ---- 2b. meanwhile v: channels [64h, 64h + 64) of this row as MN-major panels (rows >= P zero: they must add nothing to o)
        if (row < kPts) {
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t r[32];
                const int c0 = 64 * h + 32 * c2;
                tmem_ld32(tlane + kColV + c0, r);
                tmem_ld_wait();
                park32(r, s_bias + kColV + c0, smem + OFF_V + (c0 >> 3) * PV + row * 16, PV, row >= P, dump ? dbg + row * kNqkv + 2 * kDq + c0 : nullptr);
            }
        }
        mbar_wait(bar, phase); phase ^= 1;
        tc_after_sync();
        // ---- 3. softmax over the keys j of this row (model.py:304-305): this thread owns j in [48h, 48h + 48)
        // (the next position's tile is requested first: it has the rest of this position to arrive)
        if (pos + (int)gridDim.x < B) load_x_tile(x + (size_t)(pos + gridDim.x) * P * kC, P, tid, xt);
        if (h == 0) softmax_half<0>(tlane, row, P, s_sm, smem, dump ? dbg + kM * kNqkv : nullptr);
        else softmax_half<1>(tlane, row, P, s_sm, smem, dump ? dbg + kM * kNqkv : nullptr);
        fence_async_smem();
        tc_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_after_sync();
#pragma unroll
            for (int k = 0; k < kPts / 16; ++k)
                umma(tmem + kColO, smem_desc(sbase + OFF_P + k * 256, 128, PP), smem_desc(sbase + OFF_V + k * 256, 128, PV), ID_O, k > 0);
            umma_commit(bar);
        }
        // while the tensor core works: the x chunks of the junction pass (coalesced; an L2 hit, the tile was read a position ago)
        const uint4* xg = reinterpret_cast<const uint4*>(x + (size_t)pos * P * kC);
#pragma unroll
        for (int it = 0; it < 6; ++it) { const int idx = tid + kTcThreads * it; xj[it] = idx < tile_chunks ? __ldg(xg + idx) : make_uint4(0, 0, 0, 0); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_after_sync();
        // ---- 4a. channels [64h, 64h + 64) of row `row` of O -> bf16 -> the dead v panels (chunk q, row j at q * PV + j * 16)
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
            uint32_t r[32];
            const int c0 = 64 * h + 32 * c2;
            tmem_ld32(tlane + kColO + c0, r);
            tmem_ld_wait();
            if (dump) {
#pragma unroll
                for (int e = 0; e < 32; ++e) dbg[kM * kNqkv + kM * kPts + row * kC + c0 + e] = __uint_as_float(r[e]);
            }
            if (row < kPts) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(smem + OFF_V + ((c0 >> 3) + q) * PV + row * 16) =
                        make_uint4(pack_bf16(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])), pack_bf16(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                                   pack_bf16(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])), pack_bf16(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
            }
        }
        tc_before_sync();                         // these tcgen05.ld precede the next position's MMAs into the same columns
        og_prev = reinterpret_cast<uint4*>(out + (size_t)pos * P * kC);
        og2_prev = out2 ? reinterpret_cast<uint4*>(out2 + (size_t)pos * P * kC) : nullptr;
        // (the staging panels are read after the next barrier; the next v panels are written two barriers after that)
    }
    __syncthreads();
    if (og_prev) junction_pass(smem, xj, og_prev, tile_chunks, tid, gamma, sc, sh, og2_prev);
    tc_before_sync();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace tc
}  // namespace tg

extern "C" int tg_net_attention_fused2(const void* x, const void* wqkv, const void* bqkv, float gamma, const float* scale, const float* shift,
                                       void* out, const float* scale2, const float* shift2, void* out2, int64_t B, int P, int dq, int channels,
                                       float* debug, void* stream) {
    using namespace tg;
    using namespace tg::tc;
    if (!x || !wqkv || !bqkv || !scale || !shift || !out || B <= 0 || P <= 0 || P > kPts || dq != kDq || channels != kC ||
        (out2 && (!scale2 || !shift2)))
        return set_err(-1, "tg_net_attention_fused (supports P <= 96, 32 query channels, 128 value channels)");
    static unsigned long long attr_done = 0;                 // one bit per device
    int dev = 0; cudaGetDevice(&dev);
    if (!((attr_done >> (dev & 63)) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k_attention_tc<81>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attention_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return set_err((int)e, "tg_net_attention_fused (shared memory attribute)");
        attr_done |= 1ull << (dev & 63);
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = B < 2LL * sms ? B : 2LL * sms;    // persistent: two CTAs per SM, each walks positions blockIdx, + grid, ...
    auto kern = P == 81 ? k_attention_tc<81> : k_attention_tc<0>;
    kern<<<(unsigned)grid, kTcThreads, SMEM_BYTES, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)wqkv, (const __nv_bfloat16*)bqkv, gamma, scale, shift, (__nv_bfloat16*)out, (int)B, P, debug,
        scale2, shift2, (__nv_bfloat16*)out2);
    return set_err((int)cudaGetLastError(), "tg_net_attention_fused");
}

extern "C" int tg_net_attention_fused(const void* x, const void* wqkv, const void* bqkv, float gamma, const float* scale, const float* shift,
                                      void* out, int64_t B, int P, int dq, int channels, float* debug, void* stream) {
    return tg_net_attention_fused2(x, wqkv, bqkv, gamma, scale, shift, out, nullptr, nullptr, nullptr, B, P, dq, channels, debug, stream);
}
