// tg_net_tc.cu -- the self-attention block of the TransGo trunk (model.py:286-314) as ONE tcgen05 kernel for 9x9 boards:
// the q | k | v projection (three 1x1 convolutions = one [P,128] x [128,192] GEMM), the scores, the softmax, att^T . v and the
// residual junction (gamma * o + x, BatchNorm, ReLU) for one position per pass, with every contraction on the 5th-generation
// tensor cores (tcgen05.mma, operands in shared memory, accumulators in tensor memory) -- q, k, v and the scores never touch HBM.
// (tg_net.cu keeps the mma.sync kernels: they are the cross-check, and the path for boards with more than 96 points.)
//
// One CTA = 128 threads = the 128 TMEM lanes = 128 operand rows (81 points padded); two CTAs per SM (2 x 256 TMEM columns,
// 2 x 107 KB shared memory), each looping over positions.  Per position:
//   1. x tile  [128 x 128] bf16 -> shared memory (K-major panels)              D1[i][n] = sum_c x[i][c] w[n][c]     8 MMAs, N = 192
//   2. thread i reads row i of D1 (tcgen05.ld 32x32b), adds the bias, rounds to bf16 and parks q / k (K-major) and v (MN-major)
//                                                                               S[i][j]  = sum_d q[i][d] k[j][d]     2 MMAs, N = 96
//   3. thread i owns row i of S: softmax over j in registers, p[i][:] -> shared memory as the MN-major A operand p^T
//                                                                               O[j][c]  = sum_i p[i][j] v[i][c]     6 MMAs, N = 128
//   4. thread j reads row j of O, junction epilogue against x[j][:] from global memory, 16-byte stores.
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
constexpr int kTcThreads = 128;
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
constexpr int OFF_BAR = OFF_V + 16 * PV;
constexpr int SMEM_BYTES = OFF_BAR + 16;
static_assert(4 * PQ + 4 * PK <= R1_BYTES && 16 * PP <= R1_BYTES, "q | k and p^T live in the x tile");
static_assert(SMEM_BYTES * 2 + 2048 <= 227 * 1024, "two CTAs per SM");
constexpr int kTmemCols = 256;          // D1: [0,192)   S: [0,96) after D1 is parked   O: [96,224)
constexpr int kColS = 0, kColO = 96;
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }

// dbg (may be null): position 0 dumps D1 + bias [128][192], S [128][96], O [128][128] as float32 (tests check every stage)
__global__ void __launch_bounds__(kTcThreads, 2)
k_attention_tc(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ bias,
               float gamma, const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
               int B, int P, float* __restrict__ dbg) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + OFF_BAR;
    float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
    float* s_scale = reinterpret_cast<float*>(smem + OFF_SCALE);
    float* s_shift = reinterpret_cast<float*>(smem + OFF_SHIFT);

    // ---- one-time setup: TMEM columns, the barrier, the weights as K-major panels, bias / scale / shift
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int idx = tid; idx < kNqkv * 16; idx += kTcThreads) {
        const int n = idx >> 4, q = idx & 15;
        const uint4 v = *reinterpret_cast<const uint4*>(w + (size_t)n * kC + q * 8);
        *reinterpret_cast<uint4*>(smem + OFF_W + q * PW + n * 16) = v;
    }
    for (int i = tid; i < kNqkv; i += kTcThreads) s_bias[i] = __bfloat162float(bias[i]);
    for (int i = tid; i < kC; i += kTcThreads) { s_scale[i] = scale[i]; s_shift[i] = shift[i]; }
    fence_async_smem();
    tc_before_sync();
    __syncthreads();
    tc_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's 32 lanes; thread tid <-> lane tid <-> row tid
    uint32_t phase = 0;
    constexpr uint32_t ID_PROJ = instr_desc(kNqkv, 0, 0), ID_S = instr_desc(kPts, 0, 0), ID_O = instr_desc(kC, 1, 1);

    for (int pos = blockIdx.x; pos < B; pos += gridDim.x) {
        const __nv_bfloat16* xp = x + (size_t)pos * P * kC;
        // ---- 1. the x tile, rows >= P zero; all 16 loads of a thread in flight before the first store
        {
            uint4 v[16];
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int i = (tid >> 4) + 8 * it, q = tid & 15;
                v[it] = i < P ? __ldg(reinterpret_cast<const uint4*>(xp + (size_t)i * kC + q * 8)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int i = (tid >> 4) + 8 * it, q = tid & 15;
                *reinterpret_cast<uint4*>(smem + OFF_R1 + q * PX + i * 16) = v[it];
            }
        }
        fence_async_smem();
        tc_before_sync();                         // the previous position's tcgen05.ld of O precede the MMAs below
        __syncthreads();
        if (tid == 0) {
            tc_after_sync();
#pragma unroll
            for (int k = 0; k < kC / 16; ++k)
                umma(tmem, smem_desc(sbase + OFF_R1 + 2 * k * PX, PX, 128), smem_desc(sbase + OFF_W + 2 * k * PW, PW, 128), ID_PROJ, k > 0);
            umma_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1;
        tc_after_sync();
        // ---- 2. row tid of D1: + bias, bf16, q / k as K-major panels, v as MN-major panels (rows >= P zero)
#pragma unroll 1
        for (int ch = 0; ch < kNqkv / 32; ++ch) {
            uint32_t r[32];
            tmem_ld32(tlane + ch * 32, r);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(r[e]) + s_bias[ch * 32 + e];
            if (dbg && pos == 0) {
#pragma unroll
                for (int e = 0; e < 32; ++e) dbg[tid * kNqkv + ch * 32 + e] = f[e];
            }
            const bool live = ch == 0 || tid < kPts;
            const bool zero = ch >= 2 && tid >= P;
            unsigned char* dst = ch == 0 ? smem + OFF_Q + tid * 16 : ch == 1 ? smem + OFF_K + tid * 16 : smem + OFF_V + (ch - 2) * 4 * PV + tid * 16;
            const int pstride = ch == 0 ? PQ : ch == 1 ? PK : PV;
            if (live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 o;
                    o.x = pack_bf16(f[8 * q], f[8 * q + 1]); o.y = pack_bf16(f[8 * q + 2], f[8 * q + 3]);
                    o.z = pack_bf16(f[8 * q + 4], f[8 * q + 5]); o.w = pack_bf16(f[8 * q + 6], f[8 * q + 7]);
                    if (zero) o = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(dst + q * pstride) = o;
                }
            }
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
        mbar_wait(bar, phase); phase ^= 1;
        tc_after_sync();
        // ---- 3. softmax over the keys j of row tid (model.py:304-305), p^T as the MN-major A operand
        {
            float s[kPts];
#pragma unroll
            for (int ch = 0; ch < kPts / 32; ++ch) {
                uint32_t r[32];
                tmem_ld32(tlane + kColS + ch * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) s[ch * 32 + e] = __uint_as_float(r[e]);
            }
            if (dbg && pos == 0) {
#pragma unroll
                for (int e = 0; e < kPts; ++e) dbg[kM * kNqkv + tid * kPts + e] = s[e];
            }
            float m = -3.0e38f;
#pragma unroll
            for (int e = 0; e < kPts; ++e) m = (e < P) ? fmaxf(m, s[e]) : m;
            const float mb = m * kLog2e;
            float z = 0.0f;
#pragma unroll
            for (int e = 0; e < kPts; ++e) { s[e] = (e < P) ? ex2_fast(fmaf(s[e], kLog2e, -mb)) : 0.0f; z += s[e]; }
            const float inv = (tid < P) ? 1.0f / z : 0.0f;            // rows >= P do not exist: p = 0
            if (tid < kPts) {
#pragma unroll
                for (int q = 0; q < kPts / 8; ++q) {
                    uint4 o;
                    o.x = pack_bf16(s[8 * q] * inv, s[8 * q + 1] * inv); o.y = pack_bf16(s[8 * q + 2] * inv, s[8 * q + 3] * inv);
                    o.z = pack_bf16(s[8 * q + 4] * inv, s[8 * q + 5] * inv); o.w = pack_bf16(s[8 * q + 6] * inv, s[8 * q + 7] * inv);
                    *reinterpret_cast<uint4*>(smem + OFF_P + q * PP + tid * 16) = o;
                }
            }
        }
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
        mbar_wait(bar, phase); phase ^= 1;
        tc_after_sync();
        // ---- 4. row tid of O: relu(bn(gamma * o + x)) (model.py:312-313), o rounded to bf16 first like the unfused block
        __nv_bfloat16* op = out + ((size_t)pos * P + tid) * kC;
        const __nv_bfloat16* xr = xp + (size_t)tid * kC;
#pragma unroll 1
        for (int ch = 0; ch < kC / 32; ++ch) {
            uint32_t r[32];
            tmem_ld32(tlane + kColO + ch * 32, r);
            uint4 xv[4];
            if (tid < P) {
#pragma unroll
                for (int q = 0; q < 4; ++q) xv[q] = __ldg(reinterpret_cast<const uint4*>(xr + ch * 32 + q * 8));
            }
            tmem_ld_wait();
            if (dbg && pos == 0) {
#pragma unroll
                for (int e = 0; e < 32; ++e) dbg[kM * kNqkv + kM * kPts + tid * kC + ch * 32 + e] = __uint_as_float(r[e]);
            }
            if (tid < P) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t xw[4] = {xv[q].x, xv[q].y, xv[q].z, xv[q].w};
                    uint32_t ow[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = ch * 32 + q * 8 + 2 * k;
                        const float2 xf = unpack_bf16(xw[k]);
                        const float2 of = unpack_bf16(pack_bf16(__uint_as_float(r[q * 8 + 2 * k]), __uint_as_float(r[q * 8 + 2 * k + 1])));
                        const float s0 = fmaf(gamma, of.x, xf.x), s1 = fmaf(gamma, of.y, xf.y);
                        ow[k] = pack_bf16(fmaxf(fmaf(s0, s_scale[c], s_shift[c]), 0.0f), fmaxf(fmaf(s1, s_scale[c + 1], s_shift[c + 1]), 0.0f));
                    }
                    *reinterpret_cast<uint4*>(op + ch * 32 + q * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
            }
        }
    }
    tc_before_sync();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace tc
}  // namespace tg

extern "C" int tg_net_attention_fused(const void* x, const void* wqkv, const void* bqkv, float gamma, const float* scale, const float* shift,
                                      void* out, int64_t B, int P, int dq, int channels, float* debug, void* stream) {
    using namespace tg;
    using namespace tg::tc;
    if (!x || !wqkv || !bqkv || !scale || !shift || !out || B <= 0 || P <= 0 || P > kPts || dq != kDq || channels != kC)
        return set_err(-1, "tg_net_attention_fused (supports P <= 96, 32 query channels, 128 value channels)");
    static unsigned long long attr_done = 0;                 // one bit per device
    int dev = 0; cudaGetDevice(&dev);
    if (!((attr_done >> (dev & 63)) & 1ull)) {
        const cudaError_t e = cudaFuncSetAttribute(k_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return set_err((int)e, "tg_net_attention_fused (shared memory attribute)");
        attr_done |= 1ull << (dev & 63);
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = B < 2LL * sms ? B : 2LL * sms;    // persistent: two CTAs per SM, each walks positions blockIdx, + grid, ...
    k_attention_tc<<<(unsigned)grid, kTcThreads, SMEM_BYTES, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)wqkv, (const __nv_bfloat16*)bqkv, gamma, scale, shift, (__nv_bfloat16*)out, (int)B, P, debug);
    return set_err((int)cudaGetLastError(), "tg_net_attention_fused");
}
