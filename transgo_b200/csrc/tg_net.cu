// tg_net.cu -- the two memory-bound epilogues of the TransGo trunk in inference (model.py:208-262, :286-314), fused.
// The net itself stays in PyTorch / cuDNN / cuBLAS (north_star); after folding eval-mode BatchNorm into per-channel
// scale/shift what is left between the tensor-core kernels is elementwise traffic, and eager PyTorch spends ~70 % of a
// 4096-position forward on it (13 separate BN / ReLU / add kernels and layout copies).  These kernels do each residual
// junction in ONE pass over NHWC bf16 with 128-bit accesses:
//     sum = x + alpha * (y + ybias)  (residual add; ybias = the bias of the convolution that produced y, so that the
//                                     library convolution runs bias-free instead of launching a separate bias pass)
//     act = relu(sum * scale + shift) (the next block's BatchNorm + ReLU)
#include <cuda_bf16.h>
#include "../../include/transgo_b200.h"
#include "tg_common.cuh"

namespace tg {

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return q;
}

// x, y, sum_out, act_out: bf16 [rows][C] (NHWC flattened), C % 8 == 0; scale/shift: f32 [C].
// y == nullptr -> sum = x.  sum_out / act_out may be nullptr.  One thread = 8 channels of one row.
__device__ __forceinline__ void junction8(const uint4& xq, const uint4& yq, bool has_y, float alpha, const float* __restrict__ ybias,
                                          const float* __restrict__ scale, const float* __restrict__ shift, int c0,
                                          uint4* __restrict__ sum_out, uint4* __restrict__ act_out, long long i) {
    float a[8], b[8];
    unpack8(xq, a);
    if (has_y) {
        unpack8(yq, b);
        if (ybias) {
            const float4 b0 = *reinterpret_cast<const float4*>(ybias + c0), b1 = *reinterpret_cast<const float4*>(ybias + c0 + 4);
            b[0] += b0.x; b[1] += b0.y; b[2] += b0.z; b[3] += b0.w; b[4] += b1.x; b[5] += b1.y; b[6] += b1.z; b[7] += b1.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = a[k] + alpha * b[k];
    }
    uint4 s = pack8(a);
    if (sum_out) sum_out[i] = s;
    if (act_out) {
        unpack8(s, a);                      // BatchNorm sees the bf16-rounded sum, like the unfused bf16 model
        const float4 s0 = *reinterpret_cast<const float4*>(scale + c0), s1 = *reinterpret_cast<const float4*>(scale + c0 + 4);
        const float4 h0 = *reinterpret_cast<const float4*>(shift + c0), h1 = *reinterpret_cast<const float4*>(shift + c0 + 4);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fmaxf(a[k] * sc[k] + sh[k], 0.0f);
        act_out[i] = pack8(a);
    }
}

// Two 128-bit vectors per thread and iteration, all four loads issued before the first use (more bytes in flight per thread).
__global__ void __launch_bounds__(256)
k_add_bn_relu(const uint4* __restrict__ x, const uint4* __restrict__ y, float alpha, const float* __restrict__ ybias, const float* __restrict__ scale,
              const float* __restrict__ shift, uint4* __restrict__ sum_out, uint4* __restrict__ act_out, long long n_vec, int cvec, int reverse) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    // reverse: walk the tensors from the END.  y was written front to back by the kernel before this one, so its tail is what the
    // 126 MB L2 still holds; reading it first turns those reads into L2 hits instead of evicting them before they are reached.
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n_vec; j += 2 * stride) {
        const long long j2 = j + stride;
        const bool two = j2 < n_vec;
        const long long i = reverse ? n_vec - 1 - j : j, i2 = reverse ? n_vec - 1 - j2 : j2;
        const uint4 z = make_uint4(0, 0, 0, 0);
        const uint4 x0 = x[i], x1 = two ? x[i2] : z;
        const uint4 y0 = y ? y[i] : z, y1 = (y && two) ? y[i2] : z;
        junction8(x0, y0, y != nullptr, alpha, ybias, scale, shift, (int)(i % cvec) * 8, sum_out, act_out, i);
        if (two) junction8(x1, y1, y != nullptr, alpha, ybias, scale, shift, (int)(i2 % cvec) * 8, sum_out, act_out, i2);
    }
}

}  // namespace tg

// the same junction on float32 activations (the folded float32 schedule): one thread = 4 channels of one row
__global__ void __launch_bounds__(256)
k_add_bn_relu_f32(const float4* __restrict__ x, const float4* __restrict__ y, float alpha, const float* __restrict__ ybias,
                  const float* __restrict__ scale, const float* __restrict__ shift, float4* __restrict__ sum_out, float4* __restrict__ act_out,
                  long long n_vec, int cvec) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const int c0 = (int)(i % cvec) * 4;
        float4 a = x[i];
        if (y) {
            float4 b = y[i];
            if (ybias) { const float4 bb = *reinterpret_cast<const float4*>(ybias + c0); b.x += bb.x; b.y += bb.y; b.z += bb.z; b.w += bb.w; }
            a.x = a.x + alpha * b.x; a.y = a.y + alpha * b.y; a.z = a.z + alpha * b.z; a.w = a.w + alpha * b.w;
        }
        if (sum_out) sum_out[i] = a;
        if (act_out) {
            const float4 sc = *reinterpret_cast<const float4*>(scale + c0), sh = *reinterpret_cast<const float4*>(shift + c0);
            act_out[i] = make_float4(fmaxf(a.x * sc.x + sh.x, 0.0f), fmaxf(a.y * sc.y + sh.y, 0.0f), fmaxf(a.z * sc.z + sh.z, 0.0f), fmaxf(a.w * sc.w + sh.w, 0.0f));
        }
    }
}

static int g_net_reverse = 0;
extern "C" int tg_net_debug_reverse(int on) { const int old = g_net_reverse; g_net_reverse = on; return old; }

extern "C" int tg_net_add_bn_relu_f32(const float* x, const float* y, float alpha, const float* ybias, const float* scale, const float* shift,
                                       float* sum_out, float* act_out, int64_t rows, int channels, void* stream) {
    using namespace tg;
    if (!x || rows <= 0 || channels <= 0 || (channels & 3) || (!sum_out && !act_out) || (act_out && (!scale || !shift)))
        return set_err(-1, "tg_net_add_bn_relu_f32");
    const long long n_vec = rows * (long long)(channels / 4);
    long long blocks = (n_vec + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    k_add_bn_relu_f32<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)y, alpha, ybias, scale, shift,
                                                                            (float4*)sum_out, (float4*)act_out, n_vec, channels / 4);
    return set_err((int)cudaGetLastError(), "tg_net_add_bn_relu_f32");
}

extern "C" int tg_net_add_bn_relu(const void* x, const void* y, float alpha, const float* ybias, const float* scale, const float* shift,
                                   void* sum_out, void* act_out, int64_t rows, int channels, void* stream) {
    using namespace tg;
    if (!x || rows <= 0 || channels <= 0 || (channels & 7) || (!sum_out && !act_out) || (act_out && (!scale || !shift)))
        return set_err(-1, "tg_net_add_bn_relu");
    const long long n_vec = rows * (long long)(channels / 8);
    long long blocks = (n_vec + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;          // grid-stride over a multiple of the SM count
    k_add_bn_relu<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)y, alpha, ybias, scale, shift,
                                                                         (uint4*)sum_out, (uint4*)act_out, n_vec, channels / 8, g_net_reverse);
    return set_err((int)cudaGetLastError(), "tg_net_add_bn_relu");
}

// ------------------------------------------------------------------------------------------------------------------
// Fused spatial self-attention of the TransGo trunk in inference (model.py:286-314), one CTA per position:
//     S[i][j] = q_i . k_j          att = softmax_j(S)          o_j = sum_i att[i][j] * v_i
//     out = relu(bn(gamma * o + x))
// Eager PyTorch runs this as two batched GEMMs on [B,P,P] score tensors plus softmax, slicing copies and the junction
// (45 % of a 4096-position forward); here scores never leave the SM.  P <= 96 points (9x9), 32 query/key channels,
// 128 value channels, six warps:
//   1. warp w owns the queries i in [16w, 16w+16) and forms S^T[j][i] for EVERY key j (mma.sync m16n8k16, bf16 in, f32
//      accumulate), so the softmax over j is a reduction inside the warp (registers + three shuffles per column);
//   2. the normalised probabilities go to shared memory as bf16 att^T[j][i] (over the dead q / k tiles);
//   3. warp w then owns the OUTPUT rows j in [16w, 16w+16): o[j][c] = sum_i att^T[j][i] v[i][c] with the A operand from
//      ldmatrix and the B operand (v, row-major) from ldmatrix.trans, 16 channels at a time, junction epilogue on the
//      accumulators.
// (2 MFLOP and 72 KB of HBM traffic per position.)
namespace tg {

constexpr int kAttP = 96;          // padded points
constexpr int kAttDq = 32;
constexpr int kAttC = 128;
constexpr int kAttWarps = kAttP / 16;
constexpr int kQS = kAttDq + 8;    // smem row strides (bf16), padded against bank conflicts
constexpr int kVS = kAttC + 8;
constexpr int kAS = kAttP + 8;
constexpr float kLog2e = 1.4426950408889634f;
// exp(x) = ex2(x * log2 e): one FFMA + one MUFU
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t* r, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kAttWarps * 32)
k_attention(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ x, float gamma,
            const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, int P,
            const float* __restrict__ scale2, const float* __restrict__ shift2, __nv_bfloat16* __restrict__ out2) {
    constexpr int QK_ELEMS = 2 * kAttP * kQS, AT_ELEMS = kAttP * kAS;
    __shared__ __align__(16) __nv_bfloat16 sA[QK_ELEMS > AT_ELEMS ? QK_ELEMS : AT_ELEMS];   // q | k first, att^T afterwards
    __shared__ __align__(16) __nv_bfloat16 sV[kAttP * kVS];
    __nv_bfloat16* sQ = sA;
    __nv_bfloat16* sK = sA + kAttP * kQS;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
    const long long b = blockIdx.x;
    constexpr int ROWQ = (2 * kAttDq + kAttC) / 8;                   // 16-byte chunks per qkv row
    const uint4* src = reinterpret_cast<const uint4*>(qkv + b * (long long)P * (2 * kAttDq + kAttC));
    // asynchronous 16-byte copies straight into shared memory, q | k first and v as a second group, so that the score
    // product and the softmax run while v is still in flight; the residual rows of x are pulled towards L2 meanwhile
    constexpr int QKC = 2 * kAttDq / 8;
    for (int idx = tid; idx < kAttP * QKC; idx += blockDim.x) {
        const int row = idx / QKC, ch = idx - row * QKC;
        __nv_bfloat16* dst = ch < kAttDq / 8 ? &sQ[row * kQS + ch * 8] : &sK[row * kQS + (ch - kAttDq / 8) * 8];
        if (row < P) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src + row * ROWQ + ch));
        } else {
            *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("cp.async.commit_group;\n" ::);
    constexpr int VC = kAttC / 8;
    for (int idx = tid; idx < kAttP * VC; idx += blockDim.x) {
        const int row = idx / VC, ch = idx - row * VC;
        __nv_bfloat16* dst = &sV[row * kVS + ch * 8];
        if (row < P) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src + row * ROWQ + QKC + ch));
        } else {
            *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);  // rows >= P are zero: they add nothing to o
        }
    }
    asm volatile("cp.async.commit_group;\n" ::);
    {
        const char* xrow = reinterpret_cast<const char*>(x + b * (long long)P * kAttC);
        if (tid * 128 < P * kAttC * 2) asm volatile("prefetch.global.L2 [%0];\n" :: "l"(xrow + tid * 128));
    }
    asm volatile("cp.async.wait_group 1;\n" ::);
    __syncthreads();

    // ---- 1. S^T[j][i] = k_j . q_i for every key j and this warp's 16 queries i
    const int i0 = wid * 16;
    const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8, mcol = (lane >> 4) * 8;   // ldmatrix.x4 lane -> (row, col) of its 8x8 tile
    uint32_t qb[2][kAttDq / 16][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int kc = 0; kc < kAttDq / 16; ++kc) {
            qb[nt][kc][0] = *reinterpret_cast<const uint32_t*>(&sQ[(i0 + nt * 8 + g) * kQS + kc * 16 + 2 * t]);
            qb[nt][kc][1] = *reinterpret_cast<const uint32_t*>(&sQ[(i0 + nt * 8 + g) * kQS + kc * 16 + 8 + 2 * t]);
        }
    float s[kAttWarps][2][4];
#pragma unroll
    for (int mt = 0; mt < kAttWarps; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.0f;
#pragma unroll
        for (int kc = 0; kc < kAttDq / 16; ++kc) {
            uint32_t ka[4];
            ldsm_x4(ka, &sK[(mt * 16 + mrow) * kQS + kc * 16 + mcol]);
            mma_bf16(s[mt][0], ka, qb[0][kc][0], qb[0][kc][1]);
            mma_bf16(s[mt][1], ka, qb[1][kc][0], qb[1][kc][1]);
        }
    }
    // softmax over the keys j (rows of S^T) for each of this thread's 4 query columns: registers, then the 8 lanes that
    // share t.  Rows j >= P do not exist: they are left out of max and sum and get probability 0.
    const float NEG = -3.0e38f;
    float inv[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float m = NEG;
#pragma unroll
            for (int mt = 0; mt < kAttWarps; ++mt) {
                if (mt * 16 + g < P) m = fmaxf(m, s[mt][nt][e]);
                if (mt * 16 + g + 8 < P) m = fmaxf(m, s[mt][nt][2 + e]);
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float z = 0.0f;
#pragma unroll
            for (int mt = 0; mt < kAttWarps; ++mt) {
                const float e0 = (mt * 16 + g < P) ? ex2_fast(fmaf(s[mt][nt][e], kLog2e, -m * kLog2e)) : 0.0f;
                const float e1 = (mt * 16 + g + 8 < P) ? ex2_fast(fmaf(s[mt][nt][2 + e], kLog2e, -m * kLog2e)) : 0.0f;
                s[mt][nt][e] = e0; s[mt][nt][2 + e] = e1;
                z += e0 + e1;
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
            inv[nt][e] = 1.0f / z;                                   // z >= 1: the row that holds the max contributes exp(0)
        }
    __syncthreads();                                                 // every warp is done with q / k: att^T may overwrite them
    // ---- 2. att^T[j][i] as bf16
#pragma unroll
    for (int mt = 0; mt < kAttWarps; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int col = i0 + nt * 8 + 2 * t;
            *reinterpret_cast<uint32_t*>(&sA[(mt * 16 + g) * kAS + col]) = pack_bf16(s[mt][nt][0] * inv[nt][0], s[mt][nt][1] * inv[nt][1]);
            *reinterpret_cast<uint32_t*>(&sA[(mt * 16 + g + 8) * kAS + col]) = pack_bf16(s[mt][nt][2] * inv[nt][0], s[mt][nt][3] * inv[nt][1]);
        }
    asm volatile("cp.async.wait_group 0;\n" ::);                    // v has landed (each thread waits for its own copies ...)
    __syncthreads();                                                 // ... and the barrier makes all of them, and att^T, visible
    // ---- 3. o[j][c] = sum_i att^T[j][i] * v[i][c] for this warp's 16 output rows j, then the junction epilogue
    const int j0 = wid * 16;
    uint32_t pa[kAttP / 16][4];
#pragma unroll
    for (int kc = 0; kc < kAttP / 16; ++kc) ldsm_x4(pa[kc], &sA[(j0 + mrow) * kAS + kc * 16 + mcol]);
    const long long rowbase = b * (long long)P;
#pragma unroll
    for (int ct = 0; ct < kAttC / 16; ++ct) {
        float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int kc = 0; kc < kAttP / 16; ++kc) {
            uint32_t v[4];
            ldsm_x4_trans(v, &sV[(kc * 16 + mrow) * kVS + ct * 16 + mcol]);
            mma_bf16(o[0], pa[kc], v[0], v[1]);
            mma_bf16(o[1], pa[kc], v[2], v[3]);
        }
        // the accumulator layout gives each lane 2 channels of 4 (row, half) combinations; a 4x4 transpose inside the quad
        // (two shuffle rounds) turns that into 8 consecutive channels of ONE combination per lane -> 128-bit accesses
        uint32_t a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = pack_bf16(o[k >> 1][2 * (k & 1)], o[k >> 1][2 * (k & 1) + 1]);
        {
            const bool odd = t & 1, hi = t & 2;
            const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? a[0] : a[1], 1), r1 = __shfl_xor_sync(0xffffffffu, odd ? a[2] : a[3], 1);
            if (odd) { a[0] = r0; a[2] = r1; } else { a[1] = r0; a[3] = r1; }
            const uint32_t q0 = __shfl_xor_sync(0xffffffffu, hi ? a[0] : a[2], 2), q1 = __shfl_xor_sync(0xffffffffu, hi ? a[1] : a[3], 2);
            if (hi) { a[0] = q0; a[1] = q1; } else { a[2] = q0; a[3] = q1; }
        }
        const int j = j0 + g + (t & 1) * 8, c = ct * 16 + (t >> 1) * 8;
        if (j < P) {
            const long long off = (rowbase + j) * kAttC + c;
            const uint4 xq = *reinterpret_cast<const uint4*>(x + off);
            const float4 sc0 = *reinterpret_cast<const float4*>(scale + c), sc1 = *reinterpret_cast<const float4*>(scale + c + 4);
            const float4 sh0 = *reinterpret_cast<const float4*>(shift + c), sh1 = *reinterpret_cast<const float4*>(shift + c + 4);
            const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
            const float shv[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
            const uint32_t xw[4] = {xq.x, xq.y, xq.z, xq.w};
            uint32_t ow[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 xv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw[k]));
                const float2 ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a[k]));
                const float s0 = fmaf(gamma, ov.x, xv.x), s1 = fmaf(gamma, ov.y, xv.y);
                ow[k] = pack_bf16(fmaxf(fmaf(s0, scv[2 * k], shv[2 * k]), 0.0f), fmaxf(fmaf(s1, scv[2 * k + 1], shv[2 * k + 1]), 0.0f));
            }
            *reinterpret_cast<uint4*>(out + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            if (out2) {                                              // the next residual block's opening activation (like tg_net_add_bn_relu on `out`)
                uint32_t o2[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[k]));
                    o2[k] = pack_bf16(fmaxf(__fadd_rn(__fmul_rn(f.x, scale2[c + 2 * k]), shift2[c + 2 * k]), 0.0f),
                                      fmaxf(__fadd_rn(__fmul_rn(f.y, scale2[c + 2 * k + 1]), shift2[c + 2 * k + 1]), 0.0f));
                }
                *reinterpret_cast<uint4*>(out2 + off) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The same block for boards of up to 368 points (19x19 = 361): scores no longer fit in registers, so the CTA (12 warps,
// one position, q | k | v resident in 160 KB of shared memory) makes two passes over the score matrix and never stores it:
//   A. warp w owns 16-query tiles and streams all keys: S[i][j] = q_i . k_j by mma.sync, per-row running max / sum kept per
//      thread over its own columns and merged inside the quad at the end -> m_i and 1 / z_i in shared memory;
//   B. warp w owns 16-key (= output row) tiles and streams all queries 16 at a time: S^T[j][i] recomputed by mma.sync,
//      p = exp(s - m_i) / z_i, re-used from the accumulator registers as the A operand of o[j][c] += p^T[j][i] v[i][c]
//      (B operand: ldmatrix.trans on row-major v), 128 output channels in registers; junction epilogue as above.
constexpr int kBigP = 368;                      // 23 tiles of 16
constexpr int kBigTiles = kBigP / 16;
constexpr int kBigWarps = 12;

constexpr size_t kBigSmem = (size_t)kBigP * kQS * 2 * 2 + (size_t)kBigP * kVS * 2 + (size_t)kBigP * 4 * 2;

__global__ void __launch_bounds__(kBigWarps * 32, 1)
k_attention_big(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ x, float gamma,
                const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, int P,
            const float* __restrict__ scale2, const float* __restrict__ shift2, __nv_bfloat16* __restrict__ out2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* sK = sQ + kBigP * kQS;
    __nv_bfloat16* sV = sK + kBigP * kQS;
    float* sM = reinterpret_cast<float*>(sV + kBigP * kVS);
    float* sIZ = sM + kBigP;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
    const long long b = blockIdx.x;
    constexpr int ROWQ = (2 * kAttDq + kAttC) / 8, QKC = 2 * kAttDq / 8, VC = kAttC / 8;
    const uint4* src = reinterpret_cast<const uint4*>(qkv + b * (long long)P * (2 * kAttDq + kAttC));
    for (int idx = tid; idx < kBigP * QKC; idx += blockDim.x) {
        const int row = idx / QKC, ch = idx - row * QKC;
        __nv_bfloat16* dst = ch < kAttDq / 8 ? &sQ[row * kQS + ch * 8] : &sK[row * kQS + (ch - kAttDq / 8) * 8];
        if (row < P) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src + row * ROWQ + ch));
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    for (int idx = tid; idx < kBigP * VC; idx += blockDim.x) {
        const int row = idx / VC, ch = idx - row * VC;
        __nv_bfloat16* dst = &sV[row * kVS + ch * 8];
        if (row < P) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src + row * ROWQ + QKC + ch));
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);       // rows >= P are zero: they add nothing to o
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 1;\n" ::);
    __syncthreads();
    const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8, mcol = (lane >> 4) * 8;   // ldmatrix.x4 lane -> (row, col) of its 8x8 tile
    const float NEG = -3.0e38f;

    // ---- A. row statistics of S = Q K^T: this warp's query tiles against every key
    for (int it = wid; it < kBigTiles; it += kBigWarps) {
        uint32_t qa[kAttDq / 16][4];
#pragma unroll
        for (int kc = 0; kc < kAttDq / 16; ++kc) ldsm_x4(qa[kc], &sQ[(it * 16 + mrow) * kQS + kc * 16 + mcol]);
        float m[2] = {NEG, NEG}, z[2] = {0.0f, 0.0f};                  // rows g and g + 8, over this thread's own columns
        for (int jt = 0; jt < kBigTiles; ++jt) {
            float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int kc = 0; kc < kAttDq / 16; ++kc) {
                    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sK[(jt * 16 + nt * 8 + g) * kQS + kc * 16 + 2 * t]);
                    const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sK[(jt * 16 + nt * 8 + g) * kQS + kc * 16 + 8 + 2 * t]);
                    mma_bf16(s[nt], qa[kc], b0, b1);
                }
            const bool full = jt * 16 + 16 <= P;                         // warp-uniform: only the last key tile is ragged
#pragma unroll
            for (int r = 0; r < 2; ++r) {                               // accumulator rows: c0,c1 -> row g ; c2,c3 -> row g + 8
                float v[4];                                             // scores in units of log2(e): exp(x) = exp2(x * log2 e)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const bool ok = full || jt * 16 + nt * 8 + 2 * t + e < P;    // keys >= P do not exist
                        v[nt * 2 + e] = ok ? s[nt][2 * r + e] * kLog2e : NEG;
                    }
                const float nm = fmaxf(fmaxf(m[r], fmaxf(v[0], v[1])), fmaxf(v[2], v[3]));
                float acc = z[r] * ex2_fast(m[r] - nm);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc += ex2_fast(v[q] - nm);    // ex2_fast(NEG - nm) == 0
                m[r] = nm; z[r] = acc;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {                                   // merge the four threads of the quad (they split the columns)
            float mm = m[r];
            mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, 1));
            mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, 2));
            float zz = z[r] * ex2_fast(m[r] - mm);
            zz += __shfl_xor_sync(0xffffffffu, zz, 1);
            zz += __shfl_xor_sync(0xffffffffu, zz, 2);
            if (t == 0) { sM[it * 16 + g + 8 * r] = mm; sIZ[it * 16 + g + 8 * r] = 1.0f / zz; }   // mm in log2 units
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();                                                     // statistics and v visible to every warp

    // ---- B. o[j][c] = sum_i p[i][j] v[i][c] for this warp's output-row tiles
    const long long rowbase = b * (long long)P;
    for (int jt = wid; jt < kBigTiles; jt += kBigWarps) {
        uint32_t ka[kAttDq / 16][4];
#pragma unroll
        for (int kc = 0; kc < kAttDq / 16; ++kc) ldsm_x4(ka[kc], &sK[(jt * 16 + mrow) * kQS + kc * 16 + mcol]);
        const bool r0 = jt * 16 + g < P, r1 = jt * 16 + g + 8 < P;
        float o[kAttC / 8][4];
#pragma unroll
        for (int n = 0; n < kAttC / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f;
        for (int ic = 0; ic < kBigTiles; ++ic) {
            float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int kc = 0; kc < kAttDq / 16; ++kc) {
                    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sQ[(ic * 16 + nt * 8 + g) * kQS + kc * 16 + 2 * t]);
                    const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sQ[(ic * 16 + nt * 8 + g) * kQS + kc * 16 + 8 + 2 * t]);
                    mma_bf16(s[nt], ka[kc], b0, b1);
                }
            uint32_t pa[4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int col = ic * 16 + nt * 8 + 2 * t;
                const float2 mm = *reinterpret_cast<const float2*>(&sM[col]), iz = *reinterpret_cast<const float2*>(&sIZ[col]);
                const float p0 = r0 ? ex2_fast(fmaf(s[nt][0], kLog2e, -mm.x)) * iz.x : 0.0f, p1 = r0 ? ex2_fast(fmaf(s[nt][1], kLog2e, -mm.y)) * iz.y : 0.0f;
                const float p2 = r1 ? ex2_fast(fmaf(s[nt][2], kLog2e, -mm.x)) * iz.x : 0.0f, p3 = r1 ? ex2_fast(fmaf(s[nt][3], kLog2e, -mm.y)) * iz.y : 0.0f;
                pa[2 * nt] = pack_bf16(p0, p1);
                pa[2 * nt + 1] = pack_bf16(p2, p3);
            }
#pragma unroll
            for (int ct = 0; ct < kAttC / 16; ++ct) {
                uint32_t v[4];
                ldsm_x4_trans(v, &sV[(ic * 16 + mrow) * kVS + ct * 16 + mcol]);
                mma_bf16(o[2 * ct], pa, v[0], v[1]);
                mma_bf16(o[2 * ct + 1], pa, v[2], v[3]);
            }
        }
#pragma unroll
        for (int ct = 0; ct < kAttC / 16; ++ct) {
            uint32_t a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = pack_bf16(o[2 * ct + (k >> 1)][2 * (k & 1)], o[2 * ct + (k >> 1)][2 * (k & 1) + 1]);
            {
                const bool odd = t & 1, hi = t & 2;
                const uint32_t x0 = __shfl_xor_sync(0xffffffffu, odd ? a[0] : a[1], 1), x1 = __shfl_xor_sync(0xffffffffu, odd ? a[2] : a[3], 1);
                if (odd) { a[0] = x0; a[2] = x1; } else { a[1] = x0; a[3] = x1; }
                const uint32_t q0 = __shfl_xor_sync(0xffffffffu, hi ? a[0] : a[2], 2), q1 = __shfl_xor_sync(0xffffffffu, hi ? a[1] : a[3], 2);
                if (hi) { a[0] = q0; a[1] = q1; } else { a[2] = q0; a[3] = q1; }
            }
            const int j = jt * 16 + g + (t & 1) * 8, c = ct * 16 + (t >> 1) * 8;
            if (j < P) {
                const long long off = (rowbase + j) * kAttC + c;
                const uint4 xq = *reinterpret_cast<const uint4*>(x + off);
                const float4 sc0 = *reinterpret_cast<const float4*>(scale + c), sc1 = *reinterpret_cast<const float4*>(scale + c + 4);
                const float4 sh0 = *reinterpret_cast<const float4*>(shift + c), sh1 = *reinterpret_cast<const float4*>(shift + c + 4);
                const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
                const float shv[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
                const uint32_t xw[4] = {xq.x, xq.y, xq.z, xq.w};
                uint32_t ow[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 xv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw[k]));
                    const float2 ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a[k]));
                    const float s0 = fmaf(gamma, ov.x, xv.x), s1 = fmaf(gamma, ov.y, xv.y);
                    ow[k] = pack_bf16(fmaxf(fmaf(s0, scv[2 * k], shv[2 * k]), 0.0f), fmaxf(fmaf(s1, scv[2 * k + 1], shv[2 * k + 1]), 0.0f));
                }
                *reinterpret_cast<uint4*>(out + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            if (out2) {                                              // the next residual block's opening activation (like tg_net_add_bn_relu on `out`)
                uint32_t o2[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[k]));
                    o2[k] = pack_bf16(fmaxf(__fadd_rn(__fmul_rn(f.x, scale2[c + 2 * k]), shift2[c + 2 * k]), 0.0f),
                                      fmaxf(__fadd_rn(__fmul_rn(f.y, scale2[c + 2 * k + 1]), shift2[c + 2 * k + 1]), 0.0f));
                }
                *reinterpret_cast<uint4*>(out2 + off) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
            }
            }
        }
    }
}

}  // namespace tg

// ------------------------------------------------------------------------------------------------------------------
// The two head convolutions of the net (model.py:93-94, :105-106: 3x3, 128 -> 2 / 4 channels, BatchNorm folded, ReLU).  A library
// implicit GEMM pads the 2 / 4 output channels to a 64-wide tile and takes 87 us for 4096 positions; the work is one pass over x
// (85 MB).  Here a position's [81][128] bf16 tile is copied by cp.async into a zero-bordered 11x11 FRAME in shared memory (two
// buffers: the next position loads while this one is multiplied) and the nine taps become extra OUTPUT COLUMNS of one plain GEMM:
//     Y[f][tap * nout + co] = sum_ci frame[f][ci] * w[co][tap][ci]          (mma.sync m16n8k16: M = 128 frame rows, N = 9 nout, K = 128)
//     out[p][co] = relu(bias[co] + sum_tap Y[f(p) + 11 dy + dx][tap * nout + co])
// so that every A fragment is loaded once and feeds all N tiles, and the B fragments (the weights) live in registers for the whole
// launch -- a tap-by-tap formulation re-reads the tile nine times for 8 output columns and is bound by shared-memory bandwidth
// (75 us measured).  Eight warps, one 16-row M tile each; out bf16 [B][P][nout].
namespace tg {

constexpr int kHcN = 9, kHcP = 81, kHcF = 11;                             // board, points, frame side (121 frame rows)
constexpr int kHcRS = kAttC * 2 + 16;                                     // frame row stride (bytes): 4 banks per row
constexpr int kHcTileRows = 128;                                          // 8 M tiles; rows >= 121 stay zero
constexpr int kHcTile = kHcTileRows * kHcRS;
constexpr int kHcWarps = 8;
template <int NT> __host__ __device__ constexpr int hc_ystride() { return NT * 8 + 1; }       // floats per Y row (+1: the gather walks rows)
template <int NT> __host__ __device__ constexpr int hc_smem() { return 2 * kHcTile + kHcTileRows * hc_ystride<NT>() * 4; }

template <int NT>                                                         // NT = N tiles of 8 columns: ceil(9 * nout / 8)
__global__ void __launch_bounds__(kHcWarps * 32, 2)
k_conv3x3_head(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
               __nv_bfloat16* __restrict__ out, int B, int nout) {
    extern __shared__ __align__(16) unsigned char hc_smem_raw[];
    float* sY = reinterpret_cast<float*>(hc_smem_raw + 2 * kHcTile);
    constexpr int YS = hc_ystride<NT>();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
    // zero both frames once (borders and rows >= 121 stay zero: the copies only ever touch interior rows)
    for (int i = tid; i < 2 * kHcTile / 16; i += blockDim.x) reinterpret_cast<uint4*>(hc_smem_raw)[i] = make_uint4(0, 0, 0, 0);
    // B fragments of every (N tile, k step) in registers: column n = tap * nout + co of the GEMM is w[co][tap][:]
    uint32_t bw[NT][kAttC / 16][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int n = nt * 8 + g;
        const bool live = n < 9 * nout;
        const int tap = live ? n / nout : 0, co = live ? n - tap * nout : 0;
        const uint32_t* wr = reinterpret_cast<const uint32_t*>(w + ((size_t)co * 9 + tap) * kAttC);
#pragma unroll
        for (int kc = 0; kc < kAttC / 16; ++kc) {
            bw[nt][kc][0] = live ? wr[kc * 8 + t] : 0u;
            bw[nt][kc][1] = live ? wr[kc * 8 + 4 + t] : 0u;
        }
    }
    __syncthreads();
    auto load_tile = [&](int pos, int buf) {
        const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)pos * kHcP * kAttC);
        unsigned char* dst = hc_smem_raw + buf * kHcTile;
        for (int i = tid; i < kHcP * (kAttC / 8); i += blockDim.x) {
            const int p = i >> 4, q = i & 15;
            const int fr = (p / kHcN + 1) * kHcF + (p % kHcN) + 1;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"((uint32_t)__cvta_generic_to_shared(dst + fr * kHcRS + q * 16)), "l"(src + i));
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8, mcol = (lane >> 4) * 8;   // ldmatrix.x4 lane -> (row, col) of its 8x8 tile
    int buf = 0;
    if ((int)blockIdx.x < B) load_tile(blockIdx.x, 0);
    for (int pos = blockIdx.x; pos < B; pos += gridDim.x) {
        const int nxt = pos + gridDim.x;
        if (nxt < B) load_tile(nxt, buf ^ 1);                                     // (that buffer was fully read an iteration ago)
        if (nxt < B) asm volatile("cp.async.wait_group 1;\n" ::); else asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();                                                          // the tile is in; the previous gather is done with sY
        const unsigned char* arow = hc_smem_raw + buf * kHcTile + (wid * 16 + mrow) * kHcRS + mcol * 2;
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
#pragma unroll
        for (int kc = 0; kc < kAttC / 16; ++kc) {
            uint32_t a[4];
            ldsm_x4(a, arow + kc * 32);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_bf16(acc[nt], a, bw[nt][kc][0], bw[nt][kc][1]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {                                         // rows g, g + 8 of this warp's tile; columns 2t, 2t + 1
            float* y0 = sY + (wid * 16 + g) * YS + nt * 8 + 2 * t;
            y0[0] = acc[nt][0]; y0[1] = acc[nt][1];
            y0[8 * YS] = acc[nt][2]; y0[8 * YS + 1] = acc[nt][3];
        }
        __syncthreads();
        // the nine shifted partial sums of every output
        __nv_bfloat16* o = out + (size_t)pos * kHcP * nout;
        for (int i = tid; i < kHcP * nout; i += blockDim.x) {
            const int p = i / nout, co = i - p * nout;
            const int fr = (p / kHcN + 1) * kHcF + (p % kHcN) + 1;
            float v = bias[co];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) v += sY[(fr + (tap / 3 - 1) * kHcF + (tap % 3 - 1)) * YS + tap * nout + co];
            o[i] = __float2bfloat16(fmaxf(v, 0.0f));
        }
        buf ^= 1;
    }
}

}  // namespace tg

extern "C" int tg_net_conv3x3_head(const void* x, const void* w, const float* bias, void* out, int64_t B, int board, int channels, int nout,
                                    void* stream) {
    using namespace tg;
    if (!x || !w || !bias || !out || B <= 0 || board != kHcN || channels != kAttC || nout < 1 || nout > 4)
        return set_err(-1, "tg_net_conv3x3_head (supports 9x9 boards, 128 input channels, 1..4 output channels)");
    static unsigned long long attr_done = 0;                 // one bit per device
    int dev = 0; cudaGetDevice(&dev);
    if (!((attr_done >> (dev & 63)) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k_conv3x3_head<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, hc_smem<3>());
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv3x3_head<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, hc_smem<5>());
        if (e != cudaSuccess) return set_err((int)e, "tg_net_conv3x3_head (shared memory attribute)");
        attr_done |= 1ull << (dev & 63);
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = B < 2LL * sms ? B : 2LL * sms;
    if (9 * nout <= 24)
        k_conv3x3_head<3><<<(unsigned)grid, kHcWarps * 32, hc_smem<3>(), (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias,
                                                                                               (__nv_bfloat16*)out, (int)B, nout);
    else
        k_conv3x3_head<5><<<(unsigned)grid, kHcWarps * 32, hc_smem<5>(), (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias,
                                                                                               (__nv_bfloat16*)out, (int)B, nout);
    return set_err((int)cudaGetLastError(), "tg_net_conv3x3_head");
}

// ------------------------------------------------------------------------------------------------------------------
// The stem of the net (model.py:43, :81: 3x3 convolution from the 10 feature planes to 128 channels, BatchNorm folded, ReLU)
// straight from the float32 planes [B][10][81] the leaf kernels write, to NHWC bf16 -- and, optionally, the opening activation
// relu(bn1(out)) of the first residual block as a second output.  The library route is a staging copy into a 16-channel NHWC
// tensor, an implicit GEMM (69 us for 4096 positions: K = 90 is tiny, the 85 MB output is the work) and a junction pass.
// One position at a time per CTA: planes -> zero-bordered 11x11 frame [row][16 channels] in shared memory (double buffered, the
// next position's planes are fetched into registers meanwhile); a tap is a row offset and ONE k step of mma.sync m16n8k16
// (10 of 16 channels used).  Warp w owns output channels [16w, 16w + 16) for all six 16-pixel M tiles, its 18 B fragments live in
// registers for the whole launch.
namespace tg {

constexpr int kStRS = 48;                               // frame row stride (bytes): 16 channels bf16 + 16 pad -> conflict-free ldmatrix
constexpr int kStTile = 128 * kStRS;
constexpr int kStWarps = 8;

// 4 x 4 transpose of 32-bit words inside a quad (lanes t = 0..3 of the same g), two shuffle rounds.  In: a[k], k = 2 * (row half) +
// (8-column tile), each word = this lane's columns 2t, 2t + 1 of that (row half, tile).  Out: the four words are columns 0..7 of
// ONE (row half, tile) combination: row half = t & 1, tile = t >> 1.
__device__ __forceinline__ void quad_transpose(uint32_t* a, int t) {
    const bool odd = t & 1, hi = t & 2;
    // reorder to the accumulator order the attention epilogue uses: k = 2 * tile + row half
    const uint32_t s1 = a[1]; a[1] = a[2]; a[2] = s1;
    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? a[0] : a[1], 1), r1 = __shfl_xor_sync(0xffffffffu, odd ? a[2] : a[3], 1);
    if (odd) { a[0] = r0; a[2] = r1; } else { a[1] = r0; a[3] = r1; }
    const uint32_t q0 = __shfl_xor_sync(0xffffffffu, hi ? a[0] : a[2], 2), q1 = __shfl_xor_sync(0xffffffffu, hi ? a[1] : a[3], 2);
    if (hi) { a[0] = q0; a[1] = q1; } else { a[2] = q0; a[3] = q1; }
}

__global__ void __launch_bounds__(kStWarps * 32, 2)
k_stem(const float* __restrict__ obs, int cin, const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
       const float* __restrict__ scale2, const float* __restrict__ shift2, __nv_bfloat16* __restrict__ out,
       __nv_bfloat16* __restrict__ out2, int B) {
    __shared__ __align__(16) unsigned char st_smem[2 * kStTile];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < 2 * kStTile / 16; i += blockDim.x) reinterpret_cast<uint4*>(st_smem)[i] = make_uint4(0, 0, 0, 0);
    // B fragments: column n of tap `tap` is w[n][tap][0..15]
    uint32_t bw[2][9][2];
    float bs[2][2], s2[2][2], t2[2][2];
#pragma unroll
    for (int nl = 0; nl < 2; ++nl) {
        const int n = wid * 16 + nl * 8 + g;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const uint32_t* wr = reinterpret_cast<const uint32_t*>(w + ((size_t)n * 9 + tap) * 16);
            bw[nl][tap][0] = wr[t]; bw[nl][tap][1] = wr[4 + t];
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = wid * 16 + nl * 8 + 2 * t + e;               // the accumulator columns of this lane
            bs[nl][e] = bias[c];
            s2[nl][e] = out2 ? scale2[c] : 0.0f; t2[nl][e] = out2 ? shift2[c] : 0.0f;
        }
    }
    const int per = cin * kHcP;                                      // floats of one position
    const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8, mcol = (lane >> 4) * 8;
    int fra[6];
#pragma unroll
    for (int mt = 0; mt < 6; ++mt) { int p = mt * 16 + mrow; p = p < kHcP ? p : kHcP - 1; fra[mt] = (p / kHcN + 1) * kHcF + (p % kHcN) + 1; }
    float nx[4];
    auto fetch = [&](int pos) {
        const float* src = obs + (size_t)pos * per;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int e = tid + k * 256; nx[k] = e < per ? __ldg(src + e) : 0.0f; }
    };
    if ((int)blockIdx.x < B) fetch(blockIdx.x);
    __syncthreads();
    int buf = 0;
    for (int pos = blockIdx.x; pos < B; pos += gridDim.x) {
        unsigned char* tile = st_smem + buf * kStTile;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = tid + k * 256;
            if (e < per) {
                const int c = e / kHcP, p = e - c * kHcP;
                const int fr = (p / kHcN + 1) * kHcF + (p % kHcN) + 1;
                *reinterpret_cast<__nv_bfloat16*>(tile + fr * kStRS + c * 2) = __float2bfloat16(nx[k]);
            }
        }
        if (pos + (int)gridDim.x < B) fetch(pos + gridDim.x);
        __syncthreads();
        float acc[6][2][4];
#pragma unroll
        for (int mt = 0; mt < 6; ++mt)
#pragma unroll
            for (int nl = 0; nl < 2; ++nl) acc[mt][nl][0] = acc[mt][nl][1] = acc[mt][nl][2] = acc[mt][nl][3] = 0.0f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int off = (tap / 3 - 1) * kHcF + (tap % 3 - 1);
#pragma unroll
            for (int mt = 0; mt < 6; ++mt) {
                uint32_t a[4];
                ldsm_x4(a, tile + (fra[mt] + off) * kStRS + mcol * 2);
                mma_bf16(acc[mt][0], a, bw[0][tap][0], bw[0][tap][1]);
                mma_bf16(acc[mt][1], a, bw[1][tap][0], bw[1][tap][1]);
            }
        }
        __nv_bfloat16* o = out + (size_t)pos * kHcP * kAttC;
        __nv_bfloat16* o2 = out2 ? out2 + (size_t)pos * kHcP * kAttC : nullptr;
#pragma unroll
        for (int mt = 0; mt < 6; ++mt) {
            // this lane's four bf16 pairs of the 16 x 16 tile (rows g / g + 8, channel tiles 0 / 1), then a 4 x 4 transpose inside
            // the quad: every lane ends up with 8 consecutive channels of ONE row -> 16-byte stores, full sectors
            uint32_t a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = k >> 1, nl = k & 1;
                a[k] = pack_bf16(fmaxf(acc[mt][nl][2 * r] + bs[nl][0], 0.0f), fmaxf(acc[mt][nl][2 * r + 1] + bs[nl][1], 0.0f));
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a[k]));   // like tg_net_add_bn_relu on the rounded value
                b[k] = pack_bf16(fmaxf(__fadd_rn(__fmul_rn(f.x, s2[nl][0]), t2[nl][0]), 0.0f), fmaxf(__fadd_rn(__fmul_rn(f.y, s2[nl][1]), t2[nl][1]), 0.0f));
            }
            // a[] is indexed [row half][channel tile]; the transpose wants [channel tile][row half] pairs: reorder to (r0 n0, r0 n1, r1 n0, r1 n1)
            quad_transpose(a, t);
            if (o2) quad_transpose(b, t);
            const int p = mt * 16 + g + (t & 1) * 8, c = wid * 16 + (t >> 1) * 8;
            if (p < kHcP) {
                *reinterpret_cast<uint4*>(o + p * kAttC + c) = make_uint4(a[0], a[1], a[2], a[3]);
                if (o2) *reinterpret_cast<uint4*>(o2 + p * kAttC + c) = make_uint4(b[0], b[1], b[2], b[3]);
            }
        }
        buf ^= 1;
    }
}

}  // namespace tg

extern "C" int tg_net_stem(const float* obs, int cin, const void* w, const float* bias, const float* scale2, const float* shift2, void* out,
                            void* out2, int64_t B, int board, int channels, void* stream) {
    using namespace tg;
    if (!obs || !w || !bias || !out || B <= 0 || board != kHcN || channels != kAttC || cin < 1 || cin > 12 || (out2 && (!scale2 || !shift2)))
        return set_err(-1, "tg_net_stem (supports 9x9 boards, up to 12 input planes, 128 output channels)");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = B < 2LL * sms ? B : 2LL * sms;
    k_stem<<<(unsigned)grid, kStWarps * 32, 0, (cudaStream_t)stream>>>(obs, cin, (const __nv_bfloat16*)w, bias, scale2, shift2, (__nv_bfloat16*)out,
                                                                        (__nv_bfloat16*)out2, (int)B);
    return set_err((int)cudaGetLastError(), "tg_net_stem");
}

extern "C" int tg_net_attention2(const void* qkv, const void* x, float gamma, const float* scale, const float* shift, void* out,
                                 const float* scale2, const float* shift2, void* out2, int64_t B, int P, int dq, int channels, void* stream) {
    using namespace tg;
    if (!qkv || !x || !scale || !shift || !out || B <= 0 || P <= 0 || P > kBigP || dq != kAttDq || channels != kAttC || (out2 && (!scale2 || !shift2)))
        return set_err(-1, "tg_net_attention (supports P <= 368, 32 query channels, 128 value channels)");
    if (P <= kAttP) {
        k_attention<<<(unsigned)B, kAttWarps * 32, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)x, gamma, scale,
                                                                               shift, (__nv_bfloat16*)out, P, scale2, shift2, (__nv_bfloat16*)out2);
    } else {
        static unsigned long long attr_done = 0;             // one bit per device: the attribute is per device, not per process
        int dev = 0; cudaGetDevice(&dev);
        if (!((attr_done >> (dev & 63)) & 1ull)) {
            cudaFuncSetAttribute(k_attention_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBigSmem);
            attr_done |= 1ull << (dev & 63);
        }
        k_attention_big<<<(unsigned)B, kBigWarps * 32, kBigSmem, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)x, gamma,
                                                                                          scale, shift, (__nv_bfloat16*)out, P, scale2, shift2, (__nv_bfloat16*)out2);
    }
    return set_err((int)cudaGetLastError(), "tg_net_attention");
}

extern "C" int tg_net_attention(const void* qkv, const void* x, float gamma, const float* scale, const float* shift, void* out,
                                int64_t B, int P, int dq, int channels, void* stream) {
    return tg_net_attention2(qkv, x, gamma, scale, shift, out, nullptr, nullptr, nullptr, B, P, dq, channels, stream);
}
