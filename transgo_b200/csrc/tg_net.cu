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
__global__ void __launch_bounds__(256)
k_add_bn_relu(const uint4* __restrict__ x, const uint4* __restrict__ y, float alpha, const float* __restrict__ ybias, const float* __restrict__ scale,
              const float* __restrict__ shift, uint4* __restrict__ sum_out, uint4* __restrict__ act_out, long long n_vec, int cvec) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % cvec) * 8;
        float a[8], b[8];
        unpack8(x[i], a);
        if (y) {
            unpack8(y[i], b);
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
}

}  // namespace tg

extern "C" int tg_net_add_bn_relu(const void* x, const void* y, float alpha, const float* ybias, const float* scale, const float* shift,
                                   void* sum_out, void* act_out, int64_t rows, int channels, void* stream) {
    using namespace tg;
    if (!x || rows <= 0 || channels <= 0 || (channels & 7) || (!sum_out && !act_out) || (act_out && (!scale || !shift)))
        return set_err(-1, "tg_net_add_bn_relu");
    const long long n_vec = rows * (long long)(channels / 8);
    long long blocks = (n_vec + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;          // grid-stride over a multiple of the SM count
    k_add_bn_relu<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)y, alpha, ybias, scale, shift,
                                                                         (uint4*)sum_out, (uint4*)act_out, n_vec, channels / 8);
    return set_err((int)cudaGetLastError(), "tg_net_add_bn_relu");
}

// ------------------------------------------------------------------------------------------------------------------
// Fused spatial self-attention of the TransGo trunk in inference (model.py:286-314), one CTA per position:
//     S[i][j] = q_i . k_j          att = softmax_j(S)          o_j = sum_i att[i][j] * v_i
//     out = relu(bn(gamma * o + x))
// Eager PyTorch runs this as two batched GEMMs on [B,P,P] score tensors plus softmax, slicing copies and the junction
// (45 % of a 4096-position forward); here scores never leave the SM.  P <= 96 points (9x9), 32 query/key channels,
// 128 value channels.  Warp w owns key rows j in [16w, 16w+16): it forms S^T[j][i] for all i with mma.sync m16n8k16
// (bf16 in, f32 accumulate), the per-query max / sum over j go through shared memory across the six warps, and the
// normalised probabilities are re-used straight from the accumulator registers as the A operand of the second product
// o[j][c] = sum_i att^T[j][i] v[i][c], whose B operand comes from shared memory through ldmatrix.trans.
// (The problem is 2 MFLOP per position and bound by its 70 KB of HBM traffic, not by tensor throughput.)
namespace tg {

constexpr int kAttP = 96;          // padded points
constexpr int kAttDq = 32;
constexpr int kAttC = 128;
constexpr int kAttWarps = kAttP / 16;
constexpr int kQS = kAttDq + 8;    // smem row strides (bf16), padded against bank conflicts
constexpr int kVS = kAttC + 8;

__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kAttWarps * 32)
k_attention(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ x, float gamma,
            const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, int P) {
    __shared__ __align__(16) __nv_bfloat16 sQ[kAttP * kQS];
    __shared__ __align__(16) __nv_bfloat16 sK[kAttP * kQS];
    __shared__ __align__(16) __nv_bfloat16 sV[kAttP * kVS];
    __shared__ float sPart[kAttWarps][kAttP];
    __shared__ float sStat[kAttP];
    const int tid = threadIdx.x, lane = tid & 31, wj = tid >> 5, g = lane >> 2, t = lane & 3;
    const long long b = blockIdx.x;
    constexpr int ROWQ = (2 * kAttDq + kAttC) / 8;                   // 16-byte chunks per qkv row
    const uint4* src = reinterpret_cast<const uint4*>(qkv + b * (long long)P * (2 * kAttDq + kAttC));
    // asynchronous 16-byte copies straight into shared memory: all of a thread's requests are in flight at once
    for (int idx = tid; idx < kAttP * ROWQ; idx += blockDim.x) {
        const int row = idx / ROWQ, ch = idx - row * ROWQ;
        __nv_bfloat16* dst = ch < kAttDq / 8 ? &sQ[row * kQS + ch * 8]
                           : ch < 2 * kAttDq / 8 ? &sK[row * kQS + (ch - kAttDq / 8) * 8]
                                                 : &sV[row * kVS + (ch - 2 * kAttDq / 8) * 8];
        if (row < P) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src + idx));
        } else {
            *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);  // rows >= P are zero: they add nothing to o
        }
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();

    // ---- S^T[j][i] = k_j . q_i for this warp's 16 key rows j and every query i
    const int j0 = wj * 16;
    uint32_t ka[kAttDq / 16][4];
#pragma unroll
    for (int kc = 0; kc < kAttDq / 16; ++kc) {
        ka[kc][0] = *reinterpret_cast<const uint32_t*>(&sK[(j0 + g) * kQS + kc * 16 + 2 * t]);
        ka[kc][1] = *reinterpret_cast<const uint32_t*>(&sK[(j0 + g + 8) * kQS + kc * 16 + 2 * t]);
        ka[kc][2] = *reinterpret_cast<const uint32_t*>(&sK[(j0 + g) * kQS + kc * 16 + 8 + 2 * t]);
        ka[kc][3] = *reinterpret_cast<const uint32_t*>(&sK[(j0 + g + 8) * kQS + kc * 16 + 8 + 2 * t]);
    }
    float s[kAttP / 8][4];
#pragma unroll
    for (int nt = 0; nt < kAttP / 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
#pragma unroll
        for (int kc = 0; kc < kAttDq / 16; ++kc) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sQ[(nt * 8 + g) * kQS + kc * 16 + 2 * t]);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sQ[(nt * 8 + g) * kQS + kc * 16 + 8 + 2 * t]);
            mma_bf16(s[nt], ka[kc], b0, b1);
        }
    }
    // ---- softmax over the keys j (rows of S^T; they are spread over the warps) for every query i (columns)
    const bool r0 = j0 + g < P, r1 = j0 + g + 8 < P;
    const float NEG = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < kAttP / 8; ++nt) {
        float m0 = fmaxf(r0 ? s[nt][0] : NEG, r1 ? s[nt][2] : NEG), m1 = fmaxf(r0 ? s[nt][1] : NEG, r1 ? s[nt][3] : NEG);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o)); }
        if (g == 0) { sPart[wj][nt * 8 + 2 * t] = m0; sPart[wj][nt * 8 + 2 * t + 1] = m1; }
    }
    __syncthreads();
    if (tid < kAttP) {
        float m = sPart[0][tid];
#pragma unroll
        for (int w = 1; w < kAttWarps; ++w) m = fmaxf(m, sPart[w][tid]);
        sStat[tid] = m;
    }
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < kAttP / 8; ++nt) {
        const float mx0 = sStat[nt * 8 + 2 * t], mx1 = sStat[nt * 8 + 2 * t + 1];
        s[nt][0] = r0 ? __expf(s[nt][0] - mx0) : 0.0f; s[nt][1] = r0 ? __expf(s[nt][1] - mx1) : 0.0f;
        s[nt][2] = r1 ? __expf(s[nt][2] - mx0) : 0.0f; s[nt][3] = r1 ? __expf(s[nt][3] - mx1) : 0.0f;
        float z0 = s[nt][0] + s[nt][2], z1 = s[nt][1] + s[nt][3];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { z0 += __shfl_xor_sync(0xffffffffu, z0, o); z1 += __shfl_xor_sync(0xffffffffu, z1, o); }
        if (g == 0) { sPart[wj][nt * 8 + 2 * t] = z0; sPart[wj][nt * 8 + 2 * t + 1] = z1; }
    }
    __syncthreads();
    if (tid < kAttP) {
        float z = sPart[0][tid];
#pragma unroll
        for (int w = 1; w < kAttWarps; ++w) z += sPart[w][tid];
        sStat[tid] = 1.0f / z;                                       // z >= 1: the row that holds the max contributes exp(0)
    }
    __syncthreads();
    // probabilities as bf16 A fragments (two neighbouring 8-column accumulator tiles make one 16-wide k chunk)
    uint32_t pa[kAttP / 16][4];
#pragma unroll
    for (int kc = 0; kc < kAttP / 16; ++kc) {
        const float i00 = sStat[kc * 16 + 2 * t], i01 = sStat[kc * 16 + 2 * t + 1], i10 = sStat[kc * 16 + 8 + 2 * t], i11 = sStat[kc * 16 + 8 + 2 * t + 1];
        pa[kc][0] = pack_bf16(s[2 * kc][0] * i00, s[2 * kc][1] * i01);
        pa[kc][1] = pack_bf16(s[2 * kc][2] * i00, s[2 * kc][3] * i01);
        pa[kc][2] = pack_bf16(s[2 * kc + 1][0] * i10, s[2 * kc + 1][1] * i11);
        pa[kc][3] = pack_bf16(s[2 * kc + 1][2] * i10, s[2 * kc + 1][3] * i11);
    }
    // ---- o[j][c] = sum_i att^T[j][i] * v[i][c], 16 channels at a time, then the junction epilogue
    const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8, mcol = (lane >> 4) * 8;
    const long long rowbase = b * (long long)P;
#pragma unroll
    for (int ct = 0; ct < kAttC / 16; ++ct) {
        float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int kc = 0; kc < kAttP / 16; ++kc) {
            uint32_t v0, v1, v2, v3;
            const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&sV[(kc * 16 + mrow) * kVS + ct * 16 + mcol]);
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                         : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
            mma_bf16(o[0], pa[kc], v0, v1);
            mma_bf16(o[1], pa[kc], v2, v3);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = ct * 16 + h * 8 + 2 * t;
            const float sc0 = scale[c], sc1 = scale[c + 1], sh0 = shift[c], sh1 = shift[c + 1];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int j = j0 + g + rr * 8;
                if (j >= P) continue;
                const long long off = (rowbase + j) * kAttC + c;
                const float2 xv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + off));
                // same rounding points as the unfused path: o -> bf16, x + gamma * o -> bf16, BatchNorm + ReLU -> bf16
                const float2 ov = __bfloat1622float2(__floats2bfloat162_rn(o[h][2 * rr], o[h][2 * rr + 1]));
                const float2 sv = __bfloat1622float2(__floats2bfloat162_rn(xv.x + gamma * ov.x, xv.y + gamma * ov.y));
                *reinterpret_cast<__nv_bfloat162*>(out + off) =
                    __floats2bfloat162_rn(fmaxf(sv.x * sc0 + sh0, 0.0f), fmaxf(sv.y * sc1 + sh1, 0.0f));
            }
        }
    }
}

}  // namespace tg

extern "C" int tg_net_attention(const void* qkv, const void* x, float gamma, const float* scale, const float* shift, void* out,
                                int64_t B, int P, int dq, int channels, void* stream) {
    using namespace tg;
    if (!qkv || !x || !scale || !shift || !out || B <= 0 || P <= 0 || P > kAttP || dq != kAttDq || channels != kAttC)
        return set_err(-1, "tg_net_attention (supports P <= 96, 32 query channels, 128 value channels)");
    k_attention<<<(unsigned)B, kAttWarps * 32, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)x, gamma, scale,
                                                                           shift, (__nv_bfloat16*)out, P);
    return set_err((int)cudaGetLastError(), "tg_net_attention");
}
