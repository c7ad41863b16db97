// tg_net.cu -- the two memory-bound epilogues of the TransGo trunk in inference (model.py:208-262, :286-314), fused.
// The net itself stays in PyTorch / cuDNN / cuBLAS (north_star); after folding eval-mode BatchNorm into per-channel
// scale/shift what is left between the tensor-core kernels is elementwise traffic, and eager PyTorch spends ~70 % of a
// 4096-position forward on it (13 separate BN / ReLU / add kernels and layout copies).  These kernels do each residual
// junction in ONE pass over NHWC bf16 with 128-bit accesses:
//     sum = x + alpha * y            (residual add, or gamma * attention + x)
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
k_add_bn_relu(const uint4* __restrict__ x, const uint4* __restrict__ y, float alpha, const float* __restrict__ scale,
              const float* __restrict__ shift, uint4* __restrict__ sum_out, uint4* __restrict__ act_out, long long n_vec, int cvec) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % cvec) * 8;
        float a[8], b[8];
        unpack8(x[i], a);
        if (y) {
            unpack8(y[i], b);
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

extern "C" int tg_net_add_bn_relu(const void* x, const void* y, float alpha, const float* scale, const float* shift,
                                   void* sum_out, void* act_out, int64_t rows, int channels, void* stream) {
    using namespace tg;
    if (!x || rows <= 0 || channels <= 0 || (channels & 7) || (!sum_out && !act_out) || (act_out && (!scale || !shift)))
        return set_err(-1, "tg_net_add_bn_relu");
    const long long n_vec = rows * (long long)(channels / 8);
    long long blocks = (n_vec + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;          // grid-stride over a multiple of the SM count
    k_add_bn_relu<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)y, alpha, scale, shift,
                                                                         (uint4*)sum_out, (uint4*)act_out, n_vec, channels / 8);
    return set_err((int)cudaGetLastError(), "tg_net_add_bn_relu");
}
