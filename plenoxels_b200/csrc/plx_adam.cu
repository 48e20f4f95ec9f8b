// plx_adam.cu — K3: one Adam step over the whole grid fused with `grid_grad += |grad|` and the gradient clear.
//
// Stands under scripts/train.py:89 (Adam([grid], lr)), :180 (zero_grad), :182 (step), :184 (|grad| accumulation).
// Element arithmetic follows torch/optim/adam.py `_single_tensor_adam` (non-capturable branch) as ATen's CPU
// kernels evaluate it (oracle/plenoxel_oracle.py:adam_step pins the FMA placement against torch-CPU):
//   m  = fma(1-b1, g - m, m)                         exp_avg.lerp_(grad, 1 - beta1)
//   v  = fma((1-b2) * g, g, v * b2)                  exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
//   d  = sqrt(v) / sqrt(1 - b2^t) + eps              (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
//   p  = p + ((-lr / (1 - b1^t)) * m) / d            param.addcdiv_(exp_avg, denom, value=-step_size)
// Pure streaming: 5 reads + 5 writes of 16 bytes per cell = 160 B/cell, HBM-bound.
#include "plx_device.cuh"
#include "plx_launch.h"

namespace plx {

// `bc` = hoisted reciprocal of bc2_sqrt (plx_device.cuh): both quotients and the square root are the correctly rounded
// IEEE results, computed with the same expansions ptxas uses but without the per-element range-check branches.
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamScalars& s, const FastDiv& bc) {
    m = fmaf(s.one_minus_beta1, __fsub_rn(g, m), m);
    v = fmaf(__fmul_rn(s.one_minus_beta2, g), g, __fmul_rn(v, s.beta2));
    const float denom = __fadd_rn(fdiv_exact(fsqrt_exact(v), bc), s.eps);
    p = __fadd_rn(p, fdiv_var(__fmul_rn(s.neg_step_size, m), denom));
}

template <bool HAS_ABS, bool ZERO>
__global__ void __launch_bounds__(256) k_adam(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                              float4* __restrict__ v, float4* __restrict__ ga, int64_t n4,
                                              const AdamScalars s) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const FastDiv bc = make_fastdiv(s.bc2_sqrt);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = p[i];
        const float4 G = __ldcs(g + i);
        float4 M = __ldcs(m + i);
        float4 V = __ldcs(v + i);
        adam1(P.x, G.x, M.x, V.x, s, bc);
        adam1(P.y, G.y, M.y, V.y, s, bc);
        adam1(P.z, G.z, M.z, V.z, s, bc);
        adam1(P.w, G.w, M.w, V.w, s, bc);
        p[i] = P;                       // the grid is re-read by the next step's march: default (L2-resident) policy
        __stcs(m + i, M);
        __stcs(v + i, V);
        if (HAS_ABS) {
            float4 A = __ldcs(ga + i);
            A.x += fabsf(G.x); A.y += fabsf(G.y); A.z += fabsf(G.z); A.w += fabsf(G.w);
            __stcs(ga + i, A);
        }
        if (ZERO) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// scalar tail / unaligned fallback
template <bool HAS_ABS, bool ZERO>
__global__ void k_adam_scalar(float* p, float* g, float* m, float* v, float* ga, int64_t begin, int64_t n,
                              const AdamScalars s) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const FastDiv bc = make_fastdiv(s.bc2_sqrt);
    for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float P = p[i], M = m[i], V = v[i];
        const float G = g[i];
        adam1(P, G, M, V, s, bc);
        p[i] = P; m[i] = M; v[i] = V;
        if (HAS_ABS) ga[i] += fabsf(G);
        if (ZERO) g[i] = 0.f;
    }
}

cudaError_t launch_adam(float* p, float* g, float* m, float* v, float* gabs, int64_t n, const AdamScalars& s,
                        bool zero_grad, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const bool aligned = ((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                         ((uintptr_t)v % 16 == 0) && (!gabs || (uintptr_t)gabs % 16 == 0);
    const int64_t n4 = aligned ? n / 4 : 0;
    const int threads = 256;
#define PLX_ADAM(K, ...)                                                        \
    do {                                                                        \
        if (gabs) { if (zero_grad) K<true, true> __VA_ARGS__; else K<true, false> __VA_ARGS__; } \
        else      { if (zero_grad) K<false, true> __VA_ARGS__; else K<false, false> __VA_ARGS__; } \
    } while (0)
    if (n4 > 0) {
        // 148 SMs x 8 resident blocks of 256 threads; grid-stride over the rest
        int64_t want = (n4 + threads - 1) / threads;
        const unsigned blocks = (unsigned)(want < 148 * 8 ? want : 148 * 8);
        PLX_ADAM(k_adam, <<<blocks, threads, 0, st>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, (float4*)gabs, n4, s));
    }
    if (n4 * 4 < n) {
        const int64_t rem = n - n4 * 4;
        int64_t want = (rem + threads - 1) / threads;
        const unsigned blocks = (unsigned)(want < 148 * 8 ? want : 148 * 8);
        PLX_ADAM(k_adam_scalar, <<<blocks, threads, 0, st>>>(p, g, m, v, gabs, n4 * 4, n, s));
    }
#undef PLX_ADAM
    return cudaGetLastError();
}

}  // namespace plx
