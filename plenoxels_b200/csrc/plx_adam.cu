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
#include <cstdlib>

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

template <bool HAS_ABS, bool ZERO, int UNROLL>
__global__ void __launch_bounds__(256) k_adam(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                              float4* __restrict__ v, float4* __restrict__ ga, int64_t n4,
                                              const AdamScalars s) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const FastDiv bc = make_fastdiv(s.bc2_sqrt);
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * UNROLL) {
        float4 P[UNROLL], G[UNROLL], M[UNROLL], V[UNROLL], A[UNROLL];
        // all loads of the iteration are issued before the first use: 5 * UNROLL independent 16-byte requests per thread
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n4) {
                P[u] = p[i];
                G[u] = __ldcs(g + i);
                M[u] = __ldcs(m + i);
                V[u] = __ldcs(v + i);
                if (HAS_ABS) A[u] = __ldcs(ga + i);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n4) {
                adam1(P[u].x, G[u].x, M[u].x, V[u].x, s, bc);
                adam1(P[u].y, G[u].y, M[u].y, V[u].y, s, bc);
                adam1(P[u].z, G[u].z, M[u].z, V[u].z, s, bc);
                adam1(P[u].w, G[u].w, M[u].w, V[u].w, s, bc);
                p[i] = P[u];                // the grid is re-read by the next step's march: default (L2-resident) policy
                __stcs(m + i, M[u]);
                __stcs(v + i, V[u]);
                if (HAS_ABS) {
                    A[u].x += fabsf(G[u].x); A[u].y += fabsf(G[u].y); A[u].z += fabsf(G[u].z); A[u].w += fabsf(G[u].w);
                    __stcs(ga + i, A[u]);
                }
                if (ZERO) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
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

template <bool HAS_ABS, bool ZERO, int UNROLL>
static cudaError_t launch_adam_vec(float4* p, float4* g, float4* m, float4* v, float4* ga, int64_t n4, const AdamScalars& s,
                                   int blocks_per_sm_cap, cudaStream_t st) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_adam<HAS_ABS, ZERO, UNROLL>, 256, 0);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (blocks_per_sm_cap > 0 && per_sm > blocks_per_sm_cap) per_sm = blocks_per_sm_cap;
    // one wave of resident blocks, grid-stride over the rest (no tail wave)
    int64_t want = (n4 + 256 * UNROLL - 1) / (256 * UNROLL);
    const int64_t resident = (int64_t)sms * (per_sm > 0 ? per_sm : 1);
    const unsigned blocks = (unsigned)(want < resident ? want : resident);
    k_adam<HAS_ABS, ZERO, UNROLL><<<blocks, 256, 0, st>>>(p, g, m, v, ga, n4, s);
    return cudaGetLastError();
}

static int adam_env(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

cudaError_t launch_adam(float* p, float* g, float* m, float* v, float* gabs, int64_t n, const AdamScalars& s,
                        bool zero_grad, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const bool aligned = ((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                         ((uintptr_t)v % 16 == 0) && (!gabs || (uintptr_t)gabs % 16 == 0);
    const int64_t n4 = aligned ? n / 4 : 0;
    const int threads = 256;
    static const int unroll = adam_env("PLX_ADAM_UNROLL", 1);
    static const int cap = adam_env("PLX_ADAM_BLOCKS_PER_SM", 4);
    if (n4 > 0) {
        cudaError_t e;
#define PLX_ADAM_V(U)                                                                                                         \
        do {                                                                                                                  \
            if (gabs) { e = zero_grad ? launch_adam_vec<true, true, U>((float4*)p, (float4*)g, (float4*)m, (float4*)v, (float4*)gabs, n4, s, cap, st)   \
                                      : launch_adam_vec<true, false, U>((float4*)p, (float4*)g, (float4*)m, (float4*)v, (float4*)gabs, n4, s, cap, st); } \
            else      { e = zero_grad ? launch_adam_vec<false, true, U>((float4*)p, (float4*)g, (float4*)m, (float4*)v, nullptr, n4, s, cap, st)        \
                                      : launch_adam_vec<false, false, U>((float4*)p, (float4*)g, (float4*)m, (float4*)v, nullptr, n4, s, cap, st); }    \
        } while (0)
        if (unroll == 1) PLX_ADAM_V(1); else if (unroll == 4) PLX_ADAM_V(4); else PLX_ADAM_V(2);
#undef PLX_ADAM_V
        if (e != cudaSuccess) return e;
    }
    if (n4 * 4 < n) {
        const int64_t rem = n - n4 * 4;
        int64_t want = (rem + threads - 1) / threads;
        const unsigned blocks = (unsigned)(want < 148 * 8 ? want : 148 * 8);
        if (gabs) { if (zero_grad) k_adam_scalar<true, true><<<blocks, threads, 0, st>>>(p, g, m, v, gabs, n4 * 4, n, s);
                    else           k_adam_scalar<true, false><<<blocks, threads, 0, st>>>(p, g, m, v, gabs, n4 * 4, n, s); }
        else      { if (zero_grad) k_adam_scalar<false, true><<<blocks, threads, 0, st>>>(p, g, m, v, gabs, n4 * 4, n, s);
                    else           k_adam_scalar<false, false><<<blocks, threads, 0, st>>>(p, g, m, v, gabs, n4 * 4, n, s); }
    }
    return cudaGetLastError();
}

}  // namespace plx
