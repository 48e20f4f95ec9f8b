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

__device__ __forceinline__ void step_tail(const StepTail& t) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (t.src && t.dst_host) {
            *reinterpret_cast<volatile float*>(t.dst_host) = *t.src;
            __threadfence_system();                                  // the loss is visible to the host before the step number
            *reinterpret_cast<volatile int32_t*>(t.dst_host + 1) = t.step;
        }
        if (t.clear) *t.clear = 0.f;
        if (t.counter_clear) *t.counter_clear = 0;
    }
}

template <bool HAS_ABS, bool ZERO, int UNROLL, bool SKIP>
__global__ void __launch_bounds__(256) k_adam(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                              float4* __restrict__ v, float4* __restrict__ ga, int64_t n4,
                                              const AdamScalars s, const StepTail tail) {
    step_tail(tail);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const FastDiv bc = make_fastdiv(s.bc2_sqrt);
    // parameters and gradient are re-used by the march of the next step: keep them in L2 when they fit (plx_device.cuh)
    const uint64_t pol = l2_policy(s.keep_p), pol_g = l2_policy(s.keep_g);
    const bool rev = s.reverse, cs = s.stream_state;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * UNROLL) {
        float4 P[UNROLL], G[UNROLL], M[UNROLL], V[UNROLL], A[UNROLL];
        // all loads of the iteration are issued before the first use: 5 * UNROLL independent 16-byte requests per thread
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t k = i0 + u * stride;
            if (k < n4) {
                const int64_t i = rev ? n4 - 1 - k : k;
                P[u] = ld_hint(p + i, pol);
                G[u] = ld_hint(g + i, pol_g);
                M[u] = cs ? __ldcs(m + i) : m[i];
                V[u] = cs ? __ldcs(v + i) : v[i];
                if (HAS_ABS) A[u] = cs ? __ldcs(ga + i) : ga[i];
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t k = i0 + u * stride;
            if (k < n4) {
                const int64_t i = rev ? n4 - 1 - k : k;
                adam1(P[u].x, G[u].x, M[u].x, V[u].x, s, bc);
                adam1(P[u].y, G[u].y, M[u].y, V[u].y, s, bc);
                adam1(P[u].z, G[u].z, M[u].z, V[u].z, s, bc);
                adam1(P[u].w, G[u].w, M[u].w, V[u].w, s, bc);
                st_hint(p + i, P[u], pol);
                if (cs) { __stcs(m + i, M[u]); __stcs(v + i, V[u]); } else { m[i] = M[u]; v[i] = V[u]; }
                // a cell no ray touched this step has g == 0 exactly: |g| adds nothing and the gradient is already clear,
                // so neither store is issued (about half the cells of a C2 step; saves their 32 B/cell of write-back)
                const bool touched = !SKIP || G[u].x != 0.f || G[u].y != 0.f || G[u].z != 0.f || G[u].w != 0.f;
                if (HAS_ABS && touched) {
                    A[u].x += fabsf(G[u].x); A[u].y += fabsf(G[u].y); A[u].z += fabsf(G[u].z); A[u].w += fabsf(G[u].w);
                    if (cs) __stcs(ga + i, A[u]); else ga[i] = A[u];
                }
                if (ZERO && touched) st_hint(g + i, make_float4(0.f, 0.f, 0.f, 0.f), pol_g);
            }
        }
    }
}

static int adam_env(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

// K3p: reduce-scatter + Adam + all-gather over peer-mapped memory in one pass (see plenoxel_abi.h)
struct PeerPtrs {
    float4* grids[PLX_MAX_PEERS];
    const float4* grads[PLX_MAX_PEERS];
};

template <int UNROLL>
__global__ void __launch_bounds__(256) k_adam_peer(PeerPtrs pp, int world, int world_st, float4* __restrict__ m, float4* __restrict__ v,
                                                   float4* __restrict__ ga, int64_t begin4, int64_t end4, const AdamScalars s,
                                                   const StepTail tail, const PlxPeerSync sync) {
    step_tail(tail);
    peer_wait(sync);                         // every rank's partial gradient is complete
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const FastDiv bc = make_fastdiv(s.bc2_sqrt);
    for (int64_t i0 = begin4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < end4; i0 += stride * UNROLL) {
        float4 G[UNROLL], P[UNROLL], M[UNROLL], V[UNROLL], A[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = i0 + u * stride;
            G[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < end4) {
                // all partial gradients requested before the first use: `world` independent 16-byte loads (world-1 over NVLink)
                float4 part[PLX_MAX_PEERS];
#pragma unroll
                for (int r = 0; r < PLX_MAX_PEERS; ++r)
                    if (r < world) part[r] = pp.grads[r][i];
                P[u] = pp.grids[0][i];           // slot 0 is the local replica (all replicas hold the same parameters)
                M[u] = __ldcs(m + i);
                V[u] = __ldcs(v + i);
                if (ga) A[u] = __ldcs(ga + i);
#pragma unroll
                for (int r = 0; r < PLX_MAX_PEERS; ++r)
                    if (r < world) { G[u].x += part[r].x; G[u].y += part[r].y; G[u].z += part[r].z; G[u].w += part[r].w; }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < end4) {
                adam1(P[u].x, G[u].x, M[u].x, V[u].x, s, bc);
                adam1(P[u].y, G[u].y, M[u].y, V[u].y, s, bc);
                adam1(P[u].z, G[u].z, M[u].z, V[u].z, s, bc);
                adam1(P[u].w, G[u].w, M[u].w, V[u].w, s, bc);
#pragma unroll
                for (int r = 0; r < PLX_MAX_PEERS; ++r)
                    if (r < world_st) pp.grids[r][i] = P[u];
                __stcs(m + i, M[u]);
                __stcs(v + i, V[u]);
                if (ga && (G[u].x != 0.f || G[u].y != 0.f || G[u].z != 0.f || G[u].w != 0.f)) {     // untouched cell: |g| adds nothing
                    A[u].x += fabsf(G[u].x); A[u].y += fabsf(G[u].y); A[u].z += fabsf(G[u].z); A[u].w += fabsf(G[u].w);
                    __stcs(ga + i, A[u]);
                }
            }
        }
    }
    if (sync.signal_epoch > 0) {             // this rank's slab is stored in every replica (and the peers' gradients are read)
        __syncthreads();
        if (threadIdx.x == 0) peer_signal(sync);
    }
}

// NVLS variant: in-switch reduction of the partial gradients and in-switch replication of the new parameters
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float4* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int UNROLL>
__global__ void __launch_bounds__(256) k_adam_mc(const float4* __restrict__ p_local, float4* p_mc, const float4* g_mc,
                                                 float4* __restrict__ m, float4* __restrict__ v, float4* __restrict__ ga,
                                                 int64_t begin4, int64_t end4, const AdamScalars s, const StepTail tail,
                                                 const PlxPeerSync sync) {
    step_tail(tail);
    peer_wait(sync);                         // every rank's partial gradient is complete
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const FastDiv bc = make_fastdiv(s.bc2_sqrt);
    for (int64_t i0 = begin4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < end4; i0 += stride * UNROLL) {
        float4 G[UNROLL], P[UNROLL], M[UNROLL], V[UNROLL], A[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < end4) {
                G[u] = multimem_ld_reduce_add(g_mc + i);
                P[u] = p_local[i];
                M[u] = __ldcs(m + i);
                V[u] = __ldcs(v + i);
                if (ga) A[u] = __ldcs(ga + i);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < end4) {
                adam1(P[u].x, G[u].x, M[u].x, V[u].x, s, bc);
                adam1(P[u].y, G[u].y, M[u].y, V[u].y, s, bc);
                adam1(P[u].z, G[u].z, M[u].z, V[u].z, s, bc);
                adam1(P[u].w, G[u].w, M[u].w, V[u].w, s, bc);
                multimem_st(p_mc + i, P[u]);
                __stcs(m + i, M[u]);
                __stcs(v + i, V[u]);
                if (ga && (G[u].x != 0.f || G[u].y != 0.f || G[u].z != 0.f || G[u].w != 0.f)) {     // untouched cell: |g| adds nothing
                    A[u].x += fabsf(G[u].x); A[u].y += fabsf(G[u].y); A[u].z += fabsf(G[u].z); A[u].w += fabsf(G[u].w);
                    __stcs(ga + i, A[u]);
                }
            }
        }
    }
    if (sync.signal_epoch > 0) {             // this rank's slab is stored in every replica (and the peers' gradients are read)
        __syncthreads();
        if (threadIdx.x == 0) peer_signal(sync);
    }
}

cudaError_t launch_adam_peer(const PlxAdamPeer& a, const AdamScalars& s, const StepTail& tail, cudaStream_t st) {
    const int64_t begin4 = a.begin / 4, end4 = a.end / 4;
    if (end4 <= begin4) return cudaSuccess;
    // in-switch reduction pays off once several peers would otherwise be read one by one; with 2 ranks it only adds a
    // round trip through the switch (measured: 93 vs 60 us at N=2, 69 vs 96 us at N=8).  PLX_PEER_MULTICAST=0/1 overrides.
    static const int mc_env = adam_env("PLX_PEER_MULTICAST", -1);
    const bool use_mc = mc_env >= 0 ? mc_env != 0 : a.world >= 4;
    if (a.grid_mc && a.grad_mc && use_mc) {
        int dev = 0, sms = 148, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        static const int unroll = adam_env("PLX_MC_UNROLL", 2);
        // few resident blocks on purpose: with the whole slab in flight at once the switch-reduced loads and the replicated
        // stores would run as two serial phases; a deeper grid-stride loop keeps both NVLink directions busy together
        static const int mc_cap = adam_env("PLX_MC_BLOCKS_PER_SM", 1);
        static const int mc_blocks = adam_env("PLX_MC_BLOCKS", 0);          // total CTA count override (tuning)
        const int64_t n4 = end4 - begin4;
#define PLX_MC(U)                                                                                                      \
        do {                                                                                                           \
            cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_adam_mc<U>, 256, 0);              \
            if (e != cudaSuccess) return e;                                                                            \
            if (mc_cap > 0 && per_sm > mc_cap) per_sm = mc_cap;                                                        \
            const int64_t want = (n4 + 256 * U - 1) / (256 * U);                                                       \
            const int64_t resident = (int64_t)sms * (per_sm > 0 ? per_sm : 1);                                         \
            unsigned blocks = (unsigned)(want < resident ? want : resident);                                           \
            if (mc_blocks > 0 && (unsigned)mc_blocks < blocks) blocks = (unsigned)mc_blocks;                           \
            k_adam_mc<U><<<blocks, 256, 0, st>>>((const float4*)a.grids[a.rank], (float4*)a.grid_mc, (const float4*)a.grad_mc,   \
                                                 (float4*)a.exp_avg, (float4*)a.exp_avg_sq, (float4*)a.grad_abs_sum, begin4, end4, s, tail, a.sync); \
        } while (0)
        if (unroll >= 4) PLX_MC(4); else if (unroll >= 2) PLX_MC(2); else PLX_MC(1);
#undef PLX_MC
        return cudaGetLastError();
    }
    PeerPtrs pp;
    // put the local replica first so that the parameter read (grids[0]) is local
    int order[PLX_MAX_PEERS];
    for (int r = 0; r < a.world; ++r) order[r] = (a.rank + r) % a.world;
    for (int r = 0; r < PLX_MAX_PEERS; ++r) {
        pp.grids[r] = r < a.world ? (float4*)a.grids[order[r]] : nullptr;
        pp.grads[r] = r < a.world ? (const float4*)a.grads[order[r]] : nullptr;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n4 = end4 - begin4;
    // remote loads have microseconds of latency: keep `world * UNROLL` 16-byte requests in flight per thread
    static const int unroll_env = adam_env("PLX_PEER_UNROLL", 0);
    const int unroll = unroll_env ? unroll_env : 1;
#define PLX_PEER(U)                                                                                                  \
    do {                                                                                                             \
        int per_sm = 0;                                                                                              \
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_adam_peer<U>, 256, 0);              \
        if (e != cudaSuccess) return e;                                                                              \
        const int64_t want = (n4 + 256 * U - 1) / (256 * U);                                                         \
        const int64_t resident = (int64_t)sms * (per_sm > 0 ? per_sm : 1);                                           \
        const unsigned blocks = (unsigned)(want < resident ? want : resident);                                       \
        k_adam_peer<U><<<blocks, 256, 0, st>>>(pp, a.world, a.world, (float4*)a.exp_avg, (float4*)a.exp_avg_sq,                                   \
                                               (float4*)a.grad_abs_sum, begin4, end4, s, tail, a.sync);              \
    } while (0)
    if (unroll >= 4) PLX_PEER(4); else if (unroll >= 2) PLX_PEER(2); else PLX_PEER(1);
#undef PLX_PEER
    return cudaGetLastError();
}

// cross-GPU barrier over peer-mapped flag arrays (see plenoxel_abi.h)
struct FlagPtrs { int32_t* p[PLX_MAX_PEERS]; };

__global__ void k_peer_barrier(FlagPtrs f, int rank, int world, int channel, int epoch) {
    const int r = threadIdx.x;
    if (r < world) {
        int32_t* theirs = f.p[r] + channel * PLX_MAX_PEERS + rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
        const int32_t* mine = f.p[rank] + channel * PLX_MAX_PEERS + r;
        // bounded spin: a peer that died must not hang this GPU (a hung box costs far more than a wrong step); ~10 s
        int32_t seen;
        long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        } while (seen - epoch < 0 && ++spins < (1ll << 23));
    }
}

cudaError_t launch_peer_barrier(int32_t* const* flags, int rank, int world, int channel, int epoch, cudaStream_t st) {
    FlagPtrs f;
    for (int r = 0; r < PLX_MAX_PEERS; ++r) f.p[r] = r < world ? flags[r] : nullptr;
    k_peer_barrier<<<1, 32, 0, st>>>(f, rank, world, channel, epoch);
    return cudaGetLastError();
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

template <bool HAS_ABS, bool ZERO, int UNROLL, bool SKIP>
static cudaError_t launch_adam_vec2(float4* p, float4* g, float4* m, float4* v, float4* ga, int64_t n4, const AdamScalars& s,
                                    int blocks_per_sm_cap, const StepTail& tail, cudaStream_t st) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_adam<HAS_ABS, ZERO, UNROLL, SKIP>, 256, 0);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (blocks_per_sm_cap > 0 && per_sm > blocks_per_sm_cap) per_sm = blocks_per_sm_cap;
    // one wave of resident blocks, grid-stride over the rest (no tail wave)
    int64_t want = (n4 + 256 * UNROLL - 1) / (256 * UNROLL);
    const int64_t resident = (int64_t)sms * (per_sm > 0 ? per_sm : 1);
    const unsigned blocks = (unsigned)(want < resident ? want : resident);
    k_adam<HAS_ABS, ZERO, UNROLL, SKIP><<<blocks, 256, 0, st>>>(p, g, m, v, ga, n4, s, tail);
    return cudaGetLastError();
}

template <bool HAS_ABS, bool ZERO, int UNROLL>
static cudaError_t launch_adam_vec(float4* p, float4* g, float4* m, float4* v, float4* ga, int64_t n4, const AdamScalars& s,
                                   int blocks_per_sm_cap, const StepTail& tail, cudaStream_t st) {
    // skipping the |g| / clear stores of untouched cells only matters where either store exists
    static const bool skip = adam_env("PLX_ADAM_SKIP", 1) != 0;
    if ((HAS_ABS || ZERO) && skip) return launch_adam_vec2<HAS_ABS, ZERO, UNROLL, true>(p, g, m, v, ga, n4, s, blocks_per_sm_cap, tail, st);
    return launch_adam_vec2<HAS_ABS, ZERO, UNROLL, false>(p, g, m, v, ga, n4, s, blocks_per_sm_cap, tail, st);
}

__global__ void k_step_tail_only(const StepTail tail) { step_tail(tail); }

cudaError_t launch_adam(float* p, float* g, float* m, float* v, float* gabs, int64_t n, const AdamScalars& s_in,
                        bool zero_grad, const StepTail& tail, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    AdamScalars s = s_in;
    {
        static const int mode = adam_env("PLX_L2_KEEP", 2);
        const bool fits = l2_keep_ok(n / 4);
        s.keep_p = fits && mode != 0;
        s.keep_g = fits && mode == 1;
        static const int pingpong = adam_env("PLX_ADAM_PINGPONG", 1), cs = adam_env("PLX_ADAM_CS", 1);
        s.reverse = s.reverse && pingpong != 0;
        s.stream_state = cs != 0;
    }
    const bool aligned = ((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                         ((uintptr_t)v % 16 == 0) && (!gabs || (uintptr_t)gabs % 16 == 0);
    const int64_t n4 = aligned ? n / 4 : 0;
    const int threads = 256;
    static const int unroll = adam_env("PLX_ADAM_UNROLL", 1);
    // resident 256-thread blocks per SM, measured with the store skipping / alternating walk in place: 128^3 4 -> 88.6 us per
    // step, 5 / 6 -> 88.5 (flat); 256^3 3 -> 424 us, 4 -> 410, 5 -> 410
    static const int cap_env = adam_env("PLX_ADAM_BLOCKS_PER_SM", 0);
    const int cap = cap_env > 0 ? cap_env : 4;
    if (n4 > 0) {
        cudaError_t e;
#define PLX_ADAM_V(U)                                                                                                         \
        do {                                                                                                                  \
            if (gabs) { e = zero_grad ? launch_adam_vec<true, true, U>((float4*)p, (float4*)g, (float4*)m, (float4*)v, (float4*)gabs, n4, s, cap, tail, st)   \
                                      : launch_adam_vec<true, false, U>((float4*)p, (float4*)g, (float4*)m, (float4*)v, (float4*)gabs, n4, s, cap, tail, st); } \
            else      { e = zero_grad ? launch_adam_vec<false, true, U>((float4*)p, (float4*)g, (float4*)m, (float4*)v, nullptr, n4, s, cap, tail, st)        \
                                      : launch_adam_vec<false, false, U>((float4*)p, (float4*)g, (float4*)m, (float4*)v, nullptr, n4, s, cap, tail, st); }    \
        } while (0)
        if (unroll == 1) PLX_ADAM_V(1); else if (unroll == 4) PLX_ADAM_V(4); else PLX_ADAM_V(2);
#undef PLX_ADAM_V
        if (e != cudaSuccess) return e;
    }
    if (n4 == 0 && (tail.src || tail.clear || tail.counter_clear)) k_step_tail_only<<<1, 32, 0, st>>>(tail);
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
