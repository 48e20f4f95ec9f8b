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
        float loss = 0.f;
        bool have = false;
        if (t.n_peers > 0) {
            float part[PLX_MAX_PEERS];
#pragma unroll
            for (int r = 0; r < PLX_MAX_PEERS; ++r)                  // independent (remote) loads, all in flight together
                part[r] = r < t.n_peers ? *reinterpret_cast<const volatile float*>(t.src_peers[r]) : 0.f;
#pragma unroll
            for (int r = 0; r < PLX_MAX_PEERS; ++r) loss += part[r];
            have = true;
            if (t.global_out) *t.global_out = loss;
        } else if (t.src) {
            loss = *t.src;
            have = true;
        }
        if (have && t.dst_host) {
            *reinterpret_cast<volatile float*>(t.dst_host) = loss;
            __threadfence_system();                                  // the loss is visible to the host before the step number
            *reinterpret_cast<volatile int32_t*>(t.dst_host + 1) = t.step;
        }
        if (t.clear) *t.clear = 0.f;
    }
}

__device__ __forceinline__ bool differs(const float4 a, const float4 b) { return a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w; }

// Store skipping is decided per 128-byte LINE (8 consecutive float4 = 8 consecutive lanes of the warp), never per thread: a line
// of which only some 16-byte pieces are written reaches DRAM as partial-sector writes, which HBM serves as read-modify-write
// (measured: per-thread skipping made the 256^3 step 15 % SLOWER, 402 -> 462 us).  True if any lane of this lane's line votes.
__device__ __forceinline__ bool line_any(bool vote) {
    const unsigned b = __ballot_sync(FULL, vote);
    return ((b >> ((threadIdx.x & 31u) & ~7u)) & 0xffu) != 0u;
}

// REPLAY: the graph-replay variant (PlxReplayState) is its own instantiation, so the plain kernel keeps its 4 blocks per SM
template <bool HAS_ABS, bool ZERO, bool SKIP, bool REPLAY>
__global__ void __launch_bounds__(256, 4) k_adam(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                 float4* __restrict__ v, float4* __restrict__ ga, int64_t n4,
                                                 AdamScalars s, StepTail tail, const ReplayArgs rp) {
    grid_dependency_wait();                  // launched behind the march's tail: its gradient and loss are complete from here on
    if (REPLAY) {
        // graph replay: this step's number, loss slots and bias-corrected scalars live in device memory (PlxReplayState)
        const int32_t step = *reinterpret_cast<const volatile int32_t*>(rp.step_dev) + 1;
        const int64_t row = (int64_t)step - rp.table_base - 1;
        if (step <= 0 || row < 0 || row >= rp.table_len) {           // table exhausted (or an earlier failure): apply nothing, say so
            if (blockIdx.x == 0 && threadIdx.x == 0) *rp.step_dev = -1;
            return;
        }
        s.bc2_sqrt = __ldg(rp.table + 2 * row);
        s.neg_step_size = __ldg(rp.table + 2 * row + 1);
        s.reverse = (step & 1) != 0;
        tail.src = rp.loss2 + (step & 1);
        tail.clear = rp.loss2 + ((step + 1) & 1);
        tail.step = step;
    }
    step_tail(tail);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const FastDiv bc = make_fastdiv(s.bc2_sqrt);
    // parameters and gradient are re-used by the march of the next step: keep them in L2 when they fit (plx_device.cuh)
    const uint64_t pol = l2_policy(s.keep_p), pol_g = l2_policy(s.keep_g);
    const bool rev = s.reverse;
    // the trip count is uniform over the warp (the line votes below need every lane present); lanes past the end idle
    for (int64_t k0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); k0 < n4; k0 += stride) {
        const int64_t k = k0 + (threadIdx.x & 31u);
        const bool live = k < n4;
        const int64_t i = live ? (rev ? n4 - 1 - k : k) : 0;
        // all loads of the iteration are issued before the first use: 5 independent 16-byte requests per thread; m, v and
        // |g| are touched once per step (evict-first), parameters and gradient again by the next march
        float4 P = ld_hint(p + i, pol), G = ld_hint(g + i, pol_g), M = __ldcs(m + i), V = __ldcs(v + i), A;
        if (HAS_ABS) A = __ldcs(ga + i);
        const float4 P0 = P, M0 = M, V0 = V;
        adam1(P.x, G.x, M.x, V.x, s, bc);
        adam1(P.y, G.y, M.y, V.y, s, bc);
        adam1(P.z, G.z, M.z, V.z, s, bc);
        adam1(P.w, G.w, M.w, V.w, s, bc);
        // a value that did not change is not stored: a cell no ray has EVER touched keeps m = v = 0 and its parameter, and a cell
        // whose update has decayed below half an ulp keeps its parameter — their lines stay clean in L2 and are never written back
        const bool all = !s.skip_same;
        const bool st_p = all || line_any(live && differs(P, P0)), st_m = all || line_any(live && differs(M, M0)),
                   st_v = all || line_any(live && differs(V, V0));
        // a cell no ray touched this step has g == 0 exactly: |g| adds nothing and the gradient is already clear,
        // so (when that holds for its whole line) neither store is issued — about half the lines of a C2 step
        const bool touched = !SKIP || line_any(live && (G.x != 0.f || G.y != 0.f || G.z != 0.f || G.w != 0.f));
        if (!live) continue;
        if (st_p) st_hint(p + i, P, pol);
        if (st_m) __stcs(m + i, M);
        if (st_v) __stcs(v + i, V);
        if (HAS_ABS && touched) {
            A.x += fabsf(G.x); A.y += fabsf(G.y); A.z += fabsf(G.z); A.w += fabsf(G.w);
            __stcs(ga + i, A);
        }
        if (ZERO && touched) st_hint(g + i, make_float4(0.f, 0.f, 0.f, 0.f), pol_g);
    }
    if (REPLAY) {                            // the last block to finish publishes the new step number (every block has read the old one)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicInc(reinterpret_cast<unsigned*>(rp.block_counter), gridDim.x - 1) == gridDim.x - 1) *rp.step_dev = tail.step;
        }
    }
}

// NVLS: in-switch reduction of the partial gradients / in-switch replication of the new parameters
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

// ---------------------------------------------------------------------------------------------------------------------
// K3s — optimiser step of the push exchange (PlxAdamSlab): k_adam over the owned slab, parameters stored to every replica.
// All loads are local HBM / L2 (the march already reduced every rank's contribution into this rank's gradient slab), all
// remote traffic is fire-and-forget stores: nothing in the loop waits on an NVLink round trip.
// ---------------------------------------------------------------------------------------------------------------------
struct SlabPtrs {
    float4* grids[PLX_MAX_PEERS];          // [0] = local replica, then the peers in ring order
};

template <bool MC>
__global__ void __launch_bounds__(256) k_adam_slab(const SlabPtrs sp, float4* p_mc, int world, float4* __restrict__ g,
                                                   float4* __restrict__ m, float4* __restrict__ v, float4* __restrict__ ga,
                                                   int64_t begin4, int64_t end4, const AdamScalars s, const StepTail tail,
                                                   const PlxPeerError err) {
    grid_dependency_wait();                  // launched behind the barrier kernel's tail (programmatic dependent launch)
    if (peer_failed(err)) return;            // an earlier wait gave up: the gradient slab may be incomplete — store nothing
    step_tail(tail);                         // global loss: every rank's partial is complete since the barrier before this kernel
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const FastDiv bc = make_fastdiv(s.bc2_sqrt);
    const int64_t n4 = end4 - begin4;
    const bool rev = s.reverse;
    float4* const p = sp.grids[0];
    for (int64_t k0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); k0 < n4; k0 += stride) {      // uniform trip count, see k_adam
        const int64_t k = k0 + (threadIdx.x & 31u);
        const bool live = k < n4;
        const int64_t i = begin4 + (live ? (rev ? n4 - 1 - k : k) : 0);
        float4 P = p[i], G = g[i], M = __ldcs(m + i), V = __ldcs(v + i), A;
        if (ga) A = __ldcs(ga + i);
        const float4 P0 = P, M0 = M, V0 = V;
        adam1(P.x, G.x, M.x, V.x, s, bc);
        adam1(P.y, G.y, M.y, V.y, s, bc);
        adam1(P.z, G.z, M.z, V.z, s, bc);
        adam1(P.w, G.w, M.w, V.w, s, bc);
        // replicas hold identical bits, so a parameter that did not change (a cell never touched, or an update below half an ulp)
        // need not cross NVLink at all; the same goes for untouched moments and the local HBM
        const bool all = !s.skip_same;
        const bool st_p = all || line_any(live && differs(P, P0)), st_m = all || line_any(live && differs(M, M0)),
                   st_v = all || line_any(live && differs(V, V0));
        const bool touched = line_any(live && (G.x != 0.f || G.y != 0.f || G.z != 0.f || G.w != 0.f));
        if (!live) continue;
        if (st_p) {
            if (MC) {
                multimem_st(p_mc + i, P);                            // one store, replicated into every replica by the switch
            } else {
#pragma unroll
                for (int r = 0; r < PLX_MAX_PEERS; ++r)
                    if (r < world) sp.grids[r][i] = P;
            }
        }
        if (st_m) __stcs(m + i, M);
        if (st_v) __stcs(v + i, V);
        if (touched) {                                               // untouched line: |g| adds nothing, gradient already clear
            if (ga) { A.x += fabsf(G.x); A.y += fabsf(G.y); A.z += fabsf(G.z); A.w += fabsf(G.w); __stcs(ga + i, A); }
            g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

static int resident_blocks(const void* kernel, int cap_per_sm, int64_t want) {
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (cap_per_sm > 0 && per_sm > cap_per_sm) per_sm = cap_per_sm;
    const int64_t resident = (int64_t)sms * per_sm;
    return (int)(want < resident ? (want < 1 ? 1 : want) : resident);
}

cudaError_t launch_adam_slab(const PlxAdamSlab& a, const AdamScalars& s_in, cudaStream_t st) {
    const int64_t begin4 = a.begin / 4, end4 = a.end / 4;
    AdamScalars s = s_in;
    s.skip_same = tuning().adam_skip_same != 0;          // auto = on
    SlabPtrs sp;
    for (int r = 0; r < PLX_MAX_PEERS; ++r)
        sp.grids[r] = r < a.world ? (float4*)a.grids[(a.rank + r) % a.world] : nullptr;     // local replica first
    StepTail tail{nullptr, a.loss_clear, (float*)a.result_host, (int32_t)a.step};
    if (a.loss_peers[0]) {
        for (int r = 0; r < a.world; ++r) tail.src_peers[r] = a.loss_peers[r];              // rank order: same sum on every rank
        tail.n_peers = a.world;
        tail.global_out = a.loss_out;
    }
    const int64_t want = (end4 - begin4 + 255) / 256;
    const bool mc = a.grid_mc != nullptr;
    // 4 resident blocks per SM like the single-GPU optimiser; an empty slab still runs one block for the step tail
    const int blocks = mc ? resident_blocks((const void*)k_adam_slab<true>, tuning().adam_blocks_per_sm, want) : resident_blocks((const void*)k_adam_slab<false>, tuning().adam_blocks_per_sm, want);
    if (mc) return launch_pdl(k_adam_slab<true>, (unsigned)blocks, 256u, 0, st, sp, (float4*)a.grid_mc, (int)a.world, (float4*)a.grad, (float4*)a.exp_avg,
                              (float4*)a.exp_avg_sq, (float4*)a.grad_abs_sum, begin4, end4, s, tail, a.err);
    return launch_pdl(k_adam_slab<false>, (unsigned)blocks, 256u, 0, st, sp, (float4*)nullptr, (int)a.world, (float4*)a.grad, (float4*)a.exp_avg,
                      (float4*)a.exp_avg_sq, (float4*)a.grad_abs_sum, begin4, end4, s, tail, a.err);
}

// ---------------------------------------------------------------------------------------------------------------------
// K3p — "pull" exchange: reduce-scatter + Adam + all-gather over peer-mapped memory in one pass (see plenoxel_abi.h)
// ---------------------------------------------------------------------------------------------------------------------
struct PeerPtrs {
    float4* grids[PLX_MAX_PEERS];
    const float4* grads[PLX_MAX_PEERS];
};

__global__ void __launch_bounds__(256) k_adam_peer(PeerPtrs pp, int world, float4* __restrict__ m, float4* __restrict__ v,
                                                   float4* __restrict__ ga, int64_t begin4, int64_t end4, const AdamScalars s,
                                                   const StepTail tail, const PlxPeerSync sync) {
    grid_dependency_wait();
    peer_wait(sync);                         // every rank's partial gradient (and loss) is complete
    __syncthreads();
    if (peer_failed(sync.err)) return;       // a wait gave up (here or in an earlier kernel): store nothing
    step_tail(tail);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const FastDiv bc = make_fastdiv(s.bc2_sqrt);
    for (int64_t i = begin4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end4; i += stride) {
        // all partial gradients requested before the first use: `world` independent 16-byte loads (world-1 over NVLink)
        float4 part[PLX_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < PLX_MAX_PEERS; ++r)
            if (r < world) part[r] = pp.grads[r][i];
        float4 P = pp.grids[0][i];           // slot 0 is the local replica (all replicas hold the same parameters)
        float4 M = __ldcs(m + i), V = __ldcs(v + i), A, G = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ga) A = __ldcs(ga + i);
#pragma unroll
        for (int r = 0; r < PLX_MAX_PEERS; ++r)
            if (r < world) { G.x += part[r].x; G.y += part[r].y; G.z += part[r].z; G.w += part[r].w; }
        adam1(P.x, G.x, M.x, V.x, s, bc);
        adam1(P.y, G.y, M.y, V.y, s, bc);
        adam1(P.z, G.z, M.z, V.z, s, bc);
        adam1(P.w, G.w, M.w, V.w, s, bc);
#pragma unroll
        for (int r = 0; r < PLX_MAX_PEERS; ++r)
            if (r < world) pp.grids[r][i] = P;
        __stcs(m + i, M);
        __stcs(v + i, V);
        if (ga && (G.x != 0.f || G.y != 0.f || G.z != 0.f || G.w != 0.f)) {     // untouched cell: |g| adds nothing
            A.x += fabsf(G.x); A.y += fabsf(G.y); A.z += fabsf(G.z); A.w += fabsf(G.w);
            __stcs(ga + i, A);
        }
    }
    if (sync.signal_epoch > 0) {             // this rank's slab is stored in every replica (and the peers' gradients are read)
        __syncthreads();
        if (threadIdx.x == 0) peer_signal(sync);
    }
}

// NVLS variant: two elements per thread in flight (measured best at N = 8: the switch-reduced loads and the replicated
// stores then overlap instead of running as two phases), one resident block per SM
__global__ void __launch_bounds__(256) k_adam_mc(const float4* __restrict__ p_local, float4* p_mc, const float4* g_mc,
                                                 float4* __restrict__ m, float4* __restrict__ v, float4* __restrict__ ga,
                                                 int64_t begin4, int64_t end4, const AdamScalars s, const StepTail tail,
                                                 const PlxPeerSync sync) {
    constexpr int UNROLL = 2;
    grid_dependency_wait();
    peer_wait(sync);                         // every rank's partial gradient (and loss) is complete
    __syncthreads();
    if (peer_failed(sync.err)) return;
    step_tail(tail);
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

// The caller chooses the variant: NVLS when it passes the multicast mappings (grid_mc, grad_mc), per-peer pointers otherwise.
// In-switch reduction pays off once several peers would otherwise be read one by one; with 2 ranks it only adds a round trip
// through the switch (measured: 93 vs 60 us at N=2, 69 vs 96 us at N=8) — the trainer passes the mappings from N = 4 up.
cudaError_t launch_adam_peer(const PlxAdamPeer& a, const AdamScalars& s, const StepTail& tail, cudaStream_t st) {
    const int64_t begin4 = a.begin / 4, end4 = a.end / 4;
    if (end4 <= begin4) return cudaSuccess;
    const int64_t n4 = end4 - begin4;
    if (a.grid_mc && a.grad_mc) {
        const int blocks = resident_blocks((const void*)k_adam_mc, 1, (n4 + 511) / 512);
        return launch_pdl(k_adam_mc, (unsigned)blocks, 256u, 0, st, (const float4*)a.grids[a.rank], (float4*)a.grid_mc, (const float4*)a.grad_mc,
                          (float4*)a.exp_avg, (float4*)a.exp_avg_sq, (float4*)a.grad_abs_sum, begin4, end4, s, tail, a.sync);
    }
    PeerPtrs pp;
    for (int r = 0; r < PLX_MAX_PEERS; ++r) {          // the local replica first, so that the parameter read (grids[0]) is local
        pp.grids[r] = r < a.world ? (float4*)a.grids[(a.rank + r) % a.world] : nullptr;
        pp.grads[r] = r < a.world ? (const float4*)a.grads[(a.rank + r) % a.world] : nullptr;
    }
    const int blocks = resident_blocks((const void*)k_adam_peer, 0, (n4 + 255) / 256);
    return launch_pdl(k_adam_peer, (unsigned)blocks, 256u, 0, st, pp, (int)a.world, (float4*)a.exp_avg, (float4*)a.exp_avg_sq, (float4*)a.grad_abs_sum,
                      begin4, end4, s, tail, a.sync);
}

// cross-GPU barrier over peer-mapped flag arrays (see plenoxel_abi.h)
struct FlagPtrs { int32_t* p[PLX_MAX_PEERS]; };

__global__ void k_peer_barrier(FlagPtrs f, int rank, int world, int channel, int epoch, const PlxPeerError err) {
    grid_dependency_wait();                  // everything this rank enqueued before the barrier is complete and visible
    const int r = threadIdx.x;
    if (r < world) {
        int32_t* theirs = f.p[r] + channel * PLX_MAX_PEERS + rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
        // bounded wait: a peer that died must not hang this GPU; giving up is recorded, never silent (plx_device.cuh)
        if (!spin_until(f.p[rank] + channel * PLX_MAX_PEERS + r, epoch, err.timeout_ns)) peer_fail(err, channel, epoch);
    }
}

cudaError_t launch_peer_barrier(int32_t* const* flags, int rank, int world, int channel, int epoch, const PlxPeerError& err,
                                cudaStream_t st) {
    FlagPtrs f;
    for (int r = 0; r < PLX_MAX_PEERS; ++r) f.p[r] = r < world ? flags[r] : nullptr;
    return launch_pdl(k_peer_barrier, 1u, 32u, 0, st, f, rank, world, channel, epoch, err);
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

template <bool HAS_ABS, bool ZERO>
static cudaError_t launch_adam_vec(float4* p, float4* g, float4* m, float4* v, float4* ga, int64_t n4, const AdamScalars& s,
                                   const StepTail& tail, cudaStream_t st, const ReplayArgs& rp) {
    // skipping the |g| / clear stores of untouched cells only matters where either store exists.  One wave of 4 resident
    // 256-thread blocks per SM, grid-stride over the rest (measured: 128^3 4 -> 88.6 us per step, 5 / 6 flat; 256^3 3 -> 424 us,
    // 4 -> 410, 5 -> 410)
    constexpr bool SKIP = HAS_ABS || ZERO;
    if (rp.step_dev) {
        const int blocks = resident_blocks((const void*)k_adam<HAS_ABS, ZERO, SKIP, true>, tuning().adam_blocks_per_sm, (n4 + 255) / 256);
        return launch_pdl(k_adam<HAS_ABS, ZERO, SKIP, true>, (unsigned)blocks, 256u, 0, st, p, g, m, v, ga, n4, s, tail, rp);
    }
    const int blocks = resident_blocks((const void*)k_adam<HAS_ABS, ZERO, SKIP, false>, tuning().adam_blocks_per_sm, (n4 + 255) / 256);
    return launch_pdl(k_adam<HAS_ABS, ZERO, SKIP, false>, (unsigned)blocks, 256u, 0, st, p, g, m, v, ga, n4, s, tail, rp);
}

__global__ void k_step_tail_only(const StepTail tail) { step_tail(tail); }

cudaError_t launch_adam(float* p, float* g, float* m, float* v, float* gabs, int64_t n, const AdamScalars& s_in,
                        bool zero_grad, const StepTail& tail, cudaStream_t st, const ReplayArgs& rp) {
    if (n == 0) return cudaSuccess;
    AdamScalars s = s_in;
    s.keep_p = l2_keep_ok(n / 4);            // parameters tagged evict_last when grid + gradient fit the L2 (measured: +1.3 %)
    s.keep_g = false;                        // the same tag on the gradient lost 3 us on C2
    s.skip_same = tuning().adam_skip_same < 0 ? !l2_keep_ok(n / 4) : tuning().adam_skip_same != 0;
    const bool aligned = ((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                         ((uintptr_t)v % 16 == 0) && (!gabs || (uintptr_t)gabs % 16 == 0);
    const int64_t n4 = aligned ? n / 4 : 0;
    const int threads = 256;
    if (n4 > 0) {
        cudaError_t e;
        float4 *p4 = (float4*)p, *g4 = (float4*)g, *m4 = (float4*)m, *v4 = (float4*)v, *a4 = (float4*)gabs;
        if (gabs) e = zero_grad ? launch_adam_vec<true, true>(p4, g4, m4, v4, a4, n4, s, tail, st, rp) : launch_adam_vec<true, false>(p4, g4, m4, v4, a4, n4, s, tail, st, rp);
        else      e = zero_grad ? launch_adam_vec<false, true>(p4, g4, m4, v4, nullptr, n4, s, tail, st, rp) : launch_adam_vec<false, false>(p4, g4, m4, v4, nullptr, n4, s, tail, st, rp);
        if (e != cudaSuccess) return e;
    }
    if (n4 == 0 && (tail.src || tail.clear)) k_step_tail_only<<<1, 32, 0, st>>>(tail);
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
