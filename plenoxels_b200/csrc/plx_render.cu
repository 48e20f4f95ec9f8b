// plx_render.cu — K1 (fused forward) and K2 (fused backward) of the voxel-grid volume renderer, sm_100a.
//
// One warp marches one ray: lane l of chunk c handles sample k = k0 + 32 c + l of the clipped range.
// Replaces the ATen sequence of scripts/train.py:130-151 (+ its autograd at :181) without materialising any
// (M,.) temporary: sample placement (src/ray_sampling.py:161-167), normalisation (:13), lookup through
// clip(0,1) with the in-bounds mask (src/grid_functions.py:103-114 / :7-44,:220-246), compositing
// (src/ray_sampling.py:181-191).
#include "plx_device.cuh"
#include "plx_launch.h"

namespace plx {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int THREADS = WARPS_PER_BLOCK * 32;

// One sample's lookup result.
struct Sample {
    float4 c;        // (clamped) cell value, 0 when out of bounds
    bool inb;        // the reference's mask (True = inside)
    int64_t off;     // element offset of the (nearest / floor-corner) cell in the *strided* grid
    int32_t lin;     // contiguous linear cell index (ix*ny+iy)*nz+iz of that cell, -1 when out of bounds
};

// ---- nearest neighbour: src/grid_functions.py:111 (round half even), :58-61 (mask) -------------------------
template <bool VEC>
__device__ __forceinline__ Sample lookup_nearest(const PlxMarch& m, const float* __restrict__ grid, float nx, float ny,
                                                 float nz, bool valid, bool need_value) {
    Sample s;
    s.c = make_float4(0.f, 0.f, 0.f, 0.f);
    const float rx = rintf(nx), ry = rintf(ny), rz = rintf(nz);
    s.inb = valid && rx >= 0.f && rx < (float)m.nx && ry >= 0.f && ry < (float)m.ny && rz >= 0.f && rz < (float)m.nz;
    s.off = 0;
    s.lin = -1;
    if (s.inb) {
        const int ix = (int)rx, iy = (int)ry, iz = (int)rz;
        s.lin = (ix * m.ny + iy) * m.nz + iz;
        s.off = ix * m.sx + iy * m.sy + iz * m.sz;
        if (need_value) {
            float4 c = load_cell<VEC>(grid, s.off, m.sc);
            if (m.flags & PLX_CLAMP01) { c.x = clamp01(c.x); c.y = clamp01(c.y); c.z = clamp01(c.z); c.w = clamp01(c.w); }
            s.c = c;
        }
    }
    return s;
}

// ---- trilinear: SURVEY.md §8a row T ---------------------------------------------------------------------------
struct TriGeom {
    int lo[3], hi[3];     // floor / wrapped ceil index per axis
    float f[3];           // frac per axis
};

__device__ __forceinline__ bool tri_geom(const PlxMarch& m, float nx, float ny, float nz, TriGeom& g) {
    const float n[3] = {nx, ny, nz};
    const int dim[3] = {m.nx, m.ny, m.nz};
    bool inb = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) inb = inb && (n[a] >= 0.f) && (n[a] < (float)dim[a]);   // float test, :58-61
    if (!inb) return false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float fl = floorf(n[a]);
        g.lo[a] = (int)fl;
        int hi = (int)ceilf(n[a]);
        g.hi[a] = hi >= dim[a] ? hi - dim[a] : hi;            // periodic wrap of the ceil corner, :75-77
        g.f[a] = __fsub_rn(n[a], fl);                          // torch.frac for n >= 0, :29
    }
    return true;
}

// hi*f + lo*(1-f), each op rounded — src/grid_functions.py:35,:39,:42
__device__ __forceinline__ float lerp_ref(float hi, float lo, float f) {
    return __fadd_rn(__fmul_rn(hi, f), __fmul_rn(lo, __fsub_rn(1.f, f)));
}
__device__ __forceinline__ float4 lerp4(float4 hi, float4 lo, float f) {
    return make_float4(lerp_ref(hi.x, lo.x, f), lerp_ref(hi.y, lo.y, f), lerp_ref(hi.z, lo.z, f), lerp_ref(hi.w, lo.w, f));
}

template <bool VEC>
__device__ __forceinline__ float4 tri_cell(const PlxMarch& m, const float* __restrict__ grid, int ix, int iy, int iz) {
    float4 c = load_cell<VEC>(grid, ix * m.sx + iy * m.sy + iz * m.sz, m.sc);
    if (m.flags & PLX_CLAMP01) { c.x = clamp01(c.x); c.y = clamp01(c.y); c.z = clamp01(c.z); c.w = clamp01(c.w); }
    return c;
}

template <bool VEC>
__device__ __forceinline__ float4 tri_interp(const PlxMarch& m, const float* __restrict__ grid, const TriGeom& g) {
    // x-lerp of the four (y,z) edges, then y, then z — corner order of src/grid_functions.py:238-243
    float4 x_cc = lerp4(tri_cell<VEC>(m, grid, g.hi[0], g.hi[1], g.hi[2]), tri_cell<VEC>(m, grid, g.lo[0], g.hi[1], g.hi[2]), g.f[0]);
    float4 x_cf = lerp4(tri_cell<VEC>(m, grid, g.hi[0], g.hi[1], g.lo[2]), tri_cell<VEC>(m, grid, g.lo[0], g.hi[1], g.lo[2]), g.f[0]);
    float4 x_fc = lerp4(tri_cell<VEC>(m, grid, g.hi[0], g.lo[1], g.hi[2]), tri_cell<VEC>(m, grid, g.lo[0], g.lo[1], g.hi[2]), g.f[0]);
    float4 x_ff = lerp4(tri_cell<VEC>(m, grid, g.hi[0], g.lo[1], g.lo[2]), tri_cell<VEC>(m, grid, g.lo[0], g.lo[1], g.lo[2]), g.f[0]);
    float4 y_c = lerp4(x_cc, x_fc, g.f[1]);
    float4 y_f = lerp4(x_cf, x_ff, g.f[1]);
    return lerp4(y_c, y_f, g.f[2]);
}

template <bool VEC>
__device__ __forceinline__ Sample lookup_trilinear(const PlxMarch& m, const float* __restrict__ grid, float nx, float ny,
                                                   float nz, bool valid, bool need_value, TriGeom& g) {
    Sample s;
    s.c = make_float4(0.f, 0.f, 0.f, 0.f);
    s.off = 0;
    s.lin = -1;
    s.inb = valid && tri_geom(m, nx, ny, nz, g);
    if (s.inb) {
        s.lin = (g.lo[0] * m.ny + g.lo[1]) * m.nz + g.lo[2];
        if (need_value) s.c = tri_interp<VEC>(m, grid, g);
    }
    return s;
}

template <int MODE, bool VEC>
__device__ __forceinline__ Sample lookup(const PlxMarch& m, const float* __restrict__ grid, const Ray& r, int k, bool valid,
                                         bool need_value, float& t, TriGeom& g) {
    t = step_t(m.delta_step, k);
    const float nx = norm_coord(r.ox, r.dx, t, m.gmin[0], m.points_distance);
    const float ny = norm_coord(r.oy, r.dy, t, m.gmin[1], m.points_distance);
    const float nz = norm_coord(r.oz, r.dz, t, m.gmin[2], m.points_distance);
    if (MODE == PLX_NEAREST) return lookup_nearest<VEC>(m, grid, nx, ny, nz, valid, need_value);
    return lookup_trilinear<VEC>(m, grid, nx, ny, nz, valid, need_value, g);
}

// =================================================================================================================
// K1 — forward
// =================================================================================================================
template <int MODE, bool VEC>
__global__ void __launch_bounds__(THREADS) k_render_fwd(const PlxRenderFwd a) {
    __shared__ float s_loss[WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t ray = (int64_t)blockIdx.x * WARPS_PER_BLOCK + wib;
    const PlxMarch& m = a.march;
    float ray_loss = 0.f;
    if (ray < a.rays.n_rays) {
        const Ray r = load_ray(a.rays, ray);
        const bool dump = a.sample_index != nullptr;
        const bool full = dump || a.count != nullptr || (m.flags & PLX_NO_EARLY_STOP);
        int k0, k1;
        clip_range(m, r, k0, k1);
        if (dump) { k0 = 1; k1 = m.num_samples; }
        const int nch_all = num_chunks(m.num_samples);
        float T = 1.f;                       // transmittance in front of the current chunk
        float ar = 0.f, ag = 0.f, ab = 0.f, aa = 0.f, ad = 0.f;
        int cnt = 0;
        int c = 0;
        bool alive = true;                   // false once T == 0 exactly: later samples have weight 0
        for (int kb = k0; kb <= k1; kb += CHUNK, ++c) {
            const int k = kb + lane;
            const bool valid = k <= k1;
            if (a.tcarry && lane == 0) a.tcarry[ray * nch_all + c] = T;
            float t;
            TriGeom g;
            const Sample s = lookup<MODE, VEC>(m, a.grid, r, k, valid, alive, t, g);
            cnt += s.inb ? 1 : 0;
            if (dump && valid) a.sample_index[ray * (int64_t)m.num_samples + (k - 1)] = s.lin;
            if (alive) {
                float total;
                const float ex = warp_excl_prod(1.f - s.c.w, lane, total);
                const float w = s.c.w * (T * ex);           // alpha_k * T_k, src/ray_sampling.py:184
                ar = fmaf(w, s.c.x, ar); ag = fmaf(w, s.c.y, ag); ab = fmaf(w, s.c.z, ab);
                aa += w;
                ad = fmaf(w, t, ad);
                T *= total;
                if (T == 0.f) {
                    alive = false;
                    if (!full) { ++c; break; }
                }
            }
        }
        if (a.tcarry) {                      // chunks never reached carry zero transmittance (or are unused)
            for (int cc = c + lane; cc < nch_all; cc += 32) a.tcarry[ray * nch_all + cc] = alive ? T : 0.f;
        }
        ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); aa = warp_sum(aa);
        if (a.depth) ad = warp_sum(ad);
        if (a.count) cnt = warp_sum_int(cnt);
        if (lane == 0) {
            reinterpret_cast<float4*>(a.rgba)[ray] = make_float4(ar, ag, ab, aa);
            if (a.depth) a.depth[ray] = ad;
            if (a.count) a.count[ray] = cnt;
            if (a.targets) {                 // mean-MSE over N*4 incl. alpha, scripts/train.py:156
                const float4 tg = __ldg(reinterpret_cast<const float4*>(a.targets) + ray);
                const float dr = ar - tg.x, dg = ag - tg.y, db = ab - tg.z, da = aa - tg.w;
                reinterpret_cast<float4*>(a.grad_rgba)[ray] =
                    make_float4(dr * a.grad_scale, dg * a.grad_scale, db * a.grad_scale, da * a.grad_scale);
                ray_loss = (dr * dr + dg * dg + db * db + da * da) * a.loss_scale;
            }
        }
    }
    if (a.targets && a.loss) {               // one atomic per block
        if (lane == 0) s_loss[wib] = ray_loss;
        __syncthreads();
        if (threadIdx.x == 0) {
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < WARPS_PER_BLOCK; ++i) sum += s_loss[i];
            atomicAdd(a.loss, sum);
        }
    }
}

// =================================================================================================================
// K2 — backward
// =================================================================================================================
template <int MODE, bool VEC>
__global__ void __launch_bounds__(THREADS) k_render_bwd(const PlxRenderBwd a) {
    extern __shared__ float s_tc[];          // [WARPS_PER_BLOCK][nch_all], only used when a.tcarry == NULL
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t ray = (int64_t)blockIdx.x * WARPS_PER_BLOCK + wib;
    const PlxMarch& m = a.march;
    if (ray >= a.rays.n_rays) return;
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.grad_rgba) + ray);
    const float bom = a.beta_over_m;
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f && bom == 0.f) return;
    const Ray r = load_ray(a.rays, ray);
    int k0, k1;
    clip_range(m, r, k0, k1);
    if (k0 > k1) return;
    const int nch_all = num_chunks(m.num_samples);
    const int nch = (k1 - k0) / CHUNK + 1;

    // ---- pass 1 (only without a saved tcarry): transmittance in front of every chunk
    const float* tc;
    if (a.tcarry) {
        tc = a.tcarry + ray * nch_all;
    } else {
        float* mine = s_tc + wib * nch_all;
        float T = 1.f;
        for (int c = 0; c < nch; ++c) {
            if (lane == 0) mine[c] = T;
            if (T != 0.f) {
                const int k = k0 + c * CHUNK + lane;
                float t;
                TriGeom tg;
                const Sample s = lookup<MODE, VEC>(m, a.grid, r, k, k <= k1, true, t, tg);
                float f = 1.f - s.c.w;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) f *= __shfl_xor_sync(FULL, f, d);
                T *= f;
            }
        }
        __syncwarp();
        tc = mine;
    }

    // ---- pass 2: reverse over the chunks
    float behind_carry = 0.f;                // S behind the last sample of the ray = 0
    for (int c = nch - 1; c >= 0; --c) {
        const int k = k0 + c * CHUNK + lane;
        const bool valid = k <= k1;
        const float Tc = tc[c];
        float t;
        TriGeom tg;
        const Sample s = lookup<MODE, VEC>(m, a.grid, r, k, valid, true, t, tg);
        const float alpha = s.c.w;
        const float v = fmaf(s.c.x, g.x, fmaf(s.c.y, g.y, fmaf(s.c.z, g.z, g.w)));     // c_k . g_rgb + g_A
        const float behind = warp_behind(alpha * v, 1.f - alpha, lane, behind_carry);
        if (Tc == 0.f && bom == 0.f) continue;               // every T_k of this chunk is 0: no gradient here
        float total;
        const float Tk = Tc * warp_excl_prod(1.f - alpha, lane, total);
        const float wgt = alpha * Tk;
        float dr = wgt * g.x, dg = wgt * g.y, db = wgt * g.z;
        float da = Tk * (v - behind);
        if (bom != 0.f) da += bom * (1.f / (alpha + 1e-4f) + 1.f / (1.f - alpha + 1e-4f));   // scripts/train.py:170-177
        if (MODE == PLX_NEAREST) {
            if (s.inb && (m.flags & PLX_CLAMP01)) {
                const float4 raw = load_cell<VEC>(a.grid, s.off, m.sc);
                dr *= pass01(raw.x); dg *= pass01(raw.y); db *= pass01(raw.z); da *= pass01(raw.w);
            }
            warp_scatter_add(a.grad_grid, s.inb, (int64_t)s.lin * 4, dr, dg, db, da, lane);
        } else {
            if (s.inb && (dr != 0.f || dg != 0.f || db != 0.f || da != 0.f)) {
#pragma unroll
                for (int corner = 0; corner < 8; ++corner) {
                    const bool cx = corner & 4, cy = corner & 2, cz = corner & 1;   // 1 = floor side
                    const int ix = cx ? tg.lo[0] : tg.hi[0], iy = cy ? tg.lo[1] : tg.hi[1], iz = cz ? tg.lo[2] : tg.hi[2];
                    const float w = (cx ? 1.f - tg.f[0] : tg.f[0]) * (cy ? 1.f - tg.f[1] : tg.f[1]) *
                                    (cz ? 1.f - tg.f[2] : tg.f[2]);
                    if (w == 0.f) continue;
                    float px = 1.f, py = 1.f, pz = 1.f, pw = 1.f;
                    if (m.flags & PLX_CLAMP01) {
                        const float4 raw = load_cell<VEC>(a.grid, ix * m.sx + iy * m.sy + iz * m.sz, m.sc);
                        px = pass01(raw.x); py = pass01(raw.y); pz = pass01(raw.z); pw = pass01(raw.w);
                    }
                    const int64_t lin = ((int64_t)ix * m.ny + iy) * m.nz + iz;
                    red_add_v4(a.grad_grid + lin * 4, dr * w * px, dg * w * py, db * w * pz, da * w * pw);
                }
            }
        }
    }
}

// =================================================================================================================
// launchers
// =================================================================================================================
static inline bool vec_ok(const PlxMarch& m, const float* grid) {
    return m.sc == 1 && (m.sx % 4 == 0) && (m.sy % 4 == 0) && (m.sz % 4 == 0) && ((uintptr_t)grid % 16 == 0);
}

cudaError_t launch_render_fwd(const PlxRenderFwd& a, cudaStream_t st) {
    if (a.rays.n_rays == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((a.rays.n_rays + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    const bool vec = vec_ok(a.march, a.grid);
    if (a.march.mode == PLX_NEAREST) {
        if (vec) k_render_fwd<PLX_NEAREST, true><<<blocks, THREADS, 0, st>>>(a);
        else     k_render_fwd<PLX_NEAREST, false><<<blocks, THREADS, 0, st>>>(a);
    } else {
        if (vec) k_render_fwd<PLX_TRILINEAR, true><<<blocks, THREADS, 0, st>>>(a);
        else     k_render_fwd<PLX_TRILINEAR, false><<<blocks, THREADS, 0, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_render_bwd(const PlxRenderBwd& a, cudaStream_t st) {
    if (a.rays.n_rays == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((a.rays.n_rays + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    const size_t smem = a.tcarry ? 0 : (size_t)WARPS_PER_BLOCK * num_chunks(a.march.num_samples) * sizeof(float);
    const bool vec = vec_ok(a.march, a.grid);
#define PLX_BWD(MODE, VEC)                                                                                            \
    do {                                                                                                              \
        if (smem > 48 * 1024) {                                                                                       \
            cudaError_t e = cudaFuncSetAttribute(k_render_bwd<MODE, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 (int)smem);                                                          \
            if (e != cudaSuccess) return e;                                                                           \
        }                                                                                                             \
        k_render_bwd<MODE, VEC><<<blocks, THREADS, smem, st>>>(a);                                                   \
    } while (0)
    if (a.march.mode == PLX_NEAREST) {
        if (vec) PLX_BWD(PLX_NEAREST, true); else PLX_BWD(PLX_NEAREST, false);
    } else {
        if (vec) PLX_BWD(PLX_TRILINEAR, true); else PLX_BWD(PLX_TRILINEAR, false);
    }
#undef PLX_BWD
    return cudaGetLastError();
}

}  // namespace plx
