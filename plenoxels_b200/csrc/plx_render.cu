// plx_render.cu — K1 (fused forward) and K2 (fused backward) of the voxel-grid volume renderer, sm_100a.
//
// One warp marches one ray: lane l of chunk c handles sample k = k0 + 32 c + l of the clipped range.
// Replaces the ATen sequence of scripts/train.py:130-151 (+ its autograd at :181) without materialising any
// (M,.) temporary: sample placement (src/ray_sampling.py:161-167), normalisation (:13), lookup through
// clip(0,1) with the in-bounds mask (src/grid_functions.py:103-114 / :7-44,:220-246), compositing
// (src/ray_sampling.py:181-191).
//
// Template axes:  MODE  nearest / trilinear lookup
//                 FAST  contiguous 16-byte-aligned grid + divisor in the hoisted-reciprocal range: cells are addressed
//                       by their linear index with one 128-bit load, the quotient costs 3 FMAs (plx_device.cuh);
//                       !FAST = arbitrary strides (channel-planar pooled grids), scalar loads, __fdiv_rn
//                 DBG   (forward only) full march for the per-ray count / per-sample index dump
#include <cmath>
#include <cstdlib>

#include "plx_march.cuh"
#include "plx_launch.h"

namespace plx {

// image epilogue of visulize_3d_in_2d (src/visualization.py:150-154): (pix * 255).round().clip(0, 255).astype(uint8), stored
// transposed (ray (iu, iv) of the u-major lattice -> row iv, column iu)
__device__ __forceinline__ void store_pixel_u8(const PlxRenderFwd& a, int64_t ray, float r, float g, float b, float al) {
    const int side = a.image_side;
    const int iu = (int)(ray / side), iv = (int)(ray % side);
    auto q = [](float v) { return (unsigned)fminf(fmaxf(rintf(__fmul_rn(v, 255.f)), 0.f), 255.f); };
    const unsigned px = q(r) | (q(g) << 8) | (q(b) << 16) | (q(al) << 24);
    reinterpret_cast<unsigned*>(a.image_u8)[(int64_t)iv * side + iu] = px;
}

// =================================================================================================================
// K1 — forward
// =================================================================================================================
template <int MODE, bool FAST, bool DBG>
__global__ void __launch_bounds__(MAX_WARPS_PER_BLOCK * 32) k_render_fwd(const PlxRenderFwd a) {
    __shared__ float s_loss[MAX_WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t ray = (int64_t)blockIdx.x * wpb + wib;
    const PlxMarch& m = a.march;
    float ray_loss = 0.f;
    if (ray < a.rays.n_rays) {
        const Geo g = make_geo(m);
        const Ray r = load_ray(a.rays, ray);
        const bool fast_ray = FAST && ray_in_fast_range(m, r);
        const bool dump = DBG && a.sample_index != nullptr;
        const bool full = DBG;                 // count / dump / PLX_NO_EARLY_STOP: march to the end of the range
        int k0, k1;
        clip_range(m, r, k0, k1);
        if (dump) { k0 = 1; k1 = m.num_samples; }
        const int nch_all = num_chunks(m.num_samples);
        float* tc = a.tcarry ? a.tcarry + ray * nch_all : nullptr;
        float T = 1.f;                       // transmittance in front of the current chunk
        float ar = 0.f, ag = 0.f, ab = 0.f, aa = 0.f, ad = 0.f;
        int cnt = 0;
        int c = 0;
        bool alive = true;                   // false once T == 0 exactly: later samples have weight 0
        if (MODE == PLX_NEAREST && !DBG) {
            // nearest mode, production path: software-pipelined — the cell of chunk c+1 is requested before chunk c is
            // composited, so the gather latency overlaps the scan (same scheme as K12, plx_train.cu)
            float4 rawn;
            int linn = k0 <= k1 ? fetch_nearest<FAST>(m, g, a.grid, r, fast_ray, k0 + lane, k0 + lane <= k1, rawn) : -1;
            for (int kb = k0; kb <= k1; kb += CHUNK, ++c) {
                const float4 cell = g.clamp ? clamp4(rawn) : rawn;
                const float t = __fmul_rn(g.delta, (float)(kb + lane));
                (void)linn;
                if (kb + CHUNK <= k1) linn = fetch_nearest<FAST>(m, g, a.grid, r, fast_ray, kb + CHUNK + lane, kb + CHUNK + lane <= k1, rawn);
                if (tc && lane == 0) tc[c] = T;
                float total;
                const float ex = warp_excl_prod(1.f - cell.w, lane, total);
                const float w = cell.w * (T * ex);          // alpha_k * T_k, src/ray_sampling.py:184
                ar = fmaf(w, cell.x, ar); ag = fmaf(w, cell.y, ag); ab = fmaf(w, cell.z, ab);
                aa += w;
                ad = fmaf(w, t, ad);
                T *= total;
                if (T == 0.f) { alive = false; ++c; break; }
            }
        } else
        for (int kb = k0; kb <= k1; kb += CHUNK, ++c) {
            const int k = kb + lane;
            const bool valid = k <= k1;
            if (tc && lane == 0) tc[c] = T;
            float t;
            TriGeom tg;
            const Sample s = lookup<MODE, FAST>(m, g, a.grid, r, fast_ray, k, valid, alive, t, tg);
            if (DBG) {
                cnt += s.inb ? 1 : 0;
                if (dump && valid) a.sample_index[ray * (int64_t)m.num_samples + (k - 1)] = s.lin;
            }
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
        if (tc) {                            // chunks never reached carry zero transmittance (or are unused)
            for (int cc = c + lane; cc < nch_all; cc += 32) tc[cc] = alive ? T : 0.f;
        }
        ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); aa = warp_sum(aa);
        if (a.depth) ad = warp_sum(ad);
        if (DBG && a.count) cnt = warp_sum_int(cnt);
        if (lane == 0) {
            if (a.rgba) reinterpret_cast<float4*>(a.rgba)[ray] = make_float4(ar, ag, ab, aa);
            if (a.image_u8) store_pixel_u8(a, ray, ar, ag, ab, aa);
            if (a.depth) a.depth[ray] = ad;
            if (DBG && a.count) a.count[ray] = cnt;
            if (a.targets) {                 // mean-MSE over N*4 incl. alpha, scripts/train.py:156
                const float4 tgt = __ldg(reinterpret_cast<const float4*>(a.targets) + ray);
                const float dr = ar - tgt.x, dg = ag - tgt.y, db = ab - tgt.z, da = aa - tgt.w;
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
            for (int i = 0; i < wpb; ++i) sum += s_loss[i];
            atomicAdd(a.loss, sum);
        }
    }
}

// =================================================================================================================
// K1p — forward for coherent rays (inference): one ray per thread, 32 neighbouring rays per warp
// =================================================================================================================
// Image-order rays of one view pass through neighbouring cells at equal depth, so the 32 lanes of a load share sectors
// and L1 lines; the transmittance recurrence runs sequentially in registers (no warp scan), UNR samples are looked up
// (independent loads) before they are composited.  Same exact index arithmetic, same results as K1 up to summation order.
// `side` > 0: the rays of every origin are a side x side lattice (ray = view * side^2 + iu * side + iv) and a block marches a
// 16 x 8 TILE of it (2 x 2 warps of 8 x 4 rays) instead of 128 consecutive rays of one lattice row: the cells a warp touches at
// equal depth then form a compact patch (fewer sectors per load, L1 lines shared by the warps of the block and by consecutive
// samples).  Which thread marches a ray does not change the ray's result.  Measured on C4 (512^3, 800 x 800 view, B200):
// rows 0.77 ms -> tiles 0.62 ms per frame at 64 resident warps / SM (0.61 with one lookup per loop trip); 4 x 8 and 16 x 2 warp
// tiles were within 2 % of 8 x 4.
template <int MODE, bool FAST, int UNR, int MINB = 1>
__global__ void __launch_bounds__(128, MINB) k_render_fwd_packet(const PlxRenderFwd a, const int side, const int tiles_v, const int tiles_per_view) {
    int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (side > 0) {
        const int view = blockIdx.x / tiles_per_view, tile = blockIdx.x - view * tiles_per_view;
        const int tu = tile / tiles_v, tv = tile - tu * tiles_v;
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        const int iv = (tv * 2 + (w & 1)) * 8 + (l & 7), iu = (tu * 2 + (w >> 1)) * 4 + (l >> 3);
        if (iv >= side || iu >= side) return;
        ray = ((int64_t)view * side + iu) * side + iv;
    }
    if (ray >= a.rays.n_rays) return;
    const PlxMarch& m = a.march;
    const Geo g = make_geo(m);
    const Ray r = load_ray(a.rays, ray);
    const bool fast_ray = FAST && ray_in_fast_range(m, r);
    int k0, k1;
    clip_range(m, r, k0, k1);
    float T = 1.f, ar = 0.f, ag = 0.f, ab = 0.f, aa = 0.f, ad = 0.f;
    for (int kb = k0; kb <= k1 && T != 0.f; kb += UNR) {
        float4 c[UNR];
        float t[UNR];
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            TriGeom tg;
            c[j] = lookup<MODE, FAST>(m, g, a.grid, r, fast_ray, kb + j, kb + j <= k1, true, t[j], tg).c;
        }
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const float w = c[j].w * T;                   // alpha_k * T_k, src/ray_sampling.py:184
            ar = fmaf(w, c[j].x, ar); ag = fmaf(w, c[j].y, ag); ab = fmaf(w, c[j].z, ab);
            aa += w;
            ad = fmaf(w, t[j], ad);
            T *= 1.f - c[j].w;
        }
    }
    if (a.rgba) reinterpret_cast<float4*>(a.rgba)[ray] = make_float4(ar, ag, ab, aa);
    if (a.image_u8) store_pixel_u8(a, ray, ar, ag, ab, aa);
    if (a.depth) a.depth[ray] = ad;
}

// =================================================================================================================
// K2 — backward
// =================================================================================================================
template <int MODE, bool FAST>
__global__ void __launch_bounds__(MAX_WARPS_PER_BLOCK * 32) k_render_bwd(const PlxRenderBwd a) {
    extern __shared__ float s_tc[];          // [warps per block][nch_all], only used when a.tcarry == NULL
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t ray = (int64_t)blockIdx.x * wpb + wib;
    const PlxMarch& m = a.march;
    if (ray >= a.rays.n_rays) return;
    const float4 gr = __ldg(reinterpret_cast<const float4*>(a.grad_rgba) + ray);
    const float bom = a.beta_over_m;
    if (gr.x == 0.f && gr.y == 0.f && gr.z == 0.f && gr.w == 0.f && bom == 0.f) return;
    const Geo g = make_geo(m);
    const Ray r = load_ray(a.rays, ray);
    const bool fast_ray = FAST && ray_in_fast_range(m, r);
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
                const Sample s = lookup<MODE, FAST>(m, g, a.grid, r, fast_ray, k, k <= k1, true, t, tg);
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
    // nearest mode: the cell of chunk c-1 is requested before chunk c is processed (software pipeline, as in K1 / K12)
    float4 rawn = make_float4(0.f, 0.f, 0.f, 0.f);
    int linn = -1;
    if (MODE == PLX_NEAREST) {
        const int kl = k0 + (nch - 1) * CHUNK + lane;
        linn = fetch_nearest<FAST>(m, g, a.grid, r, fast_ray, kl, kl <= k1, rawn);
    }
    for (int c = nch - 1; c >= 0; --c) {
        const int k = k0 + c * CHUNK + lane;
        const bool valid = k <= k1;
        const float Tc = tc[c];
        float t;
        TriGeom tg;
        Sample s;
        if (MODE == PLX_NEAREST) {
            s.raw = rawn;
            s.c = g.clamp ? clamp4(rawn) : rawn;
            s.lin = linn;
            s.inb = linn >= 0;
            if (c > 0) {
                const int kp = k - CHUNK;
                linn = fetch_nearest<FAST>(m, g, a.grid, r, fast_ray, kp, kp <= k1, rawn);
            }
        } else {
            s = lookup<MODE, FAST>(m, g, a.grid, r, fast_ray, k, valid, true, t, tg);
        }
        const float alpha = s.c.w;
        const float v = fmaf(s.c.x, gr.x, fmaf(s.c.y, gr.y, fmaf(s.c.z, gr.z, gr.w)));     // c_k . g_rgb + g_A
        const float behind = warp_behind(alpha * v, 1.f - alpha, lane, behind_carry);
        if (Tc == 0.f && bom == 0.f) continue;               // every T_k of this chunk is 0: no gradient here
        float total;
        const float Tk = Tc * warp_excl_prod(1.f - alpha, lane, total);
        const float wgt = alpha * Tk;
        float dr = wgt * gr.x, dg = wgt * gr.y, db = wgt * gr.z;
        float da = Tk * (v - behind);
        if (bom != 0.f) da += bom * (1.f / (alpha + 1e-4f) + 1.f / (1.f - alpha + 1e-4f));   // scripts/train.py:170-177
        if (MODE == PLX_NEAREST) {
            if (g.clamp) { dr *= pass01(s.raw.x); dg *= pass01(s.raw.y); db *= pass01(s.raw.z); da *= pass01(s.raw.w); }
            warp_scatter_add(a.grad_grid, s.inb, (int64_t)s.lin * 4, dr, dg, db, da, lane, g.pol_grad);
        } else {
            if (s.inb && (dr != 0.f || dg != 0.f || db != 0.f || da != 0.f)) {
#pragma unroll
                for (int corner = 0; corner < 8; ++corner) {
                    const bool cx = corner & 4, cy = corner & 2, cz = corner & 1;   // 1 = floor side
                    const int ix = cx ? tg.lo[0] : tg.hi[0], iy = cy ? tg.lo[1] : tg.hi[1], iz = cz ? tg.lo[2] : tg.hi[2];
                    const float w = (cx ? 1.f - tg.f[0] : tg.f[0]) * (cy ? 1.f - tg.f[1] : tg.f[1]) *
                                    (cz ? 1.f - tg.f[2] : tg.f[2]);
                    if (w == 0.f) continue;
                    const int lin = (ix * g.ny + iy) * g.nz + iz;
                    float px = 1.f, py = 1.f, pz = 1.f, pw = 1.f;
                    if (g.clamp) {
                        const float4 raw = cell_at<FAST>(m, g, a.grid, ix, iy, iz, lin);
                        px = pass01(raw.x); py = pass01(raw.y); pz = pass01(raw.z); pw = pass01(raw.w);
                    }
                    red_add_v4_hint(a.grad_grid + (int64_t)lin * 4, dr * w * px, dg * w * py, db * w * pz, da * w * pw, g.pol_grad);
                }
            }
        }
    }
}

// =================================================================================================================
// launchers
// =================================================================================================================
static int warps_per_block(const char* env, int dflt) {
    const char* e = std::getenv(env);
    if (e) {
        const int v = std::atoi(e);
        if (v >= 1 && v <= MAX_WARPS_PER_BLOCK) return v;
    }
    return dflt;
}

cudaError_t launch_render_fwd(const PlxRenderFwd& a_in, cudaStream_t st) {
    if (a_in.rays.n_rays == 0) return cudaSuccess;
    PlxRenderFwd a = a_in;
    a.march.flags = (a.march.flags & 0xffffu) | l2_keep_flags(a.march);
    static const int wpb = warps_per_block("PLX_FWD_WPB", 4);
    const unsigned blocks = (unsigned)((a.rays.n_rays + wpb - 1) / wpb);
    const bool fast = fast_ok(a.march, a.grid);
    const bool dbg = a.count || a.sample_index || (a.march.flags & PLX_NO_EARLY_STOP);
    if ((a.march.flags & PLX_COHERENT_RAYS) && !dbg && !a.tcarry && !a.targets) {
        // the coherent-ray hint promises the even-spread lattice: when every origin carries a square number of rays, march it in
        // 16 x 8 tiles (tuning().packet_tile = 0 keeps the lattice-row order: tests compare the two bit for bit)
        unsigned pblocks = (unsigned)((a.rays.n_rays + 127) / 128);
        int side = 0, tiles_v = 0, tiles_per_view = 0;
        const int64_t rpo = a.rays.rays_per_origin;
        if (tuning().packet_tile && rpo >= 64 && a.rays.n_rays % rpo == 0) {
            const int64_t sd = (int64_t)std::llround(std::sqrt((double)rpo));
            if (sd * sd == rpo && sd < (1 << 20)) {
                side = (int)sd;
                tiles_v = (side + 15) / 16;
                tiles_per_view = tiles_v * ((side + 7) / 8);
                const int64_t nb = (a.rays.n_rays / rpo) * tiles_per_view;
                if (nb < (1ll << 31)) pblocks = (unsigned)nb; else side = 0;
            }
        }
        // The kernel is gather-latency bound on a grid far beyond the L2 (C4: long_scoreboard 12 stalled warps per issue), so
        // resident warps beat samples in flight per thread: nearest = 1 lookup per loop trip at 32 registers (64 warps / SM),
        // trilinear (8 corners per sample) = 1 at 48 registers.  Sweep on C4, ms per frame: nearest (4 lookups, 54 registers) 0.94,
        // (4, 40 r) 0.70, (3, 32 r) 0.66, (2, 40 r) 0.67, (2, 32 r) 0.62, (1, 32 r) 0.61; trilinear (1, 68 r) 2.18, (1, 48 r) 2.06,
        // (1, 40 r) 2.38, (2, 64 r) 2.12.
#define PLX_PACKET(...) k_render_fwd_packet<__VA_ARGS__><<<pblocks, 128, 0, st>>>(a, side, tiles_v, tiles_per_view)
        if (a.march.mode == PLX_NEAREST) { if (fast) PLX_PACKET(PLX_NEAREST, true, 1, 16); else PLX_PACKET(PLX_NEAREST, false, 4); }
        else                             { if (fast) PLX_PACKET(PLX_TRILINEAR, true, 1, 10); else PLX_PACKET(PLX_TRILINEAR, false, 1); }
#undef PLX_PACKET
        return cudaGetLastError();
    }
#define PLX_FWD(MODE)                                                                                     \
    do {                                                                                                  \
        if (fast) { if (dbg) k_render_fwd<MODE, true, true><<<blocks, wpb * 32, 0, st>>>(a);              \
                    else     k_render_fwd<MODE, true, false><<<blocks, wpb * 32, 0, st>>>(a); }           \
        else      { if (dbg) k_render_fwd<MODE, false, true><<<blocks, wpb * 32, 0, st>>>(a);             \
                    else     k_render_fwd<MODE, false, false><<<blocks, wpb * 32, 0, st>>>(a); }          \
    } while (0)
    if (a.march.mode == PLX_NEAREST) PLX_FWD(PLX_NEAREST); else PLX_FWD(PLX_TRILINEAR);
#undef PLX_FWD
    return cudaGetLastError();
}

cudaError_t launch_render_bwd(const PlxRenderBwd& a_in, cudaStream_t st) {
    if (a_in.rays.n_rays == 0) return cudaSuccess;
    PlxRenderBwd a = a_in;
    a.march.flags = (a.march.flags & 0xffffu) | l2_keep_flags(a.march);
    static const int wpb = warps_per_block("PLX_BWD_WPB", 4);
    const unsigned blocks = (unsigned)((a.rays.n_rays + wpb - 1) / wpb);
    const size_t smem = a.tcarry ? 0 : (size_t)wpb * num_chunks(a.march.num_samples) * sizeof(float);
    const bool fast = fast_ok(a.march, a.grid);
#define PLX_BWD(MODE, FASTP)                                                                                            \
    do {                                                                                                                \
        if (smem > 48 * 1024) {                                                                                         \
            cudaError_t e = cudaFuncSetAttribute(k_render_bwd<MODE, FASTP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 (int)smem);                                                            \
            if (e != cudaSuccess) return e;                                                                             \
        }                                                                                                               \
        k_render_bwd<MODE, FASTP><<<blocks, wpb * 32, smem, st>>>(a);                                                  \
    } while (0)
    if (a.march.mode == PLX_NEAREST) {
        if (fast) PLX_BWD(PLX_NEAREST, true); else PLX_BWD(PLX_NEAREST, false);
    } else {
        if (fast) PLX_BWD(PLX_TRILINEAR, true); else PLX_BWD(PLX_TRILINEAR, false);
    }
#undef PLX_BWD
    return cudaGetLastError();
}

}  // namespace plx
