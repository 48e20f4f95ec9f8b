// plx_train.cu — K12: the whole per-ray training work in ONE kernel (nearest lookup, MSE loss), sm_100a.
//
//   [ray generation] -> forward march + compositing -> MSE gradient of this ray -> reverse march + scatter-add
//
// i.e. scripts/train.py:130-157 and the autograd of :181 for one ray per warp, without the round trips the separate
// K1 / K2 launches need (directions, targets, rgba, grad_rgba, chunk transmittances, a second index computation):
//   * the forward pass caches each sample's linear cell index in shared memory; the reverse pass reads the index back,
//     re-loads the 16-byte cell (L1/L2 hit) and needs no coordinate arithmetic;
//   * each lane owns SPL consecutive samples, so one warp scan serves 32*SPL samples and samples that fall into the
//     same cell (runs along the ray) are merged in registers before the 16-byte vector reduction;
//   * exact early termination in BOTH directions: once the transmittance is exactly 0 (an alpha == 1 sample k*),
//     later samples have zero gradient; the only later quantity the gradient of k* needs is the colour seen behind it,
//     S_k* = sum_{j>k*} alpha_j v_j prod_{k*<i<j} (1-alpha_i), which is linear in the pixel gradient g, so the forward
//     pass keeps compositing a second accumulator behind k* (until that one saturates too) and S_k* = acc2 . g.
//   * trilinear lookup: the forward pass caches each sample's interpolated value (16 B) and the clip pass-mask of its eight
//     corners (4 B) instead of a cell index; the reverse pass recomputes only the geometry (corner indices, lerp weights: no
//     loads) and never repeats the eight gathers, the masks or the seven lerps.
// Same exact index arithmetic as K1 (plx_march.cuh), so indices match the reference bit for bit.
#include <cmath>
#include <cstdlib>

#include "plx_march.cuh"
#include "plx_raygen.cuh"
#include "plx_launch.h"

namespace plx {

// composite one iteration: acc += sum_j alpha_j T_j c_j, T <- T * prod(1 - alpha); returns the product of the iteration
template <int SPL>
__device__ __forceinline__ void composite_iter(const float4 (&c)[SPL], int lane, float& T, float4& acc) {
    float pf[SPL];
    pf[0] = 1.f;
#pragma unroll
    for (int j = 1; j < SPL; ++j) pf[j] = pf[j - 1] * (1.f - c[j - 1].w);
    const float P = pf[SPL - 1] * (1.f - c[SPL - 1].w);
    float total;
    const float base = T * warp_excl_prod(P, lane, total);
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const float w = c[j].w * (base * pf[j]);          // alpha_k * T_k, src/ray_sampling.py:184
        acc.x = fmaf(w, c[j].x, acc.x); acc.y = fmaf(w, c[j].y, acc.y); acc.z = fmaf(w, c[j].z, acc.z);
        acc.w += w;
    }
    T *= total;
}

// where the gradient of cell `lin` goes: the local buffer, or (multi-GPU push exchange, PlxPeerGrad) the buffer of the rank
// that owns the cell's slab — a peer-mapped pointer, so the reduction travels over NVLink and lands in the owner's L2
template <bool PEER>
struct GradDst {
    float* local;
    float* const* peers;     // shared-memory copy of PlxPeerGrad.grads
    uint32_t owner_mul;
    uint64_t pol;
    __device__ __forceinline__ void add(int lin, float x, float y, float z, float w) const {
        if (PEER) red_add_v4(peers[__umulhi((uint32_t)lin, owner_mul)] + (int64_t)lin * 4, x, y, z, w);
        else      red_add_v4_hint(local + (int64_t)lin * 4, x, y, z, w, pol);
    }
};

// per-warp scratch of one ray
struct RayScratch {
    int* lc;        // nearest: cached linear cell index of every visited sample
    float* tcs;     // transmittance in front of every iteration
    float4* cv;     // trilinear: cached interpolated value of every visited sample ...
    uint32_t* mk;   // ... and the clip pass-mask bits of its eight corners
};

// One ray, start to finish: forward march + compositing, MSE gradient, reverse march + scatter-add.  Returns the ray's loss.
// FR (fast ray): every coordinate of the ray is in the hoisted-division range — the common case, compiled without any
// slow-path call in its loops; the FR = false instantiation (coordinates near the ends of the fp32 range, NaN / inf) is the
// same code on __fdiv_rn and the float bounds test, kept out of line.
// CACHE (trilinear only): the ray's iterations fit the per-warp value / mask cache; a ray that needs more (the cache is sized for
// unit-length directions through the grid's diagonal, launch_render_train) runs with CACHE = false and recomputes the
// interpolation in its reverse pass, as every trilinear ray did before the cache existed.
template <int MODE, bool FAST, bool FR, int SPL, bool PIPE, bool PEER, bool CACHE>
__device__ __forceinline__ float train_ray(const PlxRenderTrain& a, const Geo& g, const GradDst<PEER>& dst, const RayScratch& sc,
                                           const Ray& r, const float4 tgt, const int64_t ray, const int lane, const int k0, const int k1) {
    constexpr int W = 32 * SPL;
    const PlxMarch& m = a.march;
    int* lc = sc.lc;
    float* tcs = sc.tcs;
    const float bom = a.beta_over_m;
    const bool full = bom != 0.f;        // the sparsity term touches every in-bounds sample: no early stop
    const int n_it = k0 <= k1 ? (k1 - k0) / W + 1 : 0;

    // ---------------------------------------------------------------------------------------- forward
    float T = 1.f, T2 = 1.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = acc;
    bool opaque = false;                 // segment 1 ended on an alpha == 1 sample at (zl, zj)
    int zl = 0, zj = 0;
    int it = 0;
    // software pipeline: the cells of iteration it+1 are requested before iteration it is composited, so the gather
    // latency (L2 / HBM) overlaps the scans instead of stalling the warp at the first use
    // nearest: rawn = the raw cell (clamped when consumed), linn = its linear index.  trilinear: rawn = the value interpolated
    // from the eight (clamped) corners, linn = linear index of the floor corner; tgn / maskn = interpolation geometry and the
    // corners' clip pass-mask, which only the reverse pass consumes
    float4 rawn[SPL];
    int linn[SPL];
    TriGeom tgn[MODE == PLX_TRILINEAR ? SPL : 1];
    uint32_t maskn[MODE == PLX_TRILINEAR ? SPL : 1];
    auto fetch = [&](int i) {
        const int kb = k0 + i * W + lane * SPL;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            if (MODE == PLX_NEAREST) {
                linn[j] = fetch_nearest<FAST>(m, g, a.grid, r, FR, kb + j, kb + j <= k1, rawn[j]);
            } else {
                float nx, ny, nz;
                norm3<FAST>(m, g, r, FR, __fmul_rn(g.delta, (float)(kb + j)), nx, ny, nz);
                rawn[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                linn[j] = -1;
                maskn[j] = 0u;
                if (kb + j <= k1 && tri_geom(g, nx, ny, nz, tgn[j])) {
                    linn[j] = (tgn[j].lo[0] * g.ny + tgn[j].lo[1]) * g.nz + tgn[j].lo[2];
                    rawn[j] = tri_interp_mask<FAST>(m, g, a.grid, tgn[j], maskn[j]);
                }
            }
        }
    };
    constexpr bool CLAMP_ON_USE = MODE == PLX_NEAREST;   // trilinear clamps the corners, not the interpolated value
    if (PIPE && n_it > 0) fetch(0);
    for (; it < n_it; ++it) {
        float4 c[SPL];
        int lin[SPL];
        if (!PIPE) fetch(it);
#pragma unroll
        for (int j = 0; j < SPL; ++j) { c[j] = (CLAMP_ON_USE && g.clamp) ? clamp4(rawn[j]) : rawn[j]; lin[j] = linn[j]; }
        if (PIPE && it + 1 < n_it) fetch(it + 1);
        if (MODE == PLX_NEAREST) {
            if (SPL == 1) lc[it * W + lane] = lin[0];
            else *reinterpret_cast<int2*>(lc + it * W + lane * 2) = make_int2(lin[0], lin[SPL - 1]);
        } else if (CACHE) {
#pragma unroll
            for (int j = 0; j < SPL; ++j) sc.cv[it * W + lane * SPL + j] = c[j];
            if (SPL == 1) sc.mk[it * W + lane] = maskn[0];
            else *reinterpret_cast<uint2*>(sc.mk + it * W + lane * 2) = make_uint2(maskn[0], maskn[SPL - 1]);
        }
        if (lane == 0) tcs[it] = T;
        // empty space is the common case (a trained grid is mostly alpha == 0, fit() even starts from all zeros):
        // an iteration whose 32*SPL samples are all transparent changes neither T nor the pixel — skip its scan
        bool any_alpha = false;
#pragma unroll
        for (int j = 0; j < SPL; ++j) any_alpha = any_alpha || c[j].w != 0.f;
        if (!__any_sync(FULL, any_alpha)) continue;
        composite_iter<SPL>(c, lane, T, acc);
        if (T == 0.f && !full) {
            // first sample whose factor is exactly 0 (alpha == 1); none => the product merely underflowed
            int jz = SPL;
#pragma unroll
            for (int j = SPL - 1; j >= 0; --j) if (c[j].w == 1.f) jz = j;
            const unsigned hit = __ballot_sync(FULL, jz < SPL);
            if (hit) {
                opaque = true;
                zl = __ffs(hit) - 1;
                zj = __shfl_sync(FULL, jz, zl);
                float4 c2[SPL];
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    const bool after = lane > zl || (lane == zl && j > zj);
                    c2[j] = after ? c[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                composite_iter<SPL>(c2, lane, T2, acc2);
            }
            ++it;
            break;
        }
    }
    const int n_fwd = it;                // iterations of segment 1 (their indices are cached)
    if (opaque) {                        // keep compositing behind k* until that segment saturates or the range ends
        for (int i2 = n_fwd; i2 < n_it && T2 != 0.f; ++i2) {          // PIPE: iteration n_fwd is already in flight (rawn)
            float4 c[SPL];
            if (!PIPE) fetch(i2);
#pragma unroll
            for (int j = 0; j < SPL; ++j) c[j] = (CLAMP_ON_USE && g.clamp) ? clamp4(rawn[j]) : rawn[j];
            if (PIPE && i2 + 1 < n_it) fetch(i2 + 1);
            composite_iter<SPL>(c, lane, T2, acc2);
        }
    }
    acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);

    // ---------------------------------------------------------------------------------------- loss, scripts/train.py:156
    const float er = acc.x - tgt.x, eg = acc.y - tgt.y, eb = acc.z - tgt.z, ea = acc.w - tgt.w;
    const float4 gr = make_float4(er * a.grad_scale, eg * a.grad_scale, eb * a.grad_scale, ea * a.grad_scale);
    const float this_loss = (er * er + eg * eg + eb * eb + ea * ea) * a.loss_scale;
    if (lane == 0 && a.rgba) reinterpret_cast<float4*>(a.rgba)[ray] = acc;
    float s_star = 0.f;                  // colour behind k*, dotted with the pixel gradient
    if (opaque) {
        acc2.x = warp_sum(acc2.x); acc2.y = warp_sum(acc2.y); acc2.z = warp_sum(acc2.z); acc2.w = warp_sum(acc2.w);
        s_star = fmaf(acc2.x, gr.x, fmaf(acc2.y, gr.y, fmaf(acc2.z, gr.z, acc2.w * gr.w)));
    }

    // ---------------------------------------------------------------------------------------- reverse
    const bool any_grad = gr.x != 0.f || gr.y != 0.f || gr.z != 0.f || gr.w != 0.f || full;
    float carry = 0.f;                   // S behind the last visited sample (0: end of ray, or irrelevant behind k*)
    __syncwarp();
    auto refetch = [&](int i) {          // indices back from shared memory, cells requested (L1 / L2 hits mostly)
        if (MODE == PLX_TRILINEAR && !CACHE) { fetch(i); return; }       // geometry + gathers + interpolation all over again
        if (MODE == PLX_TRILINEAR) {         // geometry recomputed (no loads), value and corner masks back from shared memory
            const int kb = k0 + i * W + lane * SPL;
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                float nx, ny, nz;
                norm3<FAST>(m, g, r, FR, __fmul_rn(g.delta, (float)(kb + j)), nx, ny, nz);
                linn[j] = -1;
                if (kb + j <= k1 && tri_geom(g, nx, ny, nz, tgn[j])) linn[j] = (tgn[j].lo[0] * g.ny + tgn[j].lo[1]) * g.nz + tgn[j].lo[2];
                rawn[j] = sc.cv[i * W + lane * SPL + j];
            }
            if (SPL == 1) maskn[0] = sc.mk[i * W + lane];
            else { const uint2 q = *reinterpret_cast<const uint2*>(sc.mk + i * W + lane * 2); maskn[0] = q.x; maskn[SPL - 1] = q.y; }
            return;
        }
        if (SPL == 1) linn[0] = lc[i * W + lane];
        else { const int2 p = *reinterpret_cast<const int2*>(lc + i * W + lane * 2); linn[0] = p.x; linn[SPL - 1] = p.y; }
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            rawn[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (linn[j] >= 0) rawn[j] = FAST ? ldg_hint(reinterpret_cast<const float4*>(a.grid) + linn[j], g.pol)
                                             : cell_at<false>(m, g, a.grid, linn[j] / (g.ny * g.nz), (linn[j] / g.nz) % g.ny, linn[j] % g.nz, linn[j]);
        }
    };
    if (PIPE && n_fwd > 0 && any_grad) refetch(n_fwd - 1);
    for (int ib = n_fwd - 1; ib >= 0 && any_grad; --ib) {
        int lin[SPL];
        float4 raw[SPL], c[SPL];
        float v[SPL];
        if (!PIPE) refetch(ib);
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            lin[j] = linn[j];
            raw[j] = rawn[j];
            c[j] = (CLAMP_ON_USE && g.clamp) ? clamp4(raw[j]) : raw[j];
            v[j] = fmaf(c[j].x, gr.x, fmaf(c[j].y, gr.y, fmaf(c[j].z, gr.z, gr.w)));      // c_k . g_rgb + g_A
        }
        if (PIPE && ib > 0) refetch(ib - 1);     // next iteration's cells in flight during this iteration's scans
        bool any_alpha = false;
#pragma unroll
        for (int j = 0; j < SPL; ++j) any_alpha = any_alpha || c[j].w != 0.f;
        float behind[SPL], pf[SPL];
        float base;
        if (__any_sync(FULL, any_alpha)) {
            // lane aggregate of the affine maps s -> alpha v + (1 - alpha) s, last sample innermost
            float A = 0.f, B = 1.f;
#pragma unroll
            for (int j = SPL - 1; j >= 0; --j) { A = fmaf(1.f - c[j].w, A, c[j].w * v[j]); B *= 1.f - c[j].w; }
            behind[SPL - 1] = warp_behind(A, B, lane, carry);
#pragma unroll
            for (int j = SPL - 1; j >= 1; --j) behind[j - 1] = fmaf(1.f - c[j].w, behind[j], c[j].w * v[j]);
            if (opaque && ib == n_fwd - 1 && lane == zl) {
#pragma unroll
                for (int j = 0; j < SPL; ++j) if (j == zj) behind[j] = s_star;
            }
            // T_k inside the iteration
            pf[0] = 1.f;
#pragma unroll
            for (int j = 1; j < SPL; ++j) pf[j] = pf[j - 1] * (1.f - c[j - 1].w);
            float total;
            base = tcs[ib] * warp_excl_prod(pf[SPL - 1] * (1.f - c[SPL - 1].w), lane, total);
        } else {
            // all-transparent iteration: every map is the identity and every factor is 1 — T_k = T at the start of
            // the iteration, the colour behind each sample is the carried one; no scan needed
#pragma unroll
            for (int j = 0; j < SPL; ++j) { behind[j] = carry; pf[j] = 1.f; }
            base = tcs[ib];
        }
        float4 d[SPL];
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const float Tk = base * pf[j];
            const float wgt = c[j].w * Tk;
            d[j] = make_float4(wgt * gr.x, wgt * gr.y, wgt * gr.z, Tk * (v[j] - behind[j]));
            if (full) d[j].w += bom * (1.f / (c[j].w + 1e-4f) + 1.f / (1.f - c[j].w + 1e-4f));   // scripts/train.py:170-177
            if (CLAMP_ON_USE && g.clamp) { d[j].x *= pass01(raw[j].x); d[j].y *= pass01(raw[j].y); d[j].z *= pass01(raw[j].z); d[j].w *= pass01(raw[j].w); }
        }
        if (MODE == PLX_TRILINEAR) {
            // d = gradient w.r.t. the interpolated value; every corner receives d * (its lerp weight) through its own clip
            // pass-mask.  Two samples of a lane that share the floor cell share all eight corners: one reduction per corner.
            const bool same = SPL == 2 && lin[0] >= 0 && lin[0] == lin[SPL - 1] &&
                              tgn[0].hi[0] == tgn[SPL - 1].hi[0] && tgn[0].hi[1] == tgn[SPL - 1].hi[1] && tgn[0].hi[2] == tgn[SPL - 1].hi[2];
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                if (lin[j] < 0 || (same && j == SPL - 1)) continue;
                const bool any = d[j].x != 0.f || d[j].y != 0.f || d[j].z != 0.f || d[j].w != 0.f;
                const bool any2 = same && (d[SPL - 1].x != 0.f || d[SPL - 1].y != 0.f || d[SPL - 1].z != 0.f || d[SPL - 1].w != 0.f);
                if (!any && !any2) continue;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float w0 = tri_weight(tgn[j], k), w1 = same ? tri_weight(tgn[SPL - 1], k) : 0.f;
                    if (w0 == 0.f && w1 == 0.f) continue;
                    const uint32_t mk = maskn[j] >> (4 * k);
                    float4 o = make_float4(d[j].x * w0, d[j].y * w0, d[j].z * w0, d[j].w * w0);
                    if (same) { o.x = fmaf(d[SPL - 1].x, w1, o.x); o.y = fmaf(d[SPL - 1].y, w1, o.y); o.z = fmaf(d[SPL - 1].z, w1, o.z); o.w = fmaf(d[SPL - 1].w, w1, o.w); }
                    dst.add(tri_corner_lin(g, tgn[j], k), (mk & 1u) ? o.x : 0.f, (mk & 2u) ? o.y : 0.f, (mk & 4u) ? o.z : 0.f, (mk & 8u) ? o.w : 0.f);
                }
            }
        } else if (SPL == 1 && !PEER) {
            warp_scatter_add(dst.local, lin[0] >= 0, (int64_t)lin[0] * 4, d[0].x, d[0].y, d[0].z, d[0].w, lane, dst.pol);
        } else {
            // merge runs inside the lane, then one 16-byte reduction per surviving entry
#pragma unroll
            for (int j = SPL - 1; j >= 1; --j) {
                if (lin[j] >= 0 && lin[j] == lin[j - 1]) {
                    d[j - 1].x += d[j].x; d[j - 1].y += d[j].y; d[j - 1].z += d[j].z; d[j - 1].w += d[j].w;
                    lin[j] = -1;
                }
            }
            if (PEER) {
                // push exchange: every reduction is a 16-byte packet on NVLink, so runs that continue in the NEXT lane are
                // merged too — a lane hands the sum of its last entry to its neighbour when that one's first entry is the same
                // cell (samples enter cells monotonically along the ray, so equal cells are adjacent).  One round takes the
                // reductions per sample from 0.8 to ~0.6 at delta = pd / 2 (N = 8: 131 -> 116 us per step).  Not used for the
                // local gradient buffer: there the L2 absorbs the extra atomics and the shuffles cost more than they save (C2
                // step 89.5 -> 91.7 us, although the kernel alone profiles faster under ncu).  A handed-off entry keeps its
                // index with a zero sum: whatever the previous lane hands to IT in the same round is then reduced from here,
                // so nothing is lost however long the run is.
                int fi = -1, la = -1;                                   // first / last surviving entry of this lane
#pragma unroll
                for (int j = SPL - 1; j >= 0; --j) if (lin[j] >= 0) fi = j;
#pragma unroll
                for (int j = 0; j < SPL; ++j) if (lin[j] >= 0) la = j;
                int lin_f = -1, lin_l = -1;
                float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < SPL; ++j) { if (j == fi) lin_f = lin[j]; if (j == la) { lin_l = lin[j]; dl = d[j]; } }
                const int next_f = __shfl_down_sync(FULL, lin_f, 1);
                const bool send = lane < 31 && lin_l >= 0 && lin_l == next_f;
                const float ix = __shfl_up_sync(FULL, send ? dl.x : 0.f, 1), iy = __shfl_up_sync(FULL, send ? dl.y : 0.f, 1);
                const float iz = __shfl_up_sync(FULL, send ? dl.z : 0.f, 1), iw = __shfl_up_sync(FULL, send ? dl.w : 0.f, 1);
                const bool recv = __shfl_up_sync(FULL, (int)send, 1) != 0 && lane > 0;
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    if (send && j == la) d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (recv && j == fi) { d[j].x += ix; d[j].y += iy; d[j].z += iz; d[j].w += iw; }
                }
            }
#pragma unroll
            for (int j = 0; j < SPL; ++j)
                if (lin[j] >= 0 && (d[j].x != 0.f || d[j].y != 0.f || d[j].z != 0.f || d[j].w != 0.f))
                    dst.add(lin[j], d[j].x, d[j].y, d[j].z, d[j].w);
        }
    }
    __syncwarp();                        // the per-warp shared-memory cache is reused by the next ray
    return this_loss;
}

// this ray: precomputed, or generated here from (pose, uv) — src/ray_sampling.py:212-264
__device__ __forceinline__ void setup_ray(const PlxRenderTrain& a, int64_t ray, Ray& r, float4& tgt) {
    if (a.gen.uv) {
        const int R = a.gen.rays_per_cam;
        const int cam = (int)(ray / R);
        const float* P = a.gen.poses + (int64_t)cam * 16;
        const RayOut o = ray_from_uv(P, a.gen.fov, __ldg(a.gen.uv + ray * 2), __ldg(a.gen.uv + ray * 2 + 1), a.gen.img_h, a.gen.img_w);
        tgt = load_target(a.gen, ((int64_t)cam * a.gen.img_h + o.vp) * a.gen.img_w + o.up);
        r.ox = __ldg(P + 3); r.oy = __ldg(P + 7); r.oz = __ldg(P + 11);          // camera position, :159
        r.dx = o.dx; r.dy = o.dy; r.dz = o.dz;
    } else {
        r = load_ray(a.rays, ray);
        tgt = __ldg(reinterpret_cast<const float4*>(a.targets) + ray);
    }
}

// the rare ray outside the hoisted-division range (coordinates near the ends of the fp32 range, NaN / inf), or (trilinear) longer
// than the per-warp cache: the same march on the IEEE division and the float bounds test, without the value cache.  Out of line and self-contained (it rebuilds the ray and the geometry from the
// kernel argument), so the common path pays neither instructions nor registers nor stack traffic for it.
template <int MODE, bool FAST, int SPL, bool PEER>
__device__ __noinline__ float train_ray_slow(const PlxRenderTrain* ap, float* const* peers, const RayScratch sc, int64_t ray, int lane) {
    const PlxRenderTrain& a = *ap;
    const Geo g = make_geo(a.march);
    GradDst<PEER> dst;
    dst.local = a.grad_grid; dst.peers = peers; dst.owner_mul = a.peer_grad.owner_mul; dst.pol = g.pol_grad;
    Ray r;
    float4 tgt;
    setup_ray(a, ray, r, tgt);
    int k0, k1;
    clip_range(a.march, r, k0, k1);
    return train_ray<MODE, FAST, false, SPL, false, PEER, false>(a, g, dst, sc, r, tgt, ray, lane, k0, k1);
}

template <int MODE, bool FAST, int SPL, bool PIPE, bool PEER>
__global__ void __launch_bounds__(MODE == PLX_NEAREST ? 128 : 64, 8) k_render_train(const __grid_constant__ PlxRenderTrain a, int lin_words, int warp_words) {
    extern __shared__ __align__(16) int s_dyn[];   // per warp (warp_words ints, 16-byte multiple): lin cache [lin_words], then chunk transmittances
    __shared__ float s_loss;
    __shared__ int s_done;
    __shared__ float* s_peer[PLX_MAX_PEERS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const PlxMarch& m = a.march;
    if (threadIdx.x == 0) { s_loss = 0.f; s_done = 0; }
    grid_dependency_wait();                  // launched behind the optimiser's tail: its parameter stores are visible from here on
    if (PEER && threadIdx.x < PLX_MAX_PEERS) s_peer[threadIdx.x] = a.peer_grad.grads[threadIdx.x];
    peer_wait(a.sync);                       // multi-GPU: every peer has stored its slab of the previous step's parameters
    __syncthreads();                         // every warp arrives here at once: free
    float ray_loss = 0.f;
    // one ray per warp, ray = block * warps + warp (claiming rays dynamically from a global counter was measured slower on
    // every config: 61 vs 48 us on C2 even with the ticket drawn one ray ahead)
    const int64_t ray = (int64_t)blockIdx.x * wpb + wib;
    if (ray < a.rays.n_rays) {
        // per warp: nearest    [lin cache: lin_words ints][chunk transmittances]
        //           trilinear  [chunk transmittances][values: lin_words float4][corner masks: lin_words words]
        RayScratch sc;
        int* const base = s_dyn + wib * warp_words;
        if (MODE == PLX_NEAREST) {
            sc.lc = base;
            sc.tcs = reinterpret_cast<float*>(base + lin_words);
            sc.cv = nullptr; sc.mk = nullptr;
        } else {                             // lin_words = cache capacity in samples (may be less than the ray can visit, see `fits`)
            const int tcs_words = warp_words - 5 * lin_words;
            sc.lc = nullptr;
            sc.tcs = reinterpret_cast<float*>(base);
            sc.cv = reinterpret_cast<float4*>(base + tcs_words);
            sc.mk = reinterpret_cast<uint32_t*>(base + tcs_words + 4 * lin_words);
        }
        Ray r;
        float4 tgt;
        setup_ray(a, ray, r, tgt);
        int k0, k1;
        clip_range(m, r, k0, k1);
        // trilinear: lin_words = the samples the per-warp cache holds
        const bool fits = MODE == PLX_NEAREST || k0 > k1 || ((k1 - k0) / (32 * SPL) + 1) * (32 * SPL) <= lin_words;
        if (FAST && fits && ray_in_fast_range(m, r)) {
            const Geo g = make_geo(m);
            GradDst<PEER> dst;
            dst.local = a.grad_grid; dst.peers = s_peer; dst.owner_mul = a.peer_grad.owner_mul; dst.pol = g.pol_grad;
            ray_loss = train_ray<MODE, FAST, true, SPL, PIPE, PEER, true>(a, g, dst, sc, r, tgt, ray, lane, k0, k1);
        } else {
            ray_loss = train_ray_slow<MODE, FAST, SPL, PEER>(&a, s_peer, sc, ray, lane);
        }
    }
    // ---- loss: shared-memory partial per block, the last warp to finish adds it to the global accumulator; that warp
    // also counts the block as done for the cross-GPU signal (its lanes' gradient reductions are ordered before lane 0's
    // fence by the __syncwarp above, the other warps' by their own fence before they bump s_done)
    float* const loss_dst = a.loss && a.step_dev ? a.loss + ((__ldg(a.step_dev) + 1) & 1) : a.loss;      // graph replay: slot by device step parity
    if (lane == 0 && (a.loss || a.sync.signal_epoch > 0)) {
        if (a.loss) atomicAdd(&s_loss, ray_loss);
        if (a.sync.signal_epoch > 0) { if (PEER) __threadfence_system(); else __threadfence(); } else __threadfence_block();
        if (atomicAdd(&s_done, 1) == wpb - 1) {
            if (a.loss) atomicAdd(loss_dst, atomicAdd(&s_loss, 0.f));
            peer_signal(a.sync);
        }
    }
}

// shared memory the fused kernel needs per block; 0 = not launchable (fall back to K1 + K2)
// iterations of 32 * spl samples a ray can visit: all S of them, or (cache_it: the trilinear value cache) what a UNIT-length
// direction can spend inside the grid grown by the clip margins — the grid's diagonal; longer rays take the uncached path
static int train_iterations(const PlxMarch& m, int spl, bool cache_it) {
    const int W = 32 * spl;
    int n_it = (m.num_samples + W - 1) / W + 1;
    if (cache_it && m.delta_step > 0.f && m.points_distance > 0.f) {
        if (tuning().train_cache_it > 0) return tuning().train_cache_it < n_it ? tuning().train_cache_it : n_it;
        const double pd = m.points_distance;
        double d2 = 0.0;
        for (int n : {m.nx, m.ny, m.nz}) { const double e = pd * (n + 8); d2 += e * e; }       // 4 cells of margin per side (clip_range: < 3)
        const double samples = std::sqrt(d2) / m.delta_step + 8.0;
        if (samples < 1e9) { const int fit = (int)std::ceil(samples / W) + 1; if (fit < n_it) n_it = fit; }
    }
    return n_it;
}

size_t render_train_smem(const PlxMarch& m, int spl, int wpb) {
    const int W = 32 * spl;
    const bool tri = m.mode == PLX_TRILINEAR;
    const int n_it_max = train_iterations(m, spl, false);
    const size_t sample_words = tri ? (size_t)train_iterations(m, spl, true) * W * 5      // value (4) + corner masks (1) per cached sample
                                    : (size_t)n_it_max * W;                               // the cell index of every sample
    return (size_t)wpb * (sample_words + ((n_it_max + 3) & ~3)) * sizeof(int);
}

bool render_train_supported(const PlxRenderTrain& a) {
    if (a.march.mode == PLX_TRILINEAR && !fast_ok(a.march, a.grid)) return false;     // trilinear: contiguous grids only (K1 + K2 otherwise)
    if ((int64_t)a.march.nx * a.march.ny * a.march.nz >= (1ll << 31) / 4) return false;
    return render_train_smem(a.march, 2, 1) <= 200 * 1024;
}

template <int MODE, bool FAST, int SPL, bool PIPE, bool PEER>
static cudaError_t launch_train_inst(const PlxRenderTrain& a, unsigned blocks, int wpb, size_t smem, int lin_words, int warp_words,
                                     cudaStream_t st) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_render_train<MODE, FAST, SPL, PIPE, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return launch_pdl(k_render_train<MODE, FAST, SPL, PIPE, PEER>, blocks, (unsigned)(wpb * 32), smem, st, a, lin_words, warp_words);
}

// Launch shape (measured on C2 / B200): 4 warps per block, 8 blocks per SM (64 registers), 2 samples per lane
// from 128 samples per ray up; software-pipelined gathers when grid + gradient fit the L2 (+5 % on C2) and off on grids far
// beyond it, where the extra requests in flight only add DRAM queueing (-8 % on the 256^3 sweep).  Trilinear: 2 warps per
// block, 8 blocks per SM at 120 registers (9 blocks at 96 registers spill: 242 vs 217 us per C2 step), the value cache sized
// from the grid's diagonal instead of S (C2: 9 instead of 11 iterations = 16 instead of 14 resident warps: 237 -> 217 us).
cudaError_t launch_render_train(const PlxRenderTrain& a_in, cudaStream_t st) {
    if (a_in.rays.n_rays == 0) return cudaSuccess;
    PlxRenderTrain a = a_in;
    a.march.flags = (a.march.flags & 0xffffu) | l2_keep_flags(a.march);
    const int spl = a.march.num_samples >= 128 ? 2 : 1;
    const bool tri = a.march.mode == PLX_TRILINEAR;
    // trilinear keeps 20 B per cached sample in shared memory: two warps per block so that the blocks pack the SM (C2: 23 KB each)
    int wpb = tri && tuning().train_wpb > 2 ? 2 : tuning().train_wpb;
    while (wpb > 1 && render_train_smem(a.march, spl, wpb) > 200 * 1024) wpb >>= 1;
    const size_t smem = render_train_smem(a.march, spl, wpb);
    const int W = 32 * spl;
    const int n_it_max = train_iterations(a.march, spl, false);
    const int lin_words = train_iterations(a.march, spl, tri) * W, warp_words = lin_words * (tri ? 5 : 1) + ((n_it_max + 3) & ~3);
    const unsigned blocks = (unsigned)((a.rays.n_rays + wpb - 1) / wpb);
    const bool fast = fast_ok(a.march, a.grid);
    const bool pipe = l2_keep_ok((int64_t)a.march.nx * a.march.ny * a.march.nz);
    const bool peer = a.peer_grad.world > 0;
    if (a.march.mode == PLX_TRILINEAR) {     // 8 gathers per sample: no software pipelining (the corners would not fit the registers)
        if (peer) { if (spl == 2) return launch_train_inst<PLX_TRILINEAR, true, 2, false, true>(a, blocks, wpb, smem, lin_words, warp_words, st);
                    return launch_train_inst<PLX_TRILINEAR, true, 1, false, true>(a, blocks, wpb, smem, lin_words, warp_words, st); }
        if (spl == 2) return launch_train_inst<PLX_TRILINEAR, true, 2, false, false>(a, blocks, wpb, smem, lin_words, warp_words, st);
        return launch_train_inst<PLX_TRILINEAR, true, 1, false, false>(a, blocks, wpb, smem, lin_words, warp_words, st);
    }
#define PLX_TRAIN(F, S, P, E) return launch_train_inst<PLX_NEAREST, F, S, P, E>(a, blocks, wpb, smem, lin_words, warp_words, st)
    if (peer) {                              // validated by the ABI layer: contiguous grid only
        if (spl == 2) { if (pipe) PLX_TRAIN(true, 2, true, true); else PLX_TRAIN(true, 2, false, true); }
        else          { if (pipe) PLX_TRAIN(true, 1, true, true); else PLX_TRAIN(true, 1, false, true); }
    }
    if (fast) {
        if (spl == 2) { if (pipe) PLX_TRAIN(true, 2, true, false); else PLX_TRAIN(true, 2, false, false); }
        else          { if (pipe) PLX_TRAIN(true, 1, true, false); else PLX_TRAIN(true, 1, false, false); }
    }
    if (spl == 2) { if (pipe) PLX_TRAIN(false, 2, true, false); else PLX_TRAIN(false, 2, false, false); }
    else          { if (pipe) PLX_TRAIN(false, 1, true, false); else PLX_TRAIN(false, 1, false, false); }
#undef PLX_TRAIN
}

}  // namespace plx
