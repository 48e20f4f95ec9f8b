// plx_march.cuh — per-sample march primitives shared by K1/K2 (plx_render.cu) and the fused training kernel
// (plx_train.cu): launch-invariant geometry, exact normalised coordinates, nearest / trilinear lookup.
//
// Template axis FAST: contiguous 16-byte-aligned grid + divisor in the hoisted-reciprocal range: cells are addressed
// by their linear index with one 128-bit load and the quotient costs 3 FMAs (plx_device.cuh);
// !FAST = arbitrary strides (channel-planar pooled grids), scalar loads, __fdiv_rn.
#pragma once
#include <cstdlib>

#include "plx_device.cuh"
#include "plx_launch.h"

namespace plx {


// launch-invariant geometry, derived once per thread from the PlxMarch argument
struct Geo {
    float fnx, fny, fnz;     // grid dims as floats (the in-bounds test runs on the rounded float)
    int nx, ny, nz;
    float gx, gy, gz, delta;
    FastDiv div;
    bool clamp;
    uint64_t pol, pol_grad;  // L2 cache policies for grid reads / gradient reductions
};

__device__ __forceinline__ Geo make_geo(const PlxMarch& m) {
    Geo g;
    g.fnx = (float)m.nx; g.fny = (float)m.ny; g.fnz = (float)m.nz;
    g.nx = m.nx; g.ny = m.ny; g.nz = m.nz;
    g.gx = m.gmin[0]; g.gy = m.gmin[1]; g.gz = m.gmin[2];
    g.delta = m.delta_step;
    g.div = make_fastdiv(m.points_distance);
    g.clamp = (m.flags & PLX_CLAMP01) != 0;
    g.pol = l2_policy((m.flags & PLX_FLAG_KEEP_GRID) != 0);
    g.pol_grad = l2_policy((m.flags & PLX_FLAG_KEEP_GRAD) != 0);
    return g;
}

// every numerator of the ray stays below the hoisted-division range (tiny numerators only ever round to index 0)
__device__ __forceinline__ bool ray_in_fast_range(const PlxMarch& m, const Ray& r) {
    const float reach = fabsf(m.delta_step) * (float)m.num_samples;
    const float big = fmaxf(fmaxf(fabsf(r.ox), fabsf(r.oy)), fabsf(r.oz)) +
                      reach * fmaxf(fmaxf(fabsf(r.dx), fabsf(r.dy)), fabsf(r.dz)) +
                      fmaxf(fmaxf(fabsf(m.gmin[0]), fabsf(m.gmin[1])), fabsf(m.gmin[2]));
    return big <= 1e17f;      // false for NaN / inf too
}

template <bool FAST>
__device__ __forceinline__ void norm3(const PlxMarch& m, const Geo& g, const Ray& r, bool fast_ray, float t, float& nx,
                                      float& ny, float& nz) {
    const float x = __fsub_rn(__fadd_rn(r.ox, __fmul_rn(r.dx, t)), g.gx);     // src/ray_sampling.py:167, :13
    const float y = __fsub_rn(__fadd_rn(r.oy, __fmul_rn(r.dy, t)), g.gy);
    const float z = __fsub_rn(__fadd_rn(r.oz, __fmul_rn(r.dz, t)), g.gz);
    if (FAST && fast_ray) {
        nx = fdiv_hoisted(x, g.div); ny = fdiv_hoisted(y, g.div); nz = fdiv_hoisted(z, g.div);
    } else {
        nx = __fdiv_rn(x, m.points_distance); ny = __fdiv_rn(y, m.points_distance); nz = __fdiv_rn(z, m.points_distance);
    }
}

template <bool FAST>
__device__ __forceinline__ float4 cell_at(const PlxMarch& m, const Geo& g, const float* __restrict__ grid, int ix, int iy, int iz, int lin) {
    if (FAST) return ldg_hint(reinterpret_cast<const float4*>(grid) + lin, g.pol);
    const int64_t off = ix * m.sx + iy * m.sy + iz * m.sz;
    return make_float4(__ldg(grid + off), __ldg(grid + off + m.sc), __ldg(grid + off + 2 * m.sc), __ldg(grid + off + 3 * m.sc));
}

__device__ __forceinline__ float4 clamp4(float4 c) {
    return make_float4(__saturatef(c.x), __saturatef(c.y), __saturatef(c.z), __saturatef(c.w));
}

// round-half-even + in-bounds test of the three normalised coordinates (src/grid_functions.py:111, :58-61).
// Reference form: r = rint(n) as float, 0 <= r < dim per axis.  For a ray in the fast range every coordinate is finite,
// so the same decision is one conversion and one unsigned compare per axis: __float2int_rn rounds half to even like
// rintf, saturates beyond int32 (-> fails the unsigned test, as the float test would) and maps -0.0 to 0 (inside).
// Anything that could be NaN / inf takes the float form.
__device__ __forceinline__ bool nearest_cell(const Geo& g, bool exact_int, float nx, float ny, float nz, int& ix, int& iy, int& iz) {
    if (exact_int) {
        ix = __float2int_rn(nx); iy = __float2int_rn(ny); iz = __float2int_rn(nz);
        return (unsigned)ix < (unsigned)g.nx && (unsigned)iy < (unsigned)g.ny && (unsigned)iz < (unsigned)g.nz;
    }
    const float rx = rintf(nx), ry = rintf(ny), rz = rintf(nz);
    const bool inb = rx >= 0.f && rx < g.fnx && ry >= 0.f && ry < g.fny && rz >= 0.f && rz < g.fnz;
    ix = inb ? (int)rx : 0; iy = inb ? (int)ry : 0; iz = inb ? (int)rz : 0;
    return inb;
}

// One sample's lookup result.
struct Sample {
    float4 c;        // (clamped) value, 0 when out of bounds
    float4 raw;      // nearest mode: the unclamped cell (for the clip pass-mask)
    bool inb;        // the reference's mask (True = inside)
    int lin;         // linear index (ix*ny+iy)*nz+iz of the nearest / floor-corner cell, -1 when out of bounds
};

// ---- nearest neighbour: src/grid_functions.py:111 (round half even), :58-61 (mask) -------------------------
template <bool FAST>
__device__ __forceinline__ Sample lookup_nearest(const PlxMarch& m, const Geo& g, const float* __restrict__ grid, float nx,
                                                 float ny, float nz, bool valid, bool need_value, bool exact_int) {
    Sample s;
    s.c = make_float4(0.f, 0.f, 0.f, 0.f);
    s.raw = s.c;
    int ix, iy, iz;
    s.inb = nearest_cell(g, exact_int, nx, ny, nz, ix, iy, iz) && valid;
    s.lin = -1;
    if (s.inb) {
        s.lin = (ix * g.ny + iy) * g.nz + iz;
        if (need_value) {
            s.raw = cell_at<FAST>(m, g, grid, ix, iy, iz, s.lin);
            s.c = g.clamp ? clamp4(s.raw) : s.raw;
        }
    }
    return s;
}

// nearest lookup that only ISSUES the load: returns the linear index (-1 = out of bounds) and the raw, unclamped cell.
// The caller clamps when it consumes the value, so the load can stay in flight across other work (software pipelining).
template <bool FAST>
__device__ __forceinline__ int fetch_nearest(const PlxMarch& m, const Geo& g, const float* __restrict__ grid, const Ray& r,
                                             bool fast_ray, int k, bool valid, float4& raw) {
    const float t = __fmul_rn(g.delta, (float)k);
    float nx, ny, nz;
    norm3<FAST>(m, g, r, fast_ray, t, nx, ny, nz);
    int ix, iy, iz;
    const bool inb = nearest_cell(g, FAST && fast_ray, nx, ny, nz, ix, iy, iz) && valid;
    raw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!inb) return -1;
    const int lin = (ix * g.ny + iy) * g.nz + iz;
    raw = cell_at<FAST>(m, g, grid, ix, iy, iz, lin);
    return lin;
}

// ---- trilinear: SURVEY.md §8a row T ---------------------------------------------------------------------------
struct TriGeom {
    int lo[3], hi[3];     // floor / wrapped ceil index per axis
    float f[3];           // frac per axis
};

__device__ __forceinline__ bool tri_geom(const Geo& g, float nx, float ny, float nz, TriGeom& t) {
    const float n[3] = {nx, ny, nz};
    const float dim[3] = {g.fnx, g.fny, g.fnz};
    bool inb = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) inb = inb && (n[a] >= 0.f) && (n[a] < dim[a]);      // float test, :58-61
    if (!inb) return false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float fl = floorf(n[a]);
        const float ce = ceilf(n[a]);
        t.lo[a] = (int)fl;
        t.hi[a] = ce >= dim[a] ? 0 : (int)ce;                  // periodic wrap of the ceil corner, :75-77
        t.f[a] = __fsub_rn(n[a], fl);                          // torch.frac for n >= 0, :29
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

template <bool FAST>
__device__ __forceinline__ float4 tri_cell(const PlxMarch& m, const Geo& g, const float* __restrict__ grid, int ix, int iy, int iz) {
    const float4 c = cell_at<FAST>(m, g, grid, ix, iy, iz, (ix * g.ny + iy) * g.nz + iz);
    return g.clamp ? clamp4(c) : c;
}

template <bool FAST>
__device__ __forceinline__ float4 tri_interp(const PlxMarch& m, const Geo& g, const float* __restrict__ grid, const TriGeom& t) {
    // x-lerp of the four (y,z) edges, then y, then z — corner order of src/grid_functions.py:238-243
    const float4 x_cc = lerp4(tri_cell<FAST>(m, g, grid, t.hi[0], t.hi[1], t.hi[2]), tri_cell<FAST>(m, g, grid, t.lo[0], t.hi[1], t.hi[2]), t.f[0]);
    const float4 x_cf = lerp4(tri_cell<FAST>(m, g, grid, t.hi[0], t.hi[1], t.lo[2]), tri_cell<FAST>(m, g, grid, t.lo[0], t.hi[1], t.lo[2]), t.f[0]);
    const float4 x_fc = lerp4(tri_cell<FAST>(m, g, grid, t.hi[0], t.lo[1], t.hi[2]), tri_cell<FAST>(m, g, grid, t.lo[0], t.lo[1], t.hi[2]), t.f[0]);
    const float4 x_ff = lerp4(tri_cell<FAST>(m, g, grid, t.hi[0], t.lo[1], t.lo[2]), tri_cell<FAST>(m, g, grid, t.lo[0], t.lo[1], t.lo[2]), t.f[0]);
    const float4 y_c = lerp4(x_cc, x_fc, t.f[1]);
    const float4 y_f = lerp4(x_cf, x_ff, t.f[1]);
    return lerp4(y_c, y_f, t.f[2]);
}

// the same interpolation that also reports, for the backward pass, which channels of which corner pass the clip gradient:
// bit (corner * 4 + channel) of `mask`, corner = (x floor ? 4 : 0) | (y floor ? 2 : 0) | (z floor ? 1 : 0)   (scripts/train.py:146)
template <bool FAST>
__device__ __forceinline__ float4 tri_interp_mask(const PlxMarch& m, const Geo& g, const float* __restrict__ grid, const TriGeom& t,
                                                  uint32_t& mask) {
    float4 c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {            // all eight gathers in flight before the first use
        const int ix = (k & 4) ? t.lo[0] : t.hi[0], iy = (k & 2) ? t.lo[1] : t.hi[1], iz = (k & 1) ? t.lo[2] : t.hi[2];
        c[k] = cell_at<FAST>(m, g, grid, ix, iy, iz, (ix * g.ny + iy) * g.nz + iz);
    }
    mask = 0xffffffffu;
    if (g.clamp) {
        mask = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            mask |= (c[k].x >= 0.f && c[k].x <= 1.f ? 1u : 0u) << (4 * k);
            mask |= (c[k].y >= 0.f && c[k].y <= 1.f ? 2u : 0u) << (4 * k);
            mask |= (c[k].z >= 0.f && c[k].z <= 1.f ? 4u : 0u) << (4 * k);
            mask |= (c[k].w >= 0.f && c[k].w <= 1.f ? 8u : 0u) << (4 * k);
            c[k] = clamp4(c[k]);
        }
    }
    const float4 x_cc = lerp4(c[0], c[4], t.f[0]), x_cf = lerp4(c[1], c[5], t.f[0]);
    const float4 x_fc = lerp4(c[2], c[6], t.f[0]), x_ff = lerp4(c[3], c[7], t.f[0]);
    return lerp4(lerp4(x_cc, x_fc, t.f[1]), lerp4(x_cf, x_ff, t.f[1]), t.f[2]);
}

// weight of corner k in the interpolated value = d value / d corner (the product of the three lerp weights)
__device__ __forceinline__ float tri_weight(const TriGeom& t, int k) {
    return ((k & 4) ? 1.f - t.f[0] : t.f[0]) * ((k & 2) ? 1.f - t.f[1] : t.f[1]) * ((k & 1) ? 1.f - t.f[2] : t.f[2]);
}
__device__ __forceinline__ int tri_corner_lin(const Geo& g, const TriGeom& t, int k) {
    return (((k & 4) ? t.lo[0] : t.hi[0]) * g.ny + ((k & 2) ? t.lo[1] : t.hi[1])) * g.nz + ((k & 1) ? t.lo[2] : t.hi[2]);
}

template <int MODE, bool FAST>
__device__ __forceinline__ Sample lookup(const PlxMarch& m, const Geo& g, const float* __restrict__ grid, const Ray& r,
                                         bool fast_ray, int k, bool valid, bool need_value, float& t, TriGeom& tg) {
    t = __fmul_rn(g.delta, (float)k);                                       // src/ray_sampling.py:161
    float nx, ny, nz;
    norm3<FAST>(m, g, r, fast_ray, t, nx, ny, nz);
    if (MODE == PLX_NEAREST) return lookup_nearest<FAST>(m, g, grid, nx, ny, nz, valid, need_value, FAST && fast_ray);
    Sample s;
    s.c = make_float4(0.f, 0.f, 0.f, 0.f);
    s.raw = s.c;
    s.lin = -1;
    s.inb = valid && tri_geom(g, nx, ny, nz, tg);
    if (s.inb) {
        s.lin = (tg.lo[0] * g.ny + tg.lo[1]) * g.nz + tg.lo[2];
        if (need_value) s.c = tri_interp<FAST>(m, g, grid, tg);
    }
    return s;
}


// host side: which of grid / gradient get the evict_last tag for this launch (PLX_L2_KEEP: 0 none, 1 both, 2 grid only)
static inline uint32_t l2_keep_flags(const PlxMarch& m) {
    static const int mode = [] { const char* e = std::getenv("PLX_L2_KEEP"); return e ? std::atoi(e) : 2; }();
    if (mode == 0 || !l2_keep_ok((int64_t)m.nx * m.ny * m.nz)) return 0;
    return mode == 2 ? PLX_FLAG_KEEP_GRID : (PLX_FLAG_KEEP_GRID | PLX_FLAG_KEEP_GRAD);
}

// host-side predicate for the FAST instantiations
static inline bool fast_ok(const PlxMarch& m, const float* grid) {
    return m.sc == 1 && m.sz == 4 && m.sy == 4 * (int64_t)m.nz && m.sx == 4 * (int64_t)m.nz * m.ny &&
           ((uintptr_t)grid % 16 == 0) && fastdiv_ok(m.points_distance);
}

}  // namespace plx
