// plx_eager.cu — per-function kernels with the reference's materialised-tensor semantics.
//
// These back the stand-alone Python functions (src/ray_sampling.py, src/grid_functions.py) when a caller uses them
// outside the fused march: ray generation, sample placement, normalisation, nearest / trilinear lookup (+ autograd),
// compositing (+ autograd).  Same exact-fp32 arithmetic as the fused kernels (plx_device.cuh).
#include "plx_raygen.cuh"
#include "plx_launch.h"

namespace plx {

static inline unsigned blocks_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = 148 * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------------------------------
// generate_rays_batched — src/ray_sampling.py:195-264
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_generate_rays(const PlxRayGen gen, int n_side, float* __restrict__ dirs,
                                                       float* __restrict__ targets) {
    const int R = gen.rays_per_cam;
    const int64_t n = (int64_t)gen.n_cams * R;
    for (int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ray < n; ray += (int64_t)gridDim.x * blockDim.x) {
        const int cam = (int)(ray / R), j = (int)(ray % R);
        float u, v;
        if (gen.uv) { u = __ldg(gen.uv + ray * 2); v = __ldg(gen.uv + ray * 2 + 1); }
        else { u = linspace01(j / n_side, n_side); v = linspace01(j % n_side, n_side); }     // cartesian_prod, u-major (:223)
        const RayOut o = ray_from_uv(gen.poses + (int64_t)cam * 16, gen.fov, u, v, gen.img_h, gen.img_w);
        if (targets)                                                                                      // :248
            reinterpret_cast<float4*>(targets)[ray] = load_target(gen, ((int64_t)cam * gen.img_h + o.vp) * gen.img_w + o.up);
        dirs[ray * 3 + 0] = o.dx;
        dirs[ray * 3 + 1] = o.dy;
        dirs[ray * 3 + 2] = o.dz;
    }
}

cudaError_t launch_generate_rays(const PlxRayGen& gen, int n_side, float* dirs, float* targets, cudaStream_t st) {
    const int64_t n = (int64_t)gen.n_cams * gen.rays_per_cam;
    if (n == 0) return cudaSuccess;
    k_generate_rays<<<blocks_for(n, 256), 256, 0, st>>>(gen, n_side, dirs, targets);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// sample placement / normalisation — src/ray_sampling.py:161-167, :13
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sample_points(const PlxRays rays, int S, float delta, float* __restrict__ out) {
    const int64_t total = rays.n_rays * S * 3;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int axis = (int)(e % 3);
        const int64_t sample = e / 3;
        const int64_t ray = sample / S;
        const int k = (int)(sample % S) + 1;
        const float o = __ldg(rays.origins + (ray / rays.rays_per_origin) * rays.origin_stride + axis * rays.origin_comp_stride);
        const float d = __ldg(rays.dirs + ray * 3 + axis);
        out[e] = __fadd_rn(o, __fmul_rn(d, step_t(delta, k)));
    }
}

cudaError_t launch_sample_points(const PlxRays& rays, int S, float delta, float* out, cudaStream_t st) {
    const int64_t total = rays.n_rays * S * 3;
    if (total == 0) return cudaSuccess;
    k_sample_points<<<blocks_for(total, 256), 256, 0, st>>>(rays, S, delta, out);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_normalize_points(const float* __restrict__ in, int64_t total, float gx, float gy,
                                                          float gz, float pd, float* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int axis = (int)(e % 3);
        const float g = axis == 0 ? gx : (axis == 1 ? gy : gz);
        out[e] = __fdiv_rn(__fsub_rn(in[e], g), pd);
    }
}

cudaError_t launch_normalize_points(const float* in, int64_t m, float gx, float gy, float gz, float pd, float* out,
                                    cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    k_normalize_points<<<blocks_for(m * 3, 256), 256, 0, st>>>(in, m * 3, gx, gy, gz, pd, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// get_nearest_voxels — src/grid_functions.py:103-114 (+ find_out_of_bound :47-63, fix_out_of_bounds :66-79)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t pymod(int64_t a, int64_t n) {      // python-style non-negative modulo (A5)
    int64_t r = a % n;
    return r < 0 ? r + n : r;
}

struct Dims { int32_t n[3]; };
struct Strides { int64_t s[4]; };

__device__ __forceinline__ bool nearest_index(const float* __restrict__ ns, int64_t i, const Dims& d, int64_t w[3]) {
    bool inb = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int64_t r = __float2ll_rn(ns[i * 3 + a]);      // round half even, then .to(torch.long) (:111)
        inb = inb && r >= 0 && r < d.n[a];
        w[a] = pymod(r, d.n[a]);
    }
    return inb;
}

__global__ void __launch_bounds__(256) k_gather_nearest(const float* __restrict__ ns, int64_t m, const float* __restrict__ grid,
                                                        Dims d, Strides s, float* __restrict__ vals,
                                                        uint8_t* __restrict__ inbounds, int64_t* __restrict__ idx_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t w[3];
        const bool inb = nearest_index(ns, i, d, w);
        const int64_t off = w[0] * s.s[0] + w[1] * s.s[1] + w[2] * s.s[2];
        const float4 c = make_float4(__ldg(grid + off), __ldg(grid + off + s.s[3]), __ldg(grid + off + 2 * s.s[3]),
                                     __ldg(grid + off + 3 * s.s[3]));
        reinterpret_cast<float4*>(vals)[i] = c;
        if (inbounds) inbounds[i] = inb ? 1 : 0;
        if (idx_out) { idx_out[i * 3] = w[0]; idx_out[i * 3 + 1] = w[1]; idx_out[i * 3 + 2] = w[2]; }
    }
}

cudaError_t launch_gather_nearest(const float* ns, int64_t m, const float* grid, const int32_t* dims,
                                  const int64_t* strides, float* vals, uint8_t* inb, int64_t* idx_out, cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    Dims d{{dims[0], dims[1], dims[2]}};
    Strides s{{strides[0], strides[1], strides[2], strides[3]}};
    k_gather_nearest<<<blocks_for(m, 256), 256, 0, st>>>(ns, m, grid, d, s, vals, inb, idx_out);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_gather_nearest_bwd(const float* __restrict__ ns, int64_t m,
                                                            const float* __restrict__ gv, Dims d, float* __restrict__ gg) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t w[3];
        nearest_index(ns, i, d, w);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gv) + i);
        if (g.x != 0.f || g.y != 0.f || g.z != 0.f || g.w != 0.f)
            red_add_v4(gg + ((w[0] * d.n[1] + w[1]) * d.n[2] + w[2]) * 4, g.x, g.y, g.z, g.w);
    }
}

cudaError_t launch_gather_nearest_bwd(const float* ns, int64_t m, const float* grad_vals, const int32_t* dims,
                                      float* grad_grid, cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    Dims d{{dims[0], dims[1], dims[2]}};
    k_gather_nearest_bwd<<<blocks_for(m, 256), 256, 0, st>>>(ns, m, grad_vals, d, grad_grid);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// trilinear — src/grid_functions.py:220-246 (corners), :66-79 (wrap), :7-44 (interpolation)
// ---------------------------------------------------------------------------------------------------------------
struct TriIdx {
    int64_t lo[3], hi[3];
    float f[3];
    bool inb;
};

__device__ __forceinline__ TriIdx tri_index(const float* __restrict__ ns, int64_t i, const Dims& d) {
    TriIdx t;
    t.inb = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float n = ns[i * 3 + a];
        t.inb = t.inb && n >= 0.f && n < (float)d.n[a];
        t.lo[a] = pymod((int64_t)floorf(n), d.n[a]);
        t.hi[a] = pymod((int64_t)ceilf(n), d.n[a]);
        t.f[a] = __fsub_rn(n, truncf(n));                    // torch.frac keeps the sign (:29)
    }
    return t;
}

__device__ __forceinline__ float lerp_hl(float hi, float lo, float f) {
    return __fadd_rn(__fmul_rn(hi, f), __fmul_rn(lo, __fsub_rn(1.f, f)));
}

__global__ void __launch_bounds__(256) k_trilinear_fwd(const float* __restrict__ ns, int64_t m, const float* __restrict__ grid,
                                                       Dims d, Strides s, int masked, float* __restrict__ vals,
                                                       uint8_t* __restrict__ inbounds) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const TriIdx t = tri_index(ns, i, d);
        float out[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            auto cell = [&](int64_t ix, int64_t iy, int64_t iz) {
                return __ldg(grid + ix * s.s[0] + iy * s.s[1] + iz * s.s[2] + ch * s.s[3]);
            };
            const float x_cc = lerp_hl(cell(t.hi[0], t.hi[1], t.hi[2]), cell(t.lo[0], t.hi[1], t.hi[2]), t.f[0]);
            const float x_cf = lerp_hl(cell(t.hi[0], t.hi[1], t.lo[2]), cell(t.lo[0], t.hi[1], t.lo[2]), t.f[0]);
            const float x_fc = lerp_hl(cell(t.hi[0], t.lo[1], t.hi[2]), cell(t.lo[0], t.lo[1], t.hi[2]), t.f[0]);
            const float x_ff = lerp_hl(cell(t.hi[0], t.lo[1], t.lo[2]), cell(t.lo[0], t.lo[1], t.lo[2]), t.f[0]);
            const float y_c = lerp_hl(x_cc, x_fc, t.f[1]);
            const float y_f = lerp_hl(x_cf, x_ff, t.f[1]);
            out[ch] = lerp_hl(y_c, y_f, t.f[2]);
            if (masked && !t.inb) out[ch] = __fmul_rn(out[ch], 0.f);
        }
        reinterpret_cast<float4*>(vals)[i] = make_float4(out[0], out[1], out[2], out[3]);
        if (inbounds) inbounds[i] = t.inb ? 1 : 0;
    }
}

cudaError_t launch_trilinear_fwd(const float* ns, int64_t m, const float* grid, const int32_t* dims,
                                 const int64_t* strides, int masked, float* vals, uint8_t* inb, cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    Dims d{{dims[0], dims[1], dims[2]}};
    Strides s{{strides[0], strides[1], strides[2], strides[3]}};
    k_trilinear_fwd<<<blocks_for(m, 256), 256, 0, st>>>(ns, m, grid, d, s, masked, vals, inb);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_trilinear_bwd(const float* __restrict__ ns, int64_t m, const float* __restrict__ gv,
                                                       Dims d, int masked, float* __restrict__ gg) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const TriIdx t = tri_index(ns, i, d);
        if (masked && !t.inb) continue;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gv) + i);
        if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
#pragma unroll
        for (int corner = 0; corner < 8; ++corner) {
            const bool cx = corner & 4, cy = corner & 2, cz = corner & 1;     // 1 = floor side
            const float w = (cx ? 1.f - t.f[0] : t.f[0]) * (cy ? 1.f - t.f[1] : t.f[1]) * (cz ? 1.f - t.f[2] : t.f[2]);
            if (w == 0.f) continue;
            const int64_t ix = cx ? t.lo[0] : t.hi[0], iy = cy ? t.lo[1] : t.hi[1], iz = cz ? t.lo[2] : t.hi[2];
            red_add_v4(gg + ((ix * d.n[1] + iy) * d.n[2] + iz) * 4, g.x * w, g.y * w, g.z * w, g.w * w);
        }
    }
}

cudaError_t launch_trilinear_bwd(const float* ns, int64_t m, const float* grad_vals, const int32_t* dims, int masked,
                                 float* grad_grid, cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    Dims d{{dims[0], dims[1], dims[2]}};
    k_trilinear_bwd<<<blocks_for(m, 256), 256, 0, st>>>(ns, m, grad_vals, d, masked, grad_grid);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// compute_alpha_weighted_pixels — src/ray_sampling.py:172-192, one warp per ray, coalesced 16-byte sample loads
// ---------------------------------------------------------------------------------------------------------------
constexpr int CWARPS = 8;

__global__ void __launch_bounds__(CWARPS * 32) k_composite_fwd(const float4* __restrict__ samples, int64_t n_rays, int S,
                                                               float4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * CWARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const float4* s = samples + ray * S;
    float T = 1.f, ar = 0.f, ag = 0.f, ab = 0.f, aa = 0.f;
    for (int kb = 0; kb < S; kb += 32) {
        const int k = kb + lane;
        const float4 c = k < S ? __ldg(s + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        float total;
        const float ex = warp_excl_prod(1.f - c.w, lane, total);
        const float w = c.w * (T * ex);
        ar = fmaf(w, c.x, ar); ag = fmaf(w, c.y, ag); ab = fmaf(w, c.z, ab); aa += w;
        T *= total;
    }
    ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); aa = warp_sum(aa);
    if (lane == 0) out[ray] = make_float4(ar, ag, ab, aa);
}

cudaError_t launch_composite_fwd(const float* samples, int64_t n_rays, int S, float* out, cudaStream_t st) {
    if (n_rays == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n_rays + CWARPS - 1) / CWARPS);
    k_composite_fwd<<<blocks, CWARPS * 32, 0, st>>>((const float4*)samples, n_rays, S, (float4*)out);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(CWARPS * 32) k_composite_bwd(const float4* __restrict__ samples, int64_t n_rays, int S,
                                                               const float4* __restrict__ grad_out,
                                                               float4* __restrict__ grad_samples) {
    extern __shared__ float s_tc[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t ray = (int64_t)blockIdx.x * CWARPS + wib;
    if (ray >= n_rays) return;
    const int nch = (S + 31) / 32;
    float* tc = s_tc + wib * nch;
    const float4* s = samples + ray * S;
    float4* gs = grad_samples + ray * S;
    const float4 g = __ldg(grad_out + ray);
    float T = 1.f;
    for (int c = 0; c < nch; ++c) {
        if (lane == 0) tc[c] = T;
        const int k = c * 32 + lane;
        float f = k < S ? 1.f - __ldg(s + k).w : 1.f;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) f *= __shfl_xor_sync(FULL, f, d);
        T *= f;
    }
    __syncwarp();
    float carry = 0.f;
    for (int c = nch - 1; c >= 0; --c) {
        const int k = c * 32 + lane;
        const float4 v4 = k < S ? __ldg(s + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float alpha = v4.w;
        const float v = fmaf(v4.x, g.x, fmaf(v4.y, g.y, fmaf(v4.z, g.z, g.w)));
        const float behind = warp_behind(alpha * v, 1.f - alpha, lane, carry);
        float total;
        const float Tk = tc[c] * warp_excl_prod(1.f - alpha, lane, total);
        const float wgt = alpha * Tk;
        if (k < S) gs[k] = make_float4(wgt * g.x, wgt * g.y, wgt * g.z, Tk * (v - behind));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// self-test of the hoisted exact arithmetic (plx_device.cuh) against the compiler's IEEE intrinsics
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 16);
}

__global__ void __launch_bounds__(256) k_selftest(float y, uint64_t n, uint64_t seed, unsigned long long* bad) {
    const FastDiv d = make_fastdiv(y);
    unsigned b0 = 0, b1 = 0, b2 = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t h = mix32(seed * 0x100000001B3ull + i), h2 = mix32(~seed + 3 * i);
        float x;
        if (i & 1) {
            x = ((float)(h & 0xffffff) / 16777216.f - 0.5f) * 1024.f;            // grid-coordinate-like magnitudes
        } else {
            const uint32_t e = 127 - 40 + (h2 % 80);                               // exponents 2^-40 .. 2^39
            x = __uint_as_float(((h >> 8) & 0x80000000u) | (e << 23) | (h & 0x7fffff));
        }
        if (__float_as_uint(fdiv_exact(x, d)) != __float_as_uint(__fdiv_rn(x, y))) ++b0;
        const float a = fabsf(x);
        if (__float_as_uint(fsqrt_exact(a)) != __float_as_uint(__fsqrt_rn(a))) ++b1;
        const float den = __uint_as_float((127 - 30 + (h2 >> 8) % 40) << 23 | (h2 & 0x7fffff));   // 2^-30 .. 2^9
        if (__float_as_uint(fdiv_var(x, den)) != __float_as_uint(__fdiv_rn(x, den))) ++b2;
    }
    if (b0) atomicAdd(bad + 0, (unsigned long long)b0);
    if (b1) atomicAdd(bad + 1, (unsigned long long)b1);
    if (b2) atomicAdd(bad + 2, (unsigned long long)b2);
}

cudaError_t launch_selftest(float y, uint64_t n, uint64_t seed, unsigned long long* bad, cudaStream_t st) {
    k_selftest<<<148 * 8, 256, 0, st>>>(y, n, seed, bad);
    return cudaGetLastError();
}

cudaError_t launch_composite_bwd(const float* samples, int64_t n_rays, int S, const float* grad_out, float* grad_samples,
                                 cudaStream_t st) {
    if (n_rays == 0 || S == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n_rays + CWARPS - 1) / CWARPS);
    const size_t smem = (size_t)CWARPS * ((S + 31) / 32) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_composite_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_composite_bwd<<<blocks, CWARPS * 32, smem, st>>>((const float4*)samples, n_rays, S, (const float4*)grad_out,
                                                       (float4*)grad_samples);
    return cudaGetLastError();
}

}  // namespace plx
