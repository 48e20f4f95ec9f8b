// plx_view.cu — GPU version of the reference's point-splat preview, visulize_3d_in_2d_fast (src/visualization.py:157-232),
// the function scripts/compare_inference_to_image.py:58 actually calls.
//
// Reference algorithm (CPU, torch + numpy): every voxel with alpha > 0.1 becomes a point at its cell centre; its direction
// cosines against the camera's X / Y axes give normalised screen coordinates; points outside [0,1)^2 are dropped; the rest
// are sorted by DECREASING distance and assigned to an (xs, ys, 3) image of ones with numpy fancy indexing, so for every
// pixel the LAST assignment — the NEAREST point — wins (painter's order).
// Here: one pass over the grid with a 64-bit atomicMin per point on  (distance bits << 32 | cell index)  — for positive floats
// the bit pattern orders like the value — and one pass over the pixels that resolves the winner's colour.  No sort, no
// point list.  Equal distances: the reference's argsort is not stable, so either point may win there; we take the lower index.
#include "plx_device.cuh"
#include "plx_launch.h"

namespace plx {

struct SplatCam {
    float pos[3], ax[3], ay[3];     // camera position, X axis (pose[:3,0]), Y axis (pose[:3,1])
    float fov, aspect;
};

__global__ void __launch_bounds__(256) k_splat_points(const float4* __restrict__ grid, int X, int Y, int Z, float pd, SplatCam c,
                                                      int xs, int ys, unsigned long long* __restrict__ zbuf) {
    const int64_t n = (int64_t)X * Y * Z;
    const float hx = ceilf(__fdiv_rn((float)X, 2.f)), hy = ceilf(__fdiv_rn((float)Y, 2.f)), hz = ceilf(__fdiv_rn((float)Z, 2.f));
    const float nax = norm3_plain_f(c.ax[0], c.ax[1], c.ax[2]);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        if (!(__ldg(&grid[e].w) > 0.1f)) continue;                                             // :176
        const int iz = (int)(e % Z), iy = (int)((e / Z) % Y), ix = (int)(e / ((int64_t)Z * Y));
        // world position of the cell, :182-184:  ((i - ceil(dim / 2)) + 1) * pd, every op rounded
        const float px = __fmul_rn(__fadd_rn(__fsub_rn((float)ix, hx), 1.f), pd);
        const float py = __fmul_rn(__fadd_rn(__fsub_rn((float)iy, hy), 1.f), pd);
        const float pz = __fmul_rn(__fadd_rn(__fsub_rn((float)iz, hz), 1.f), pd);
        const float rx = __fsub_rn(px, c.pos[0]), ry = __fsub_rn(py, c.pos[1]), rz = __fsub_rn(pz, c.pos[2]);     // :188
        const float dyz = fmaf(rz, c.ax[2], fmaf(ry, c.ax[1], __fmul_rn(rx, c.ax[0])));        // :190 (matmul)
        const float dxz = fmaf(rz, c.ay[2], fmaf(ry, c.ay[1], __fmul_rn(rx, c.ay[0])));        // :191
        const float dist = norm3_plain_f(rx, ry, rz);                                          // :193, :222
        const float nm = __fmul_rn(dist, nax);
        const float ang_x = __fdiv_rn(dyz, nm), ang_y = __fdiv_rn(dxz, nm);                    // :194-195
        const float xn = __fadd_rn(0.5f, __fdiv_rn(ang_y, -c.fov));                            // :197
        const float yn = __fadd_rn(0.5f, __fdiv_rn(ang_x, __fdiv_rn(c.fov, c.aspect)));        // :198
        if (!(xn < 1.f && xn >= 0.f && yn < 1.f && yn >= 0.f)) continue;                       // :202-208 (drops NaN too)
        const int xx = (int)fminf(fmaxf(rintf(__fmul_rn((float)xs, xn)), 0.f), (float)(xs - 1));   // :219
        const int yy = (int)fminf(fmaxf(rintf(__fmul_rn((float)ys, yn)), 0.f), (float)(ys - 1));   // :220
        const unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)(uint32_t)e;
        atomicMin(zbuf + (int64_t)xx * ys + yy, key);
    }
}

__global__ void __launch_bounds__(256) k_splat_resolve(const float4* __restrict__ grid, const unsigned long long* __restrict__ zbuf,
                                                       int64_t n_pix, float* __restrict__ image) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = zbuf[p];
        float r = 1.f, g = 1.f, b = 1.f;                                                       // np.ones, :228
        if (key != ~0ull) {
            const float4 c = __ldg(grid + (uint32_t)(key & 0xffffffffull));
            r = c.x; g = c.y; b = c.z;                                                         // :229
        }
        image[p * 3 + 0] = r; image[p * 3 + 1] = g; image[p * 3 + 2] = b;
    }
}

cudaError_t launch_splat_view(const float* grid, const int32_t* dims, float pd, const float* pose16, float fov, int xs, int ys,
                              unsigned long long* zbuf, float* image, cudaStream_t st) {
    SplatCam c;
    for (int a = 0; a < 3; ++a) { c.pos[a] = pose16[a * 4 + 3]; c.ax[a] = pose16[a * 4 + 0]; c.ay[a] = pose16[a * 4 + 1]; }
    c.fov = fov;
    auto nrm = [](const float* v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
    c.aspect = nrm(c.ax) / nrm(c.ay);                                                          // :171
    const int64_t n_pix = (int64_t)xs * ys, n = (int64_t)dims[0] * dims[1] * dims[2];
    cudaError_t e = cudaMemsetAsync(zbuf, 0xff, n_pix * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    auto blocks = [](int64_t k) { int64_t b = (k + 255) / 256; return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b)); };
    k_splat_points<<<blocks(n), 256, 0, st>>>((const float4*)grid, dims[0], dims[1], dims[2], pd, c, xs, ys, zbuf);
    k_splat_resolve<<<blocks(n_pix), 256, 0, st>>>((const float4*)grid, zbuf, n_pix, image);
    return cudaGetLastError();
}

}  // namespace plx
