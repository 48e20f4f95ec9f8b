// plx_pool.cu — strided 3-D average pooling of an (X,Y,Z,4) grid and its backward, as three separable box passes.
//
// Stands under average_pool3d_grid (src/grid_functions.py:173-181 = F.avg_pool3d(kernel k, stride s, no padding)) that
// fit() applies to the whole grid on each of its first 230 steps (scripts/train.py:110-118, windows 93^3 .. 3^3, stride
// max(1, k // 4)).  A k^3 window is the product of three 1-D box sums, so the forward is  Z-pass -> Y-pass -> X-pass with
// k additions per output each (instead of k^3), and the backward is the three transposed passes in gather form (every
// input cell sums the <= ceil(k/s) windows that cover it): no atomics, every access a 16-byte cell.
#include "plx_device.cuh"
#include "plx_launch.h"

namespace plx {

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// out[a][o][c] = scale * sum_{d<k} in[a][o*s + d][c]   over a tensor viewed as (A, L, C) -> (A, O, C), C in float4 units
__global__ void __launch_bounds__(256) k_box_fwd(const float4* __restrict__ in, float4* __restrict__ out, int64_t A, int L,
                                                 int O, int64_t C, int k, int s, float scale) {
    const int64_t total = A * O * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = e % C, o = (e / C) % O, a = e / (C * O);
        const float4* src = in + (a * L + o * s) * C + c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int d = 0; d < k; ++d) acc = add4(acc, __ldg(src + d * C));
        out[e] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
    }
}

// transposed pass: gin[a][l][c] = scale * sum_{o : o*s <= l < o*s + k} gout[a][o][c]
__global__ void __launch_bounds__(256) k_box_bwd(const float4* __restrict__ gout, float4* __restrict__ gin, int64_t A, int L,
                                                 int O, int64_t C, int k, int s, float scale) {
    const int64_t total = A * L * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = e % C, l = (e / C) % L, a = e / (C * L);
        int o_lo = (int)l - k + 1;
        o_lo = o_lo <= 0 ? 0 : (o_lo + s - 1) / s;
        int o_hi = (int)(l / s);
        if (o_hi > O - 1) o_hi = O - 1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int o = o_lo; o <= o_hi; ++o) acc = add4(acc, __ldg(gout + (a * O + o) * C + c));
        gin[e] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
    }
}

static unsigned pool_blocks(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = 148 * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_avgpool3d_fwd(const float* in, const int32_t* dims, int k, int s, float* tmp1, float* tmp2, float* out,
                                 cudaStream_t st) {
    const int X = dims[0], Y = dims[1], Z = dims[2];
    const int Ox = (X - k) / s + 1, Oy = (Y - k) / s + 1, Oz = (Z - k) / s + 1;
    const float inv = 1.f / (float)k;
    // Z pass: (X*Y, Z, 1) -> (X*Y, Oz, 1)
    k_box_fwd<<<pool_blocks((int64_t)X * Y * Oz), 256, 0, st>>>((const float4*)in, (float4*)tmp1, (int64_t)X * Y, Z, Oz, 1, k, s, inv);
    // Y pass: (X, Y, Oz) -> (X, Oy, Oz)
    k_box_fwd<<<pool_blocks((int64_t)X * Oy * Oz), 256, 0, st>>>((const float4*)tmp1, (float4*)tmp2, X, Y, Oy, Oz, k, s, inv);
    // X pass: (1, X, Oy*Oz) -> (1, Ox, Oy*Oz)
    k_box_fwd<<<pool_blocks((int64_t)Ox * Oy * Oz), 256, 0, st>>>((const float4*)tmp2, (float4*)out, 1, X, Ox, (int64_t)Oy * Oz, k, s, inv);
    return cudaGetLastError();
}

cudaError_t launch_avgpool3d_bwd(const float* gout, const int32_t* dims, int k, int s, float* tmp2, float* tmp1, float* gin,
                                 cudaStream_t st) {
    const int X = dims[0], Y = dims[1], Z = dims[2];
    const int Ox = (X - k) / s + 1, Oy = (Y - k) / s + 1, Oz = (Z - k) / s + 1;
    const float inv = 1.f / (float)k;
    k_box_bwd<<<pool_blocks((int64_t)X * Oy * Oz), 256, 0, st>>>((const float4*)gout, (float4*)tmp2, 1, X, Ox, (int64_t)Oy * Oz, k, s, inv);
    k_box_bwd<<<pool_blocks((int64_t)X * Y * Oz), 256, 0, st>>>((const float4*)tmp2, (float4*)tmp1, X, Y, Oy, Oz, k, s, inv);
    k_box_bwd<<<pool_blocks((int64_t)X * Y * Z), 256, 0, st>>>((const float4*)tmp1, (float4*)gin, (int64_t)X * Y, Z, Oz, 1, k, s, inv);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// tv_loss — scripts/train.py:44-65: sqrt( sum over the three grid axes of squared neighbour differences, 4 channels )
// and its gradient, as two dense passes: (1) the sum of squares into a double accumulator, (2) the 6-neighbour stencil
// grad += tv / sqrt(S) * (sum_axis (g - g_next) [has next] - (g_prev - g) [has prev]).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sq4(float4 a, float4 b) {
    const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
    return x * x + y * y + z * z + w * w;
}

__global__ void __launch_bounds__(256) k_tv_sumsq(const float4* __restrict__ g, int X, int Y, int Z, double* __restrict__ sum) {
    __shared__ float s_part[8];
    const int64_t n = (int64_t)X * Y * Z;
    float acc = 0.f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(e % Z), y = (int)((e / Z) % Y), x = (int)(e / ((int64_t)Z * Y));
        const float4 c = __ldg(g + e);
        if (z + 1 < Z) acc += sq4(c, __ldg(g + e + 1));
        if (y + 1 < Y) acc += sq4(c, __ldg(g + e + Z));
        if (x + 1 < X) acc += sq4(c, __ldg(g + e + (int64_t)Z * Y));
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += (double)s_part[i];
        atomicAdd(sum, t);
    }
}

__global__ void __launch_bounds__(256) k_tv_grad(const float4* __restrict__ g, int X, int Y, int Z, const double* __restrict__ sum,
                                                 float tv, float4* __restrict__ grad, int64_t cell_begin, int64_t cell_end,
                                                 bool atomic, float* __restrict__ loss_out) {
    const double S = *sum;
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss_out) *loss_out = tv * (float)sqrt(S);
    if (!(S > 0.0) || !grad) return;               // the reference's gradient is 0/0 here; we add nothing
    const float scale = tv / (float)sqrt(S);
    const int64_t sx = (int64_t)Z * Y;
    for (int64_t e = cell_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < cell_end; e += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(e % Z), y = (int)((e / Z) % Y), x = (int)(e / sx);
        const float4 c = __ldg(g + e);
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        auto towards = [&](const float4 nb) { d.x += c.x - nb.x; d.y += c.y - nb.y; d.z += c.z - nb.z; d.w += c.w - nb.w; };
        if (z + 1 < Z) towards(__ldg(g + e + 1));
        if (z > 0)     towards(__ldg(g + e - 1));
        if (y + 1 < Y) towards(__ldg(g + e + Z));
        if (y > 0)     towards(__ldg(g + e - Z));
        if (x + 1 < X) towards(__ldg(g + e + sx));
        if (x > 0)     towards(__ldg(g + e - sx));
        if (atomic) {                            // other GPUs may be reducing into the same cells right now (push exchange)
            red_add_v4(reinterpret_cast<float*>(grad + e), scale * d.x, scale * d.y, scale * d.z, scale * d.w);
        } else {
            float4 o = grad[e];
            o.x = fmaf(scale, d.x, o.x); o.y = fmaf(scale, d.y, o.y); o.z = fmaf(scale, d.z, o.z); o.w = fmaf(scale, d.w, o.w);
            grad[e] = o;
        }
    }
}

cudaError_t launch_tv_loss(const float* grid, const int32_t* dims, float tv, float* grad, int64_t cell_begin, int64_t cell_end,
                           bool atomic, double* scratch, float* loss_out, cudaStream_t st) {
    const int X = dims[0], Y = dims[1], Z = dims[2];
    const int64_t n = (int64_t)X * Y * Z;
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double), st);
    if (e != cudaSuccess) return e;
    k_tv_sumsq<<<pool_blocks(n), 256, 0, st>>>((const float4*)grid, X, Y, Z, scratch);
    k_tv_grad<<<pool_blocks(cell_end - cell_begin), 256, 0, st>>>((const float4*)grid, X, Y, Z, scratch, tv, (float4*)grad, cell_begin, cell_end, atomic, loss_out);
    return cudaGetLastError();
}

}  // namespace plx
