// plx_pool.cu — strided 3-D average pooling of an (X,Y,Z,4) grid and its backward, as three separable box passes.
//
// Stands under average_pool3d_grid (src/grid_functions.py:173-181 = F.avg_pool3d(kernel k, stride s, no padding)) that
// fit() applies to the whole grid on each of its first 230 steps (scripts/train.py:110-118, windows 93^3 .. 3^3, stride
// max(1, k // 4)).  A k^3 window is the product of three 1-D box sums, so the forward is  Z-pass -> Y-pass -> X-pass with
// k additions per output each (instead of k^3), and the backward is the three transposed passes in gather form (every
// input cell sums the <= ceil(k/s) windows that cover it): no atomics, every access a 16-byte cell.  Stride-1 windows (the
// last pooled stage, window 3) take a one-pass sliding-window kernel instead (k_box3_stride1 below).
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

// Stride-1 windows (the last pooled stage of fit(): window 3, stride max(1, 3 // 4) = 1; scripts/train.py:110-118) in ONE pass
// instead of three: out = in (*) box_K^3 with `in` read as zero outside its bounds,
//   out[x][y][z] = sum_{dx,dy,dz < K} in[x - pad + dx][y - pad + dy][z - pad + dz] / K^3.
// pad = 0 is the forward (out = in - (K-1) per axis), pad = K - 1 the backward (out = in + (K-1): every input cell gathers the
// windows that cover it).  One thread owns an (y, z) column of a chunk of x and slides along x keeping the last K plane sums
// (K^2 cells each) in registers: K (K + 1) / 2 loads per output instead of K^3 (two z-neighbours per thread), neighbouring
// threads share them through L1, DRAM sees the input once and the output once (the three separable passes move each three
// times).  256^3, window 3, forward + backward: 0.93 ms -> 0.33 ms (0.15 + 0.18; 497 MB of DRAM traffic each).
template <int K, bool PADDED>
__global__ void __launch_bounds__(256) k_box3_stride1(const float4* __restrict__ in, float4* __restrict__ out, int IX, int IY, int IZ,
                                                      int OX, int OY, int OZ, int xchunk) {
    const int pad = PADDED ? K - 1 : 0;
    const int OZP = (OZ + 1) / 2;            // a thread owns TWO neighbouring z outputs: their windows share K - 1 of K + 1 cells per row
    const int64_t cols = (int64_t)OY * OZP;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = (int)(e / cols);
    const int x0 = chunk * xchunk;
    if (x0 >= OX) return;
    const int x1 = x0 + xchunk < OX ? x0 + xchunk : OX;
    const int oy = (int)((e % cols) / OZP), oz = 2 * (int)(e % OZP);
    const bool second = oz + 1 < OZ;
    const int y_lo = oy - pad, z_lo = oz - pad;                                      // first row / cell under the two windows
    const bool yz_inside = y_lo >= 0 && y_lo + K <= IY && z_lo >= 0 && z_lo + K + 1 <= IZ;     // K rows x (K + 1) cells all exist
    // sums of the K x K cells of input plane ix under the window of output z (s0) and of output z + 1 (s1)
    auto plane = [&](int ix, float4& s0, float4& s1) {
        s0 = make_float4(0.f, 0.f, 0.f, 0.f);
        s1 = s0;
        if (PADDED && (ix < 0 || ix >= IX)) return;
        const float4* base = in + ((int64_t)ix * IY + y_lo) * IZ + z_lo;
        if (yz_inside) {                     // interior column: no per-cell tests
#pragma unroll
            for (int dy = 0; dy < K; ++dy) {
                const float4* row = base + (int64_t)dy * IZ;
                float4 c[K + 1];
#pragma unroll
                for (int dz = 0; dz <= K; ++dz) c[dz] = __ldg(row + dz);
                float4 mid = c[1];
#pragma unroll
                for (int dz = 2; dz < K; ++dz) mid = add4(mid, c[dz]);
                s0 = add4(s0, K > 1 ? add4(c[0], mid) : c[0]);
                s1 = add4(s1, K > 1 ? add4(mid, c[K]) : c[K]);
            }
        } else {
#pragma unroll
            for (int dy = 0; dy < K; ++dy) {
                const int iy = y_lo + dy;
                if (iy < 0 || iy >= IY) continue;
#pragma unroll
                for (int dz = 0; dz <= K; ++dz) {
                    const int iz = z_lo + dz;
                    if (iz < 0 || iz >= IZ) continue;
                    const float4 c = __ldg(base + (int64_t)dy * IZ + dz);
                    if (dz < K) s0 = add4(s0, c);
                    if (dz > 0) s1 = add4(s1, c);
                }
            }
        }
    };
    float4 r0[K], r1[K];
#pragma unroll
    for (int j = 0; j < K - 1; ++j) plane(x0 - pad + j, r0[j], r1[j]);
    const FastDiv dn = make_fastdiv((float)(K * K * K));     // sum / K^3 as the correctly rounded quotient (3 FMAs, plx_device.cuh)
    for (int xb = x0; xb < x1; xb += K) {
#pragma unroll
        for (int j = 0; j < K; ++j) {        // ring slot (K - 1 + j) % K receives plane x + K - 1: compile-time slots, no local memory
            const int x = xb + j;
            if (x < x1) {
                plane(x - pad + K - 1, r0[(K - 1 + j) % K], r1[(K - 1 + j) % K]);
                float4 a0 = r0[j % K], a1 = r1[j % K];
#pragma unroll
                for (int t = 1; t < K; ++t) { a0 = add4(a0, r0[(j + t) % K]); a1 = add4(a1, r1[(j + t) % K]); }
                float4* dst = out + ((int64_t)x * OY + oy) * OZ + oz;
                dst[0] = make_float4(fdiv_hoisted(a0.x, dn), fdiv_hoisted(a0.y, dn), fdiv_hoisted(a0.z, dn), fdiv_hoisted(a0.w, dn));
                if (second) dst[1] = make_float4(fdiv_hoisted(a1.x, dn), fdiv_hoisted(a1.y, dn), fdiv_hoisted(a1.z, dn), fdiv_hoisted(a1.w, dn));
            }
        }
    }
}

template <bool PADDED>
static bool launch_box3_stride1(const float* in, float* out, int IX, int IY, int IZ, int k, cudaStream_t st) {
    const int d = PADDED ? k - 1 : -(k - 1);
    const int OX = IX + d, OY = IY + d, OZ = IZ + d;
    const int64_t cols = (int64_t)OY * ((OZ + 1) / 2);
    // about one wave of 2048 threads per SM in 128-thread blocks (measured on 256^3, window 3: x chunks of 16 .. 64 planes and
    // blocks of 64 .. 256 threads are within 10 % of each other; longer chunks starve the SMs, shorter ones re-read K - 1 planes
    // too often)
    int chunks = (int)((148ll * 2048 + cols - 1) / cols);
    if (chunks < 1) chunks = 1;
    if (chunks > (OX + 7) / 8) chunks = (OX + 7) / 8;
    const int xchunk = (OX + chunks - 1) / chunks;
    chunks = (OX + xchunk - 1) / xchunk;
    const unsigned blocks = (unsigned)((cols * chunks + 127) / 128);
    const float4* i4 = (const float4*)in;
    float4* o4 = (float4*)out;
    switch (k) {
        case 2: k_box3_stride1<2, PADDED><<<blocks, 128, 0, st>>>(i4, o4, IX, IY, IZ, OX, OY, OZ, xchunk); return true;
        case 3: k_box3_stride1<3, PADDED><<<blocks, 128, 0, st>>>(i4, o4, IX, IY, IZ, OX, OY, OZ, xchunk); return true;
        case 4: k_box3_stride1<4, PADDED><<<blocks, 128, 0, st>>>(i4, o4, IX, IY, IZ, OX, OY, OZ, xchunk); return true;
        case 5: k_box3_stride1<5, PADDED><<<blocks, 128, 0, st>>>(i4, o4, IX, IY, IZ, OX, OY, OZ, xchunk); return true;
        default: return false;
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
    if (s == 1 && launch_box3_stride1<false>(in, out, X, Y, Z, k, st)) return cudaGetLastError();
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
    if (s == 1 && launch_box3_stride1<true>(gout, gin, Ox, Oy, Oz, k, st)) return cudaGetLastError();
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
