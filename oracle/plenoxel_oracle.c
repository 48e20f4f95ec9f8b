/*
 * plenoxel_oracle.c — plain-C restatement of the voxel-grid volume-rendering hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A second, independent CPU statement of the algorithm DanJbk/Plenoxels runs through torch ATen ops, next to the numpy
 * one (oracle/plenoxel_oracle.py).  It exists to CHECK the CUDA path at sizes the numpy oracle is too slow for; nothing
 * under plenoxels_b200/ may link or load it.  Only tests/, __graft_entry__ (build()/smoke()) and bench.py's CPU legs do.
 *
 * Parity status: pinned.  tests/test_oracle_c.py checks it bit for bit against the numpy oracle (which
 * oracle/validate_against_reference.py pins against /root/reference) and against the golden vectors generated from the
 * reference itself (the .npz fixtures under tests/golden/): linear indices, masks, counts and gathered values equal, pixels equal to the
 * bit, gradients to summation order.
 *
 * Arithmetic contract: every fp32 operation is rounded separately — build with
 * -ffp-contract=off (oracle/c_oracle.py does) — true IEEE division, round half to even, non-negative modulo.
 * Citations are file:line in the reference tree.
 *
 * Build:  gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC oracle/plenoxel_oracle.c -o oracle/_build/libplenoxel_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PLXO_NEAREST 0
#define PLXO_TRILINEAR 1

static inline int64_t pymod(int64_t a, int64_t n) {          /* python-style modulo, src/grid_functions.py:75-77 */
    int64_t r = a % n;
    return r < 0 ? r + n : r;
}

static inline float clamp01(float x, int clamp) {            /* grid.clip(0, 1), scripts/train.py:146 */
    if (!clamp) return x;
    return x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
}

/* normalised sample coordinate of sample k (1-based) on one axis:
 * t = fl(fl32(delta) * fl32(k))                               src/ray_sampling.py:161
 * p = fl(o + fl(d * t))                                        :164-167 (product and sum rounded separately)
 * n = fl(fl(p - gmin) / fl32(pd))                              :13 (true division)                                  */
static inline float norm_coord(float o, float d, float t, float gmin, float pd) {
    const float dt = d * t;
    const float p = o + dt;
    const float q = p - gmin;
    return q / pd;
}

/* float -> int64 as numpy / torch do it for values an int64 cannot hold (NaN, inf, |x| >= 2^63): INT64_MIN, i.e. out of bounds;
 * the plain C cast would be undefined behaviour there */
static inline int64_t to_index(float r) {
    return (r >= -9.0e18f && r <= 9.0e18f) ? (int64_t)r : INT64_MIN;
}

typedef struct {
    const float* grid;
    int64_t nx, ny, nz;
    int clamp;
} Grid;

static inline const float* cell(const Grid* g, int64_t ix, int64_t iy, int64_t iz) {
    return g->grid + ((ix * g->ny + iy) * g->nz + iz) * 4;
}

/* one sample: masked (and clamped) value + linear index (-1 = out of bounds).  Nearest: src/grid_functions.py:103-114
 * (+ :58-61 mask, :75-77 wrap) and the caller's mask multiply scripts/train.py:147; trilinear: :7-44, :220-246 composed
 * as SURVEY.md section 8a row T. */
static inline int64_t lookup(const Grid* g, int mode, float nx_, float ny_, float nz_, float out[4]) {
    if (mode == PLXO_NEAREST) {
        const int64_t ix = to_index(rintf(nx_)), iy = to_index(rintf(ny_)), iz = to_index(rintf(nz_));   /* half to even */
        const int inb = ix >= 0 && ix < g->nx && iy >= 0 && iy < g->ny && iz >= 0 && iz < g->nz;
        if (!inb) { out[0] = out[1] = out[2] = out[3] = 0.f; return -1; }
        const float* c = cell(g, ix, iy, iz);
        for (int ch = 0; ch < 4; ++ch) out[ch] = clamp01(c[ch], g->clamp);
        return (ix * g->ny + iy) * g->nz + iz;
    }
    const int inb = nx_ >= 0.f && nx_ < (float)g->nx && ny_ >= 0.f && ny_ < (float)g->ny && nz_ >= 0.f && nz_ < (float)g->nz;
    if (!inb) { out[0] = out[1] = out[2] = out[3] = 0.f; return -1; }
    const int64_t cx = pymod((int64_t)ceilf(nx_), g->nx), fx_ = pymod((int64_t)floorf(nx_), g->nx);
    const int64_t cy = pymod((int64_t)ceilf(ny_), g->ny), fy_ = pymod((int64_t)floorf(ny_), g->ny);
    const int64_t cz = pymod((int64_t)ceilf(nz_), g->nz), fz_ = pymod((int64_t)floorf(nz_), g->nz);
    const float fx = nx_ - truncf(nx_), fy = ny_ - truncf(ny_), fz = nz_ - truncf(nz_);              /* torch.frac, :29 */
    const float gx = 1.f - fx, gy = 1.f - fy, gz = 1.f - fz;
    for (int ch = 0; ch < 4; ++ch) {
#define V(ix, iy, iz) clamp01(cell(g, ix, iy, iz)[ch], g->clamp)
#define LERP(hi, lo, f, gw) ((hi) * (f) + (lo) * (gw))                /* mul, mul, add: three roundings (-ffp-contract=off) */
        const float x_cc = LERP(V(cx, cy, cz), V(fx_, cy, cz), fx, gx);                               /* :34-35 */
        const float x_cf = LERP(V(cx, cy, fz_), V(fx_, cy, fz_), fx, gx);
        const float x_fc = LERP(V(cx, fy_, cz), V(fx_, fy_, cz), fx, gx);
        const float x_ff = LERP(V(cx, fy_, fz_), V(fx_, fy_, fz_), fx, gx);
        const float y_c = LERP(x_cc, x_fc, fy, gy);                                                   /* :38-39 */
        const float y_f = LERP(x_cf, x_ff, fy, gy);
        out[ch] = LERP(y_c, y_f, fz, gz);                                                             /* :41-42 */
#undef LERP
#undef V
    }
    return 0;
}

/*
 * Forward: rays -> rgba (N,4), depth (N), count (N), optional lin (N,S) int64 — scripts/train.py:130-151.
 * Compositing src/ray_sampling.py:181-191: w = alpha * T; rgb += c * w; A += w; T *= 1 - alpha, all fp32, in order.
 * depth = sum_k w_k t_k (not in the reference).
 */
int plxo_render_forward(const float* grid, const int64_t dims[3], const float* origins, const float* dirs, int64_t n_rays,
                        int32_t num_samples, float delta, const float gmin[3], float pd, int32_t mode, int32_t clamp,
                        float* rgba, float* depth, int32_t* count, int64_t* lin) {
    const Grid g = {grid, dims[0], dims[1], dims[2], clamp};
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* o = origins + r * 3;
        const float* d = dirs + r * 3;
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, dep = 0.f, T = 1.f;
        int32_t cnt = 0;
        for (int32_t k = 1; k <= num_samples; ++k) {
            const float t = delta * (float)k;
            float c[4];
            const int64_t li = lookup(&g, mode, norm_coord(o[0], d[0], t, gmin[0], pd), norm_coord(o[1], d[1], t, gmin[1], pd),
                                      norm_coord(o[2], d[2], t, gmin[2], pd), c);
            if (lin) lin[r * num_samples + (k - 1)] = li;
            cnt += li >= 0;
            const float w = c[3] * T;
            for (int ch = 0; ch < 3; ++ch) { const float cw = c[ch] * w; acc[ch] = acc[ch] + cw; }
            acc[3] = acc[3] + w;
            const float wt = w * t;
            dep = dep + wt;
            const float om = 1.f - c[3];
            T = T * om;
        }
        for (int ch = 0; ch < 4; ++ch) rgba[r * 4 + ch] = acc[ch];
        if (depth) depth[r] = dep;
        if (count) count[r] = cnt;
    }
    return 0;
}

/*
 * Backward: gradient of sum(rgba * grad_rgba) [+ beta term] w.r.t. the RAW grid, accumulated in double into grad_out
 * (X*Y*Z*4, zeroed here).  Autograd of scripts/train.py:146-181 restated:
 *   v_k = c_k . g_rgb + g_A;  d c_k = alpha_k T_k g_rgb;  d alpha_k = T_k (v_k - S_k) [+ beta/M (1/(a+eps) + 1/(1-a+eps))]
 *   S_k = alpha_{k+1} v_{k+1} + (1 - alpha_{k+1}) S_{k+1},  S_last = 0
 * scattered with += at the nearest cell (or the 8 trilinear corners with their weights) and gated per channel by the clip
 * pass-mask 0 <= raw <= 1 (inclusive).  beta_over_m = beta / (number of samples of the whole batch).
 */
int plxo_render_backward(const float* grid, const int64_t dims[3], const float* origins, const float* dirs, int64_t n_rays,
                         int32_t num_samples, float delta, const float gmin[3], float pd, const double* grad_rgba,
                         int32_t mode, int32_t clamp, double beta_over_m, double* grad_out) {
    const Grid g = {grid, dims[0], dims[1], dims[2], clamp};
    const int64_t ncell = dims[0] * dims[1] * dims[2];
    memset(grad_out, 0, sizeof(double) * (size_t)ncell * 4);
    int failed = 0;
#pragma omp parallel
    {
        /* per-thread scratch for one ray; cells are shared between rays, so the scatter uses atomic adds (double: the
         * order of the additions changes the sum by ~1e-16 relative, far below every tolerance it is compared at) */
        double* vals = (double*)malloc(sizeof(double) * (size_t)(num_samples > 0 ? num_samples : 1) * 4);
        double* Tk = (double*)malloc(sizeof(double) * (size_t)(num_samples > 0 ? num_samples : 1));
        float* nsx = (float*)malloc(sizeof(float) * (size_t)(num_samples > 0 ? num_samples : 1) * 3);
        int64_t* li = (int64_t*)malloc(sizeof(int64_t) * (size_t)(num_samples > 0 ? num_samples : 1));
        const int ok = vals && Tk && nsx && li;
        if (!ok) {
#pragma omp atomic write
            failed = 1;
        }
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_rays; ++r) {
            if (!ok) continue;
            const float* o = origins + r * 3;
            const float* d = dirs + r * 3;
            const double* gr = grad_rgba + r * 4;
            double T = 1.0;
            for (int32_t k = 0; k < num_samples; ++k) {
                const float t = delta * (float)(k + 1);
                float c[4];
                nsx[k * 3 + 0] = norm_coord(o[0], d[0], t, gmin[0], pd);
                nsx[k * 3 + 1] = norm_coord(o[1], d[1], t, gmin[1], pd);
                nsx[k * 3 + 2] = norm_coord(o[2], d[2], t, gmin[2], pd);
                li[k] = lookup(&g, mode, nsx[k * 3], nsx[k * 3 + 1], nsx[k * 3 + 2], c);
                for (int ch = 0; ch < 4; ++ch) vals[k * 4 + ch] = (double)c[ch];
                Tk[k] = T;
                T *= 1.0 - (double)c[3];
            }
            double behind = 0.0;
            for (int32_t k = num_samples - 1; k >= 0; --k) {
                const double a = vals[k * 4 + 3];
                const double v = vals[k * 4] * gr[0] + vals[k * 4 + 1] * gr[1] + vals[k * 4 + 2] * gr[2] + gr[3];
                double dv[4] = {a * Tk[k] * gr[0], a * Tk[k] * gr[1], a * Tk[k] * gr[2], Tk[k] * (v - behind)};
                behind = a * v + (1.0 - a) * behind;
                if (li[k] < 0) continue;                              /* masked sample: no gradient path */
                if (beta_over_m != 0.0) dv[3] += beta_over_m * (1.0 / (a + 1e-4) + 1.0 / (1.0 - a + 1e-4));   /* scripts/train.py:170-177 */
                if (mode == PLXO_NEAREST) {
                    const float* raw = grid + li[k] * 4;
                    for (int ch = 0; ch < 4; ++ch)
                        if ((!clamp || (raw[ch] >= 0.f && raw[ch] <= 1.f)) && dv[ch] != 0.0) {
#pragma omp atomic
                            grad_out[li[k] * 4 + ch] += dv[ch];
                        }
                } else {
                    const double x = nsx[k * 3], y = nsx[k * 3 + 1], z = nsx[k * 3 + 2];
                    const double fx = x - trunc(x), fy = y - trunc(y), fz = z - trunc(z);
                    const int64_t ix[2] = {pymod((int64_t)ceil(x), g.nx), pymod((int64_t)floor(x), g.nx)};
                    const int64_t iy[2] = {pymod((int64_t)ceil(y), g.ny), pymod((int64_t)floor(y), g.ny)};
                    const int64_t iz[2] = {pymod((int64_t)ceil(z), g.nz), pymod((int64_t)floor(z), g.nz)};
                    const double wx[2] = {fx, 1.0 - fx}, wy[2] = {fy, 1.0 - fy}, wz[2] = {fz, 1.0 - fz};
                    for (int a_ = 0; a_ < 2; ++a_)
                        for (int b_ = 0; b_ < 2; ++b_)
                            for (int c_ = 0; c_ < 2; ++c_) {
                                const int64_t l = (ix[a_] * g.ny + iy[b_]) * g.nz + iz[c_];
                                const double w = wx[a_] * wy[b_] * wz[c_];
                                const float* raw = grid + l * 4;
                                for (int ch = 0; ch < 4; ++ch)
                                    if ((!clamp || (raw[ch] >= 0.f && raw[ch] <= 1.f)) && dv[ch] * w != 0.0) {
#pragma omp atomic
                                        grad_out[l * 4 + ch] += dv[ch] * w;
                                    }
                            }
                }
            }
        }
        free(vals); free(Tk); free(nsx); free(li);
    }
    return failed ? -1 : 0;
}

/* mean over N*4 elements incl. alpha (scripts/train.py:156); d loss / d pixels into grad (N,4) double. */
double plxo_mse_loss(const float* pixels, const float* targets, int64_t n_rays, int64_t n_global, double* grad) {
    const double denom = 4.0 * (double)(n_global > 0 ? n_global : n_rays);
    double sum = 0.0;
    for (int64_t i = 0; i < n_rays * 4; ++i) {
        const double diff = (double)pixels[i] - (double)targets[i];
        sum += diff * diff;
        if (grad) grad[i] = 2.0 * diff / denom;
    }
    return sum / denom;
}

/*
 * One torch.optim.Adam step (torch/optim/adam.py `_single_tensor_adam`, non-capturable branch, as driven by
 * scripts/train.py:89,:180-184) plus `grid_grad += |grad|`, in place.  fp32 element ops, double scalars; FMA placement as
 * ATen's CPU kernels: m = fma(1-b1, g-m, m); v = fma((1-b2)*g, g, v*b2); p += ((-lr/bc1)*m) / (sqrt(v)/sqrt(bc2) + eps).
 * The two multiply-adds are evaluated as (float)(double product + double addend), exactly like the numpy oracle does.
 */
void plxo_adam_step(float* p, const float* g, float* m, float* v, float* gabs, int64_t n, double lr, int64_t step,
                    double beta1, double beta2, double eps) {
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2);
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const float neg_step = (float)(-(lr / bc1)), bc2_sqrt = (float)pow(bc2, 0.5), epsf = (float)eps;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float diff = g[i] - m[i];
        const float m2 = (float)((double)m[i] + (double)w1 * (double)diff);          /* fma(w, g - m, m) */
        const float vb = v[i] * b2;
        const float vg = w2 * g[i];
        const float v2 = (float)((double)vg * (double)g[i] + (double)vb);            /* fma((1-b2) g, g, v b2) */
        const float sq = sqrtf(v2);
        const float q = sq / bc2_sqrt;
        const float denom = q + epsf;
        const float num = neg_step * m2;
        const float upd = num / denom;
        p[i] = p[i] + upd;
        m[i] = m2;
        v[i] = v2;
        if (gabs) gabs[i] = gabs[i] + fabsf(g[i]);
    }
}

/*
 * Ray directions and target pixels for given uv (C,R,2) — generate_rays_batched, src/ray_sampling.py:212-264.
 * Quirks kept: u is scaled by H, v by W, lookup imgs[cam, v_pix, u_pix]; angles linear in u, v; v negated.
 * The norm of the direction is x0^2 then two multiply-adds rounded once each (ATen's contiguous-axis kernel), the pose-column
 * norms are plain mul/add (strided slices) — as pinned by the numpy oracle against torch-CPU.
 */
static inline float fused_sq_add(float a, float acc) { return (float)((double)a * (double)a + (double)acc); }

int plxo_generate_rays(const float* imgs, int32_t n_cams, int32_t H, int32_t W, const float* poses, float fov, const float* uv,
                       int32_t R, float* dirs, float* targets, int64_t* pix) {
    for (int32_t c = 0; c < n_cams; ++c) {
        const float* P = poses + (int64_t)c * 16;
        const float X[3] = {P[0], P[4], P[8]}, Y[3] = {P[1], P[5], P[9]}, Zn[3] = {-P[2], -P[6], -P[10]};
        float nx = X[0] * X[0]; nx = nx + X[1] * X[1]; nx = nx + X[2] * X[2];         /* plain mul / add, in order */
        float ny = Y[0] * Y[0]; ny = ny + Y[1] * Y[1]; ny = ny + Y[2] * Y[2];
        const float sx = sqrtf(nx), sy = sqrtf(ny);
        const float aspect = sx / sy;                                                       /* :218 */
        const float inv_aspect = 1.f / aspect;
        const float v_scale = fov * inv_aspect;                                             /* :235 */
        for (int32_t r = 0; r < R; ++r) {
            const int64_t ray = (int64_t)c * R + r;
            const float u = uv[ray * 2], v = uv[ray * 2 + 1];
            const float uh = u - 0.5f, vh = v - 0.5f;
            const float u_ang = fov * uh;                                                   /* :234 */
            const float v_pos = v_scale * vh;
            const float v_ang = -v_pos;
            const float hu = (float)H * u, wv = (float)W * v;
            const int64_t up = (int64_t)fminf(rintf(hu), (float)(H - 1));                      /* :238 */
            const int64_t vp = (int64_t)fminf(rintf(wv), (float)(W - 1));                      /* :239 */
            if (pix) { pix[ray * 2] = up; pix[ray * 2 + 1] = vp; }
            if (targets && imgs)
                memcpy(targets + ray * 4, imgs + (((int64_t)c * H + vp) * W + up) * 4, 4 * sizeof(float));   /* :248 */
            float d[3];
            for (int a = 0; a < 3; ++a) {
                const float ux = u_ang * X[a];                                              /* :261 */
                const float vy = v_ang * Y[a];
                const float s = ux + vy;
                d[a] = s + Zn[a];
            }
            float acc = d[0] * d[0];
            acc = fused_sq_add(d[1], acc);
            acc = fused_sq_add(d[2], acc);
            const float nrm = sqrtf(acc);
            for (int a = 0; a < 3; ++a) dirs[ray * 3 + a] = d[a] / nrm;                        /* :262 */
        }
    }
    return 0;
}

/*
 * `torch.linspace(0, 1, n)` in fp32 (src/ray_sampling.py:222) and the even-spread uv lattice built from it (:220-223):
 * lower half 0 + step * i (mul, add), upper half 1 - step * (n - 1 - i) contracted to one multiply-add (ATen's CPU kernel,
 * pinned by the numpy oracle); uv (C, n*n, 2) u-major, n = round(sqrt(R)).
 */
void plxo_linspace01(int32_t n, float* out) {
    if (n == 1) { out[0] = 0.f; return; }
    const float step = 1.f / (float)(n - 1);
    for (int32_t i = 0; i < n; ++i)
        out[i] = i < n / 2 ? 0.f + step * (float)i : (float)(1.0 - (double)step * (double)(n - 1 - i));
}

int plxo_even_spread_uv(int32_t n_cams, int32_t n_side, float* uv) {
    float* line = (float*)malloc(sizeof(float) * (size_t)(n_side > 0 ? n_side : 1));
    if (!line) return -1;
    plxo_linspace01(n_side, line);
    for (int32_t c = 0; c < n_cams; ++c)
        for (int32_t iu = 0; iu < n_side; ++iu)
            for (int32_t iv = 0; iv < n_side; ++iv) {
                float* o = uv + (((int64_t)c * n_side + iu) * n_side + iv) * 2;
                o[0] = line[iu];
                o[1] = line[iv];
            }
    free(line);
    return 0;
}

/*
 * `tv_loss` of scripts/train.py:44-65 in double: sqrt of the summed squared neighbour differences along the three grid axes
 * (all four channels) and its gradient w.r.t. the grid (autograd of :63).  An all-equal grid gives loss 0 and gradient 0 (the
 * reference's autograd gives 0/0 = NaN there).
 */
double plxo_tv_loss(const float* grid, const int64_t dims[3], double* grad) {
    const int64_t X = dims[0], Y = dims[1], Z = dims[2];
    const int64_t sx = Y * Z * 4, sy = Z * 4, sz = 4, n = X * sx;
    double s = 0.0;
    for (int64_t x = 0; x < X; ++x)
        for (int64_t y = 0; y < Y; ++y)
            for (int64_t z = 0; z < Z; ++z)
                for (int ch = 0; ch < 4; ++ch) {
                    const int64_t i = x * sx + y * sy + z * sz + ch;
                    const double v = grid[i];
                    if (x + 1 < X) { const double d = v - (double)grid[i + sx]; s += d * d; }
                    if (y + 1 < Y) { const double d = v - (double)grid[i + sy]; s += d * d; }
                    if (z + 1 < Z) { const double d = v - (double)grid[i + sz]; s += d * d; }
                }
    const double loss = sqrt(s);
    if (grad) {
        memset(grad, 0, sizeof(double) * (size_t)n);
        if (loss > 0.0)
            for (int64_t x = 0; x < X; ++x)
                for (int64_t y = 0; y < Y; ++y)
                    for (int64_t z = 0; z < Z; ++z)
                        for (int ch = 0; ch < 4; ++ch) {
                            const int64_t i = x * sx + y * sy + z * sz + ch;
                            const double v = grid[i];
                            if (x + 1 < X) { const double d = (v - (double)grid[i + sx]) / loss; grad[i] += d; grad[i + sx] -= d; }
                            if (y + 1 < Y) { const double d = (v - (double)grid[i + sy]) / loss; grad[i] += d; grad[i + sy] -= d; }
                            if (z + 1 < Z) { const double d = (v - (double)grid[i + sz]) / loss; grad[i] += d; grad[i + sz] -= d; }
                        }
    }
    return loss;
}

int plxo_version(void) { return 1; }
