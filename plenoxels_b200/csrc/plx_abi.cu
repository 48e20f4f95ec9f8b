// plx_abi.cu — the extern "C" boundary of libplenoxel_b200.so (include/plenoxel_abi.h): argument validation, error
// strings, and the single-call training-step driver.  No torch types, no allocation, no synchronisation.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "plx_device.cuh"
#include "plx_launch.h"

namespace plx {
Tuning& tuning() {
    static Tuning t;
    return t;
}
}  // namespace plx

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_result(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return PLX_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

int check_march(const PlxMarch& m, const float* grid) {
    if (!grid) return fail(PLX_E_NULL, "grid is NULL");
    if (m.nx <= 0 || m.ny <= 0 || m.nz <= 0) return fail(PLX_E_SHAPE, "grid dims must be positive (%d,%d,%d)", m.nx, m.ny, m.nz);
    if ((int64_t)m.nx * m.ny * m.nz > 0x7fffffffLL / 4) return fail(PLX_E_SHAPE, "grid too large for 32-bit cell indices");
    if (m.num_samples < 0) return fail(PLX_E_SHAPE, "num_samples < 0");
    if (m.mode != PLX_NEAREST && m.mode != PLX_TRILINEAR) return fail(PLX_E_UNSUPPORTED, "unknown lookup mode %d", m.mode);
    if (!(m.points_distance == m.points_distance) || m.points_distance == 0.f)
        return fail(PLX_E_SHAPE, "points_distance must be a non-zero number");
    return PLX_OK;
}

int check_rays(const PlxRays& r) {
    if (r.n_rays < 0) return fail(PLX_E_SHAPE, "n_rays < 0");
    if (r.n_rays > 0 && (!r.origins || !r.dirs)) return fail(PLX_E_NULL, "ray origins/dirs are NULL");
    if (r.rays_per_origin <= 0) return fail(PLX_E_SHAPE, "rays_per_origin must be >= 1");
    if (r.origin_comp_stride == 0 && r.n_rays > 0) return fail(PLX_E_SHAPE, "origin_comp_stride must be non-zero");
    return PLX_OK;
}

}  // namespace

extern "C" {

int plx_version(void) { return PLX_ABI_VERSION; }

const char* plx_last_error(void) { return g_err; }

int32_t plx_num_chunks(int32_t num_samples) { return plx::num_chunks(num_samples < 0 ? 0 : num_samples); }

int plx_render_fwd(const PlxRenderFwd* a, void* stream) {
    if (!a) return fail(PLX_E_NULL, "args is NULL");
    int rc;
    if ((rc = check_march(a->march, a->grid)) != PLX_OK) return rc;
    if ((rc = check_rays(a->rays)) != PLX_OK) return rc;
    if (a->rays.n_rays > 0 && !a->rgba && !a->image_u8) return fail(PLX_E_NULL, "rgba is NULL");
    if ((uintptr_t)a->rgba % 16) return fail(PLX_E_ALIGN, "rgba must be 16-byte aligned");
    if (a->image_u8) {
        if (a->image_side <= 0 || (int64_t)a->image_side * a->image_side != a->rays.n_rays)
            return fail(PLX_E_SHAPE, "image epilogue: n_rays (%lld) must equal image_side^2 (%d^2)", (long long)a->rays.n_rays, a->image_side);
        if ((uintptr_t)a->image_u8 % 4) return fail(PLX_E_ALIGN, "image_u8 must be 4-byte aligned");
        if (a->targets && !a->rgba) return fail(PLX_E_NULL, "rgba is required with targets");
    }
    if (a->targets) {
        if (!a->grad_rgba) return fail(PLX_E_NULL, "grad_rgba is required with targets");
        if ((uintptr_t)a->targets % 16 || (uintptr_t)a->grad_rgba % 16) return fail(PLX_E_ALIGN, "targets/grad_rgba must be 16-byte aligned");
    }
    return cuda_result(plx::launch_render_fwd(*a, (cudaStream_t)stream), "plx_render_fwd");
}

int plx_render_bwd(const PlxRenderBwd* a, void* stream) {
    if (!a) return fail(PLX_E_NULL, "args is NULL");
    int rc;
    if ((rc = check_march(a->march, a->grid)) != PLX_OK) return rc;
    if ((rc = check_rays(a->rays)) != PLX_OK) return rc;
    if (a->rays.n_rays > 0 && (!a->grad_rgba || !a->grad_grid)) return fail(PLX_E_NULL, "grad_rgba/grad_grid is NULL");
    if ((uintptr_t)a->grad_rgba % 16 || (uintptr_t)a->grad_grid % 16) return fail(PLX_E_ALIGN, "grad_rgba/grad_grid must be 16-byte aligned");
    if (!a->tcarry && (size_t)plx::num_chunks(a->march.num_samples) * 8 * sizeof(float) > 200 * 1024)
        return fail(PLX_E_UNSUPPORTED, "num_samples %d too large for the in-kernel transmittance cache; pass tcarry", a->march.num_samples);
    return cuda_result(plx::launch_render_bwd(*a, (cudaStream_t)stream), "plx_render_bwd");
}

// PlxPeerSync (optional fused cross-GPU ordering): all-zero = none; `has_work` = the launch will actually run a grid
static int check_sync(const PlxPeerSync& s, bool has_work) {
    if (s.wait_epoch <= 0 && s.signal_epoch <= 0) return PLX_OK;
    if (s.world < 1 || s.world > PLX_MAX_PEERS || s.rank < 0 || s.rank >= s.world) return fail(PLX_E_SHAPE, "peer sync: bad world / rank");
    if (s.wait_channel < 0 || s.wait_channel >= PLX_BARRIER_CHANNELS || s.signal_channel < 0 || s.signal_channel >= PLX_BARRIER_CHANNELS)
        return fail(PLX_E_SHAPE, "peer sync: channel out of range");
    for (int r = 0; r < s.world; ++r) if (!s.flags[r]) return fail(PLX_E_NULL, "peer sync: flag array of rank %d is NULL", r);
    if (s.signal_epoch > 0 && !s.block_counter) return fail(PLX_E_NULL, "peer sync: block_counter is NULL");
    if (s.signal_epoch > 0 && !has_work) return fail(PLX_E_SHAPE, "peer sync: a launch without work cannot signal its peers");
    return PLX_OK;
}

// push exchange: `world` peer-mapped gradient buffers and the owner multiplier of plx_slab_partition; contiguous grids only
static int check_peer_grad(const PlxPeerGrad& p, const PlxMarch& m, const float* grid) {
    if (p.world == 0) return PLX_OK;
    if (p.world < 1 || p.world > PLX_MAX_PEERS) return fail(PLX_E_SHAPE, "peer_grad: world must be 1..%d", PLX_MAX_PEERS);
    for (int r = 0; r < p.world; ++r) {
        if (!p.grads[r]) return fail(PLX_E_NULL, "peer_grad: gradient buffer of rank %d is NULL", r);
        if ((uintptr_t)p.grads[r] % 16) return fail(PLX_E_ALIGN, "peer_grad: gradient buffers must be 16-byte aligned");
    }
    const int64_t cells = (int64_t)m.nx * m.ny * m.nz;
    uint32_t mul = 0;
    if (plx_slab_partition(cells, p.world, 0, &mul, nullptr, nullptr) != PLX_OK) return PLX_E_SHAPE;
    if (mul != p.owner_mul) return fail(PLX_E_SHAPE, "peer_grad: owner_mul %u does not match plx_slab_partition (%u)", p.owner_mul, mul);
    if (!(m.sc == 1 && m.sz == 4 && m.sy == 4 * (int64_t)m.nz && m.sx == 4 * (int64_t)m.nz * m.ny) || (uintptr_t)grid % 16)
        return fail(PLX_E_UNSUPPORTED, "peer_grad: the push exchange needs a contiguous, 16-byte aligned (X,Y,Z,4) grid");
    if (m.mode != PLX_NEAREST && m.mode != PLX_TRILINEAR) return fail(PLX_E_UNSUPPORTED, "peer_grad: unknown mode");
    return PLX_OK;
}

int plx_slab_partition(int64_t n_cells, int32_t world, int32_t rank, uint32_t* owner_mul, int64_t* begin_cell, int64_t* end_cell) {
    if (world < 1 || world > PLX_MAX_PEERS || rank < 0 || rank >= world) return fail(PLX_E_SHAPE, "slab partition: bad world / rank");
    if (n_cells <= world || n_cells > 0x7fffffffLL) return fail(PLX_E_SHAPE, "slab partition: need world < n_cells < 2^31 (got %lld)", (long long)n_cells);
    const unsigned __int128 one = (unsigned __int128)1 << 32;
    const uint64_t mul = (uint64_t)((one * (unsigned)world) / (uint64_t)n_cells);        // < 2^32 because n_cells > world
    auto first_of = [&](int r) -> int64_t {                                              // smallest lin with umulhi(lin, mul) >= r
        const unsigned __int128 num = one * (unsigned)r;
        const int64_t b = (int64_t)((num + mul - 1) / mul);
        return b > n_cells ? n_cells : b;
    };
    if (owner_mul) *owner_mul = (uint32_t)mul;
    if (begin_cell) *begin_cell = first_of(rank);
    if (end_cell) *end_cell = rank + 1 == world ? n_cells : first_of(rank + 1);
    return PLX_OK;
}

int plx_render_train(const PlxRenderTrain* a, void* stream) {
    if (!a) return fail(PLX_E_NULL, "args is NULL");
    int rc;
    if ((rc = check_march(a->march, a->grid)) != PLX_OK) return rc;
    if (a->rays.n_rays < 0) return fail(PLX_E_SHAPE, "n_rays < 0");
    if ((rc = check_sync(a->sync, a->rays.n_rays > 0)) != PLX_OK) return rc;
    if (a->rays.n_rays == 0) return PLX_OK;
    if (!a->grad_grid) return fail(PLX_E_NULL, "grad_grid is NULL");
    if ((uintptr_t)a->grad_grid % 16 || (uintptr_t)a->rgba % 16) return fail(PLX_E_ALIGN, "grad_grid/rgba must be 16-byte aligned");
    if (a->gen.uv) {
        if (!a->gen.imgs || !a->gen.poses) return fail(PLX_E_NULL, "gen.imgs/gen.poses is NULL");
        if (a->gen.img_h <= 0 || a->gen.img_h != a->gen.img_w)
            return fail(PLX_E_UNSUPPORTED, "in-kernel ray generation needs square images (src/ray_sampling.py:238-248), got %dx%d", a->gen.img_h, a->gen.img_w);
        if ((int64_t)a->gen.n_cams * a->gen.rays_per_cam != a->rays.n_rays) return fail(PLX_E_SHAPE, "n_rays != n_cams * rays_per_cam");
        if ((uintptr_t)a->gen.imgs % (a->gen.img_format == PLX_IMG_U8 ? 4 : 16)) return fail(PLX_E_ALIGN, "imgs must be 16-byte (fp32) / 4-byte (uint8) aligned");
    } else {
        if ((rc = check_rays(a->rays)) != PLX_OK) return rc;
        if (!a->targets) return fail(PLX_E_NULL, "targets is NULL");
        if ((uintptr_t)a->targets % 16) return fail(PLX_E_ALIGN, "targets must be 16-byte aligned");
    }
    if (a->gen.uv && a->gen.img_format != PLX_IMG_F32 && a->gen.img_format != PLX_IMG_U8) return fail(PLX_E_UNSUPPORTED, "unknown image format %d", a->gen.img_format);
    if (!plx::render_train_supported(*a)) return fail(PLX_E_UNSUPPORTED, "fused training march: num_samples beyond the shared-memory cache or grid beyond 32-bit cell indices");
    if ((rc = check_peer_grad(a->peer_grad, a->march, a->grid)) != PLX_OK) return rc;
    return cuda_result(plx::launch_render_train(*a, (cudaStream_t)stream), "plx_render_train");
}

int plx_adam_table(double lr, double beta1, double beta2, int64_t first_step, int32_t n, float* table_host) {
    if (!table_host) return fail(PLX_E_NULL, "table_host is NULL");
    if (first_step < 1 || n < 0) return fail(PLX_E_SHAPE, "steps count from 1");
    for (int32_t i = 0; i < n; ++i) {        // the same double-precision scalar math as adam_scalars() below
        const double step = (double)(first_step + i);
        table_host[2 * i] = (float)std::sqrt(1.0 - std::pow(beta2, step));
        table_host[2 * i + 1] = (float)(-(lr / (1.0 - std::pow(beta1, step))));
    }
    return PLX_OK;
}

static plx::AdamScalars adam_scalars(double lr, double beta1, double beta2, double eps, int64_t step) {
    // python-double scalar math of torch/optim/adam.py, cast to fp32 where ATen casts the Scalar
    const double bc1 = 1.0 - std::pow(beta1, (double)step);
    const double bc2 = 1.0 - std::pow(beta2, (double)step);
    plx::AdamScalars s;
    s.one_minus_beta1 = (float)(1.0 - beta1);
    s.beta2 = (float)beta2;
    s.one_minus_beta2 = (float)(1.0 - beta2);
    s.bc2_sqrt = (float)std::sqrt(bc2);
    s.eps = (float)eps;
    s.neg_step_size = (float)(-(lr / bc1));
    s.keep_p = s.keep_g = false;
    s.reverse = (step & 1) != 0;
    s.skip_same = true;
    return s;
}

int plx_adam_step(float* p, float* g, float* m, float* v, float* gabs, int64_t n, double lr, double beta1, double beta2,
                  double eps, int64_t step, int32_t zero_grad, void* stream) {
    if (n < 0) return fail(PLX_E_SHAPE, "n < 0");
    if (n > 0 && (!p || !g || !m || !v)) return fail(PLX_E_NULL, "p/g/m/v is NULL");
    if (step < 1) return fail(PLX_E_SHAPE, "step counts from 1");
    return cuda_result(plx::launch_adam(p, g, m, v, gabs, n, adam_scalars(lr, beta1, beta2, eps, step), zero_grad != 0,
                                        plx::StepTail{nullptr, nullptr, nullptr, 0}, (cudaStream_t)stream), "plx_adam_step");
}

int plx_adam_step_peer(const PlxAdamPeer* a, void* stream) {
    if (!a) return fail(PLX_E_NULL, "args is NULL");
    if (a->world < 1 || a->world > PLX_MAX_PEERS || a->rank < 0 || a->rank >= a->world)
        return fail(PLX_E_SHAPE, "world must be 1..%d and 0 <= rank < world (got world %d, rank %d)", PLX_MAX_PEERS, a->world, a->rank);
    if (a->begin < 0 || a->end < a->begin || a->begin % 4 || a->end % 4) return fail(PLX_E_SHAPE, "owned range must be multiples of 4 floats");
    if (a->step < 1) return fail(PLX_E_SHAPE, "step counts from 1");
    if (!a->exp_avg || !a->exp_avg_sq) return fail(PLX_E_NULL, "exp_avg/exp_avg_sq is NULL");
    for (int r = 0; r < a->world; ++r) {
        if (!a->grids[r] || !a->grads[r]) return fail(PLX_E_NULL, "grid/grad pointer of rank %d is NULL", r);
        if ((uintptr_t)a->grids[r] % 16 || (uintptr_t)a->grads[r] % 16) return fail(PLX_E_ALIGN, "peer buffers must be 16-byte aligned");
    }
    if ((uintptr_t)a->exp_avg % 16 || (uintptr_t)a->exp_avg_sq % 16 || (uintptr_t)a->grad_abs_sum % 16)
        return fail(PLX_E_ALIGN, "optimizer state must be 16-byte aligned");
    int rc;
    if ((rc = check_sync(a->sync, a->end > a->begin)) != PLX_OK) return rc;
    plx::StepTail tail{a->loss_src, a->loss_clear, (float*)a->result_host, (int32_t)a->step};
    if (a->loss_peers[0]) {
        for (int r = 0; r < a->world; ++r) {
            if (!a->loss_peers[r]) return fail(PLX_E_NULL, "loss accumulator of rank %d is NULL", r);
            tail.src_peers[r] = a->loss_peers[r];
        }
        tail.n_peers = a->world;
        tail.global_out = a->loss_out;
    }
    return cuda_result(plx::launch_adam_peer(*a, adam_scalars(a->lr, a->beta1, a->beta2, a->eps, a->step), tail, (cudaStream_t)stream),
                       "plx_adam_step_peer");
}

int plx_adam_step_slab(const PlxAdamSlab* a, void* stream) {
    if (!a) return fail(PLX_E_NULL, "args is NULL");
    if (a->world < 1 || a->world > PLX_MAX_PEERS || a->rank < 0 || a->rank >= a->world)
        return fail(PLX_E_SHAPE, "world must be 1..%d and 0 <= rank < world (got world %d, rank %d)", PLX_MAX_PEERS, a->world, a->rank);
    if (a->begin < 0 || a->end < a->begin || a->begin % 4 || a->end % 4) return fail(PLX_E_SHAPE, "owned range must be multiples of 4 floats");
    if (a->step < 1) return fail(PLX_E_SHAPE, "step counts from 1");
    if (!a->exp_avg || !a->exp_avg_sq || !a->grad) return fail(PLX_E_NULL, "grad/exp_avg/exp_avg_sq is NULL");
    for (int r = 0; r < a->world; ++r) {
        if (!a->grids[r]) return fail(PLX_E_NULL, "grid pointer of rank %d is NULL", r);
        if ((uintptr_t)a->grids[r] % 16) return fail(PLX_E_ALIGN, "peer buffers must be 16-byte aligned");
    }
    if ((uintptr_t)a->exp_avg % 16 || (uintptr_t)a->exp_avg_sq % 16 || (uintptr_t)a->grad_abs_sum % 16 || (uintptr_t)a->grad % 16 ||
        (uintptr_t)a->grid_mc % 16)
        return fail(PLX_E_ALIGN, "optimizer state must be 16-byte aligned");
    if ((uintptr_t)a->result_host % 8) return fail(PLX_E_ALIGN, "result_host must be 8-byte aligned");
    if (a->loss_peers[0])
        for (int r = 0; r < a->world; ++r) if (!a->loss_peers[r]) return fail(PLX_E_NULL, "loss accumulator of rank %d is NULL", r);
    return cuda_result(plx::launch_adam_slab(*a, adam_scalars(a->lr, a->beta1, a->beta2, a->eps, a->step), (cudaStream_t)stream),
                       "plx_adam_step_slab");
}

int plx_peer_barrier(int32_t* const* flags, int32_t rank, int32_t world, int32_t channel, int32_t epoch, const PlxPeerError* err,
                     void* stream) {
    if (!flags) return fail(PLX_E_NULL, "flags is NULL");
    if (world < 1 || world > PLX_MAX_PEERS || rank < 0 || rank >= world) return fail(PLX_E_SHAPE, "bad world / rank");
    if (channel < 0 || channel >= PLX_BARRIER_CHANNELS) return fail(PLX_E_SHAPE, "channel out of range");
    for (int r = 0; r < world; ++r) if (!flags[r]) return fail(PLX_E_NULL, "flag array of rank %d is NULL", r);
    const PlxPeerError none{nullptr, nullptr, 0};
    return cuda_result(plx::launch_peer_barrier(flags, rank, world, channel, epoch, err ? *err : none, (cudaStream_t)stream), "plx_peer_barrier");
}

int plx_generate_rays(const float* imgs, int32_t n_cams, int32_t img_h, int32_t img_w, const float* poses, float fov,
                      const float* uv, int32_t rays_per_cam, int32_t n_side, float* dirs, float* targets, void* stream) {
    if (n_cams < 0 || rays_per_cam < 0) return fail(PLX_E_SHAPE, "negative camera / ray count");
    if (n_cams == 0 || rays_per_cam == 0) return PLX_OK;
    if (!poses || !dirs) return fail(PLX_E_NULL, "poses/dirs is NULL");
    if (!uv && (n_side <= 0 || n_side * n_side != rays_per_cam))
        return fail(PLX_E_SHAPE, "even-spread lattice needs rays_per_cam == n_side^2 (got %d, n_side %d)", rays_per_cam, n_side);
    if (targets) {
        if (!imgs) return fail(PLX_E_NULL, "imgs is NULL but targets requested");
        if (img_h <= 0 || img_w <= 0) return fail(PLX_E_SHAPE, "bad image size");
        // the reference scales u by shape[1] and v by shape[2] but indexes imgs[cam, v_pix, u_pix] (src/ray_sampling.py:238-248):
        // only square images keep both indices in range for every uv
        if (img_h != img_w) return fail(PLX_E_UNSUPPORTED, "non-square images index out of range in the reference (H=%d, W=%d)", img_h, img_w);
        if ((uintptr_t)imgs % 16 || (uintptr_t)targets % 16) return fail(PLX_E_ALIGN, "imgs/targets must be 16-byte aligned");
    }
    PlxRayGen gen;
    gen.imgs = imgs; gen.n_cams = n_cams; gen.img_h = img_h; gen.img_w = img_w; gen.poses = poses; gen.fov = fov; gen.uv = uv;
    gen.rays_per_cam = rays_per_cam; gen.img_format = PLX_IMG_F32;
    return cuda_result(plx::launch_generate_rays(gen, n_side, dirs, targets, (cudaStream_t)stream), "plx_generate_rays");
}

int plx_generate_rays_gen(const PlxRayGen* gen, int32_t n_side, float* dirs, float* targets, void* stream) {
    if (!gen) return fail(PLX_E_NULL, "gen is NULL");
    if (gen->img_format == PLX_IMG_F32)
        return plx_generate_rays((const float*)gen->imgs, gen->n_cams, gen->img_h, gen->img_w, gen->poses, gen->fov, gen->uv, gen->rays_per_cam,
                                 n_side, dirs, targets, stream);
    if (gen->img_format != PLX_IMG_U8) return fail(PLX_E_UNSUPPORTED, "unknown image format %d", gen->img_format);
    if (gen->n_cams < 0 || gen->rays_per_cam < 0) return fail(PLX_E_SHAPE, "negative camera / ray count");
    if (gen->n_cams == 0 || gen->rays_per_cam == 0) return PLX_OK;
    if (!gen->poses || !dirs) return fail(PLX_E_NULL, "poses/dirs is NULL");
    if (!gen->uv && (n_side <= 0 || n_side * n_side != gen->rays_per_cam)) return fail(PLX_E_SHAPE, "even-spread lattice needs rays_per_cam == n_side^2");
    if (targets) {
        if (!gen->imgs) return fail(PLX_E_NULL, "imgs is NULL but targets requested");
        if (gen->img_h <= 0 || gen->img_h != gen->img_w) return fail(PLX_E_UNSUPPORTED, "non-square images index out of range in the reference (H=%d, W=%d)", gen->img_h, gen->img_w);
        if ((uintptr_t)gen->imgs % 4 || (uintptr_t)targets % 16) return fail(PLX_E_ALIGN, "imgs (uint8 RGBA) must be 4-byte, targets 16-byte aligned");
    }
    return cuda_result(plx::launch_generate_rays(*gen, n_side, dirs, targets, (cudaStream_t)stream), "plx_generate_rays_gen");
}

int plx_sample_points(const PlxRays* rays, int32_t num_samples, float delta_step, float* samples, void* stream) {
    if (!rays) return fail(PLX_E_NULL, "rays is NULL");
    int rc;
    if ((rc = check_rays(*rays)) != PLX_OK) return rc;
    if (num_samples < 0) return fail(PLX_E_SHAPE, "num_samples < 0");
    if (rays->n_rays * num_samples > 0 && !samples) return fail(PLX_E_NULL, "samples is NULL");
    return cuda_result(plx::launch_sample_points(*rays, num_samples, delta_step, samples, (cudaStream_t)stream), "plx_sample_points");
}

int plx_normalize_points(const float* samples, int64_t m, const float gmin[3], float points_distance, float* out, void* stream) {
    if (m < 0) return fail(PLX_E_SHAPE, "m < 0");
    if (m > 0 && (!samples || !out || !gmin)) return fail(PLX_E_NULL, "samples/out/gmin is NULL");
    return cuda_result(plx::launch_normalize_points(samples, m, gmin[0], gmin[1], gmin[2], points_distance, out,
                                                    (cudaStream_t)stream), "plx_normalize_points");
}

static int check_lookup(const float* ns, int64_t m, const void* grid, const int32_t* dims, const void* out) {
    if (m < 0) return fail(PLX_E_SHAPE, "m < 0");
    if (!dims) return fail(PLX_E_NULL, "dims is NULL");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail(PLX_E_SHAPE, "grid dims must be positive");
    if (m > 0 && (!ns || !grid || !out)) return fail(PLX_E_NULL, "ns/grid/out is NULL");
    if ((uintptr_t)out % 16) return fail(PLX_E_ALIGN, "value buffers must be 16-byte aligned");
    return PLX_OK;
}

int plx_gather_nearest(const float* ns, int64_t m, const float* grid, const int32_t dims[3], const int64_t strides[4],
                       float* vals, uint8_t* inbounds, int64_t* idx_out, void* stream) {
    int rc;
    if ((rc = check_lookup(ns, m, grid, dims, vals)) != PLX_OK) return rc;
    if (!strides) return fail(PLX_E_NULL, "strides is NULL");
    return cuda_result(plx::launch_gather_nearest(ns, m, grid, dims, strides, vals, inbounds, idx_out, (cudaStream_t)stream),
                       "plx_gather_nearest");
}

int plx_gather_nearest_bwd(const float* ns, int64_t m, const float* grad_vals, const int32_t dims[3], float* grad_grid, void* stream) {
    int rc;
    if ((rc = check_lookup(ns, m, grad_grid, dims, grad_vals)) != PLX_OK) return rc;
    if ((uintptr_t)grad_grid % 16) return fail(PLX_E_ALIGN, "grad_grid must be 16-byte aligned");
    return cuda_result(plx::launch_gather_nearest_bwd(ns, m, grad_vals, dims, grad_grid, (cudaStream_t)stream), "plx_gather_nearest_bwd");
}

int plx_trilinear_fwd(const float* ns, int64_t m, const float* grid, const int32_t dims[3], const int64_t strides[4],
                      int32_t masked, float* vals, uint8_t* inbounds, void* stream) {
    int rc;
    if ((rc = check_lookup(ns, m, grid, dims, vals)) != PLX_OK) return rc;
    if (!strides) return fail(PLX_E_NULL, "strides is NULL");
    return cuda_result(plx::launch_trilinear_fwd(ns, m, grid, dims, strides, masked, vals, inbounds, (cudaStream_t)stream),
                       "plx_trilinear_fwd");
}

int plx_trilinear_bwd(const float* ns, int64_t m, const float* grad_vals, const int32_t dims[3], int32_t masked,
                      float* grad_grid, void* stream) {
    int rc;
    if ((rc = check_lookup(ns, m, grad_grid, dims, grad_vals)) != PLX_OK) return rc;
    if ((uintptr_t)grad_grid % 16) return fail(PLX_E_ALIGN, "grad_grid must be 16-byte aligned");
    return cuda_result(plx::launch_trilinear_bwd(ns, m, grad_vals, dims, masked, grad_grid, (cudaStream_t)stream), "plx_trilinear_bwd");
}

int plx_composite_fwd(const float* samples, int64_t n_rays, int32_t num_samples, float* out, void* stream) {
    if (n_rays < 0 || num_samples < 0) return fail(PLX_E_SHAPE, "negative size");
    if (n_rays > 0 && (!out || (num_samples > 0 && !samples))) return fail(PLX_E_NULL, "samples/out is NULL");
    if ((uintptr_t)samples % 16 || (uintptr_t)out % 16) return fail(PLX_E_ALIGN, "samples/out must be 16-byte aligned");
    return cuda_result(plx::launch_composite_fwd(samples, n_rays, num_samples, out, (cudaStream_t)stream), "plx_composite_fwd");
}

int plx_composite_bwd(const float* samples, int64_t n_rays, int32_t num_samples, const float* grad_out, float* grad_samples, void* stream) {
    if (n_rays < 0 || num_samples < 0) return fail(PLX_E_SHAPE, "negative size");
    if (n_rays > 0 && num_samples > 0 && (!samples || !grad_out || !grad_samples)) return fail(PLX_E_NULL, "samples/grad_out/grad_samples is NULL");
    if ((uintptr_t)samples % 16 || (uintptr_t)grad_out % 16 || (uintptr_t)grad_samples % 16)
        return fail(PLX_E_ALIGN, "composite buffers must be 16-byte aligned");
    if ((size_t)((num_samples + 31) / 32) * 8 * sizeof(float) > 200 * 1024) return fail(PLX_E_UNSUPPORTED, "num_samples too large");
    return cuda_result(plx::launch_composite_bwd(samples, n_rays, num_samples, grad_out, grad_samples, (cudaStream_t)stream),
                       "plx_composite_bwd");
}

static int check_pool(const void* a, const int32_t* dims, int32_t k, int32_t s, const void* t1, const void* t2, const void* b) {
    if (!dims) return fail(PLX_E_NULL, "dims is NULL");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail(PLX_E_SHAPE, "grid dims must be positive");
    if (k < 1 || s < 1) return fail(PLX_E_SHAPE, "kernel and stride must be >= 1");
    if (k > dims[0] || k > dims[1] || k > dims[2]) return fail(PLX_E_SHAPE, "pooling window %d larger than the grid (%d,%d,%d)", k, dims[0], dims[1], dims[2]);
    if (!a || !b || !t1 || !t2) return fail(PLX_E_NULL, "pooling buffer is NULL");
    if ((uintptr_t)a % 16 || (uintptr_t)b % 16 || (uintptr_t)t1 % 16 || (uintptr_t)t2 % 16) return fail(PLX_E_ALIGN, "pooling buffers must be 16-byte aligned");
    return PLX_OK;
}

int plx_avgpool3d_fwd(const float* in, const int32_t dims[3], int32_t kernel, int32_t stride, float* tmp1, float* tmp2,
                      float* out, void* stream) {
    int rc;
    if ((rc = check_pool(in, dims, kernel, stride, tmp1, tmp2, out)) != PLX_OK) return rc;
    return cuda_result(plx::launch_avgpool3d_fwd(in, dims, kernel, stride, tmp1, tmp2, out, (cudaStream_t)stream), "plx_avgpool3d_fwd");
}

int plx_avgpool3d_bwd(const float* grad_out, const int32_t dims[3], int32_t kernel, int32_t stride, float* tmp2, float* tmp1,
                      float* grad_in, void* stream) {
    int rc;
    if ((rc = check_pool(grad_out, dims, kernel, stride, tmp1, tmp2, grad_in)) != PLX_OK) return rc;
    return cuda_result(plx::launch_avgpool3d_bwd(grad_out, dims, kernel, stride, tmp2, tmp1, grad_in, (cudaStream_t)stream), "plx_avgpool3d_bwd");
}

int plx_tv_loss_range(const float* grid, const int32_t dims[3], float tv, float* grad, int64_t cell_begin, int64_t cell_end,
                      int32_t atomic, double* scratch, float* loss_out, void* stream) {
    if (!dims || !grid || !scratch) return fail(PLX_E_NULL, "grid/dims/scratch is NULL");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail(PLX_E_SHAPE, "grid dims must be positive");
    if ((uintptr_t)grid % 16 || (uintptr_t)grad % 16 || (uintptr_t)scratch % 8) return fail(PLX_E_ALIGN, "tv buffers are misaligned");
    const int64_t n = (int64_t)dims[0] * dims[1] * dims[2];
    if (cell_begin < 0 || cell_end < cell_begin || cell_end > n) return fail(PLX_E_SHAPE, "cell range outside the grid");
    return cuda_result(plx::launch_tv_loss(grid, dims, tv, grad, cell_begin, cell_end, atomic != 0, scratch, loss_out, (cudaStream_t)stream), "plx_tv_loss");
}

int plx_tv_loss(const float* grid, const int32_t dims[3], float tv, float* grad, double* scratch, float* loss_out, void* stream) {
    if (!dims) return fail(PLX_E_NULL, "grid/dims/scratch is NULL");
    return plx_tv_loss_range(grid, dims, tv, grad, 0, (int64_t)dims[0] * dims[1] * dims[2], 0, scratch, loss_out, stream);
}

int plx_splat_view(const float* grid, const int32_t dims[3], float points_distance, const float* pose_host, float fov,
                   int32_t xs, int32_t ys, uint64_t* zbuf, float* image, void* stream) {
    if (!grid || !dims || !pose_host || !zbuf || !image) return fail(PLX_E_NULL, "splat: grid/dims/pose/zbuf/image is NULL");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail(PLX_E_SHAPE, "grid dims must be positive");
    if ((int64_t)dims[0] * dims[1] * dims[2] > 0xffffffffLL) return fail(PLX_E_SHAPE, "grid too large for 32-bit cell indices");
    if (xs <= 0 || ys <= 0) return fail(PLX_E_SHAPE, "image size must be positive (%d x %d)", xs, ys);
    if ((uintptr_t)grid % 16 || (uintptr_t)zbuf % 8) return fail(PLX_E_ALIGN, "splat buffers are misaligned");
    return cuda_result(plx::launch_splat_view(grid, dims, points_distance, pose_host, fov, xs, ys, (unsigned long long*)zbuf, image,
                                              (cudaStream_t)stream), "plx_splat_view");
}

int plx_tune(const char* name, int32_t value) {
    if (!name) return fail(PLX_E_NULL, "name is NULL");
    plx::Tuning& t = plx::tuning();
    if (!std::strcmp(name, "adam_skip_same")) t.adam_skip_same = value;
    else if (!std::strcmp(name, "adam_blocks_per_sm")) { if (value < 1 || value > 8) return fail(PLX_E_SHAPE, "adam_blocks_per_sm must be 1..8"); t.adam_blocks_per_sm = value; }
    else if (!std::strcmp(name, "train_wpb")) { if (value != 1 && value != 2 && value != 4) return fail(PLX_E_SHAPE, "train_wpb must be 1, 2 or 4"); t.train_wpb = value; }
    else if (!std::strcmp(name, "pdl")) t.pdl = value != 0;
    else if (!std::strcmp(name, "train_cache_it")) { if (value < 0) return fail(PLX_E_SHAPE, "train_cache_it must be >= 0"); t.train_cache_it = value; }
    else if (!std::strcmp(name, "packet_tile")) t.packet_tile = value != 0;
    else return fail(PLX_E_UNSUPPORTED, "unknown tuning switch '%s'", name);
    return PLX_OK;
}

int plx_selftest_arith(float y, uint64_t n, uint64_t seed, uint64_t* mismatches, void* stream) {
    if (!mismatches) return fail(PLX_E_NULL, "mismatches is NULL");
    if (!(y == y) || y == 0.f) return fail(PLX_E_SHAPE, "divisor must be a non-zero number");
    return cuda_result(plx::launch_selftest(y, n, seed, (unsigned long long*)mismatches, (cudaStream_t)stream), "plx_selftest_arith");
}

// shared body of plx_train_step / plx_train_step_host: `uv` is where the march reads this step's uv from (device memory,
// or pinned host memory for the zero-copy end-to-end path); `result_host` receives {loss, step} from the optimiser kernel
static int train_step_impl(const PlxTrainStep* a, const float* uv, void* result_host, int32_t phase, void* stream) {
    if (!a) return fail(PLX_E_NULL, "args is NULL");
    int rc;
    const int64_t n_rays = (int64_t)a->n_cams * a->rays_per_cam;
    if (!a->loss) return fail(PLX_E_NULL, "train step: loss (2 floats) is NULL");
    float* loss_now = a->loss + (a->step & 1);
    float* loss_next = a->loss + ((a->step + 1) & 1);
    if (phase & PLX_STEP_RENDER) {
        if (!uv || !a->dirs || !a->targets || !a->rgba || !a->grad_rgba || !a->grad)
            return fail(PLX_E_NULL, "train step: uv/scratch/grad pointer is NULL");
        if (a->n_rays_global < n_rays) return fail(PLX_E_SHAPE, "n_rays_global < local ray count");
        const float grad_scale = (float)(2.0 / (4.0 * (double)a->n_rays_global));
        const float loss_scale = (float)(1.0 / (4.0 * (double)a->n_rays_global));
        // preferred: one fused kernel (ray generation + forward + loss + backward); PLX_STEP_UNFUSED asks for the 3-kernel path
        const bool allow_fused = !(phase & PLX_STEP_UNFUSED);
        PlxRenderTrain t;
        std::memset(&t, 0, sizeof(t));
        t.march = a->march;
        t.rays.n_rays = n_rays;
        t.gen.imgs = a->imgs; t.gen.n_cams = a->n_cams; t.gen.img_h = a->img_h; t.gen.img_w = a->img_w;
        t.gen.poses = a->poses; t.gen.fov = a->fov; t.gen.uv = uv; t.gen.rays_per_cam = a->rays_per_cam;
        t.gen.img_format = a->img_format;
        if (a->peer_grad) t.peer_grad = *a->peer_grad;
        if (a->replay) {
            if (!a->replay->step_dev || !a->replay->table || !a->replay->block_counter || a->replay->table_len <= 0)
                return fail(PLX_E_NULL, "train step: incomplete replay state");
            t.step_dev = a->replay->step_dev;   // the kernel picks the loss slot from the device-resident step number
        }
        t.grid = a->grid; t.grad_grid = a->grad; t.rgba = a->rgba; t.loss = a->replay ? a->loss : loss_now;
        t.grad_scale = grad_scale; t.loss_scale = loss_scale; t.beta_over_m = a->beta_over_m;
        if (a->render_sync) t.sync = *a->render_sync;
        const bool synced = t.sync.wait_epoch > 0 || t.sync.signal_epoch > 0;
        if ((synced || t.peer_grad.world > 0 || a->replay) && !(allow_fused && a->img_h == a->img_w && plx::render_train_supported(t)))
            return fail(PLX_E_UNSUPPORTED, "train step: render_sync / peer_grad / replay need the fused march (square images, samples within the shared-memory cache)");
        if (allow_fused && a->img_h == a->img_w && plx::render_train_supported(t)) {
            if ((rc = plx_render_train(&t, stream)) != PLX_OK) return rc;
        } else {
            if ((rc = plx_generate_rays_gen(&t.gen, 0, a->dirs, a->targets, stream)) != PLX_OK) return rc;
            PlxRenderFwd f;
            std::memset(&f, 0, sizeof(f));
            f.march = a->march;
            // camera positions = poses[:, :3, 3] (src/ray_sampling.py:159), read in place through the strided view
            f.rays.origins = a->poses + 3;
            f.rays.origin_stride = 16;
            f.rays.origin_comp_stride = 4;
            f.rays.dirs = a->dirs;
            f.rays.n_rays = n_rays;
            f.rays.rays_per_origin = a->rays_per_cam;
            f.grid = a->grid;
            f.rgba = a->rgba; f.tcarry = a->tcarry;
            f.targets = a->targets; f.grad_rgba = a->grad_rgba; f.loss = loss_now;
            f.grad_scale = grad_scale;
            f.loss_scale = loss_scale;
            if ((rc = plx_render_fwd(&f, stream)) != PLX_OK) return rc;
            PlxRenderBwd b;
            std::memset(&b, 0, sizeof(b));
            b.march = a->march; b.rays = f.rays; b.grid = a->grid; b.grad_rgba = a->grad_rgba; b.tcarry = a->tcarry;
            b.grad_grid = a->grad; b.beta_over_m = a->beta_over_m;
            if ((rc = plx_render_bwd(&b, stream)) != PLX_OK) return rc;
        }
    }
    if (phase & PLX_STEP_OPTIM) {
        const int64_t n = (int64_t)a->march.nx * a->march.ny * a->march.nz * 4;
        if (n > 0 && (!a->grid || !a->grad || !a->exp_avg || !a->exp_avg_sq)) return fail(PLX_E_NULL, "train step: optimiser state is NULL");
        if (a->step < 1 && !a->replay) return fail(PLX_E_SHAPE, "step counts from 1");
        const plx::StepTail tail{loss_now, loss_next, (float*)result_host, (int32_t)a->step};
        plx::ReplayArgs rp;
        if (a->replay) {
            if (!a->replay->step_dev || !a->replay->table || !a->replay->block_counter || a->replay->table_len <= 0)
                return fail(PLX_E_NULL, "train step: incomplete replay state");
            rp.step_dev = a->replay->step_dev; rp.table = a->replay->table; rp.table_base = a->replay->table_base;
            rp.table_len = a->replay->table_len; rp.block_counter = a->replay->block_counter; rp.loss2 = a->loss;
        }
        return cuda_result(plx::launch_adam(a->grid, a->grad, a->exp_avg, a->exp_avg_sq, a->grad_abs_sum, n,
                                            adam_scalars(a->lr, a->beta1, a->beta2, a->eps, a->replay ? 1 : a->step), true, tail,
                                            (cudaStream_t)stream, rp), "plx_train_step(optim)");
    }
    return PLX_OK;
}

int plx_train_step(const PlxTrainStep* a, int32_t phase, void* stream) {
    return train_step_impl(a, a ? a->uv : nullptr, nullptr, phase, stream);
}

int plx_train_step_host(const PlxTrainStep* a, const float* uv_host, void* result_host, int32_t phase, void* stream) {
    if ((phase & PLX_STEP_RENDER) && !uv_host) return fail(PLX_E_NULL, "uv_host is NULL");
    if ((uintptr_t)result_host % 8) return fail(PLX_E_ALIGN, "result_host must be 8-byte aligned");
    return train_step_impl(a, (phase & PLX_STEP_RENDER) ? uv_host : (a ? a->uv : nullptr), result_host, phase, stream);
}

}  // extern "C"
