/*
 * plenoxel_abi.h — C ABI of libplenoxel_b200.so (sm_100a CUDA kernels for the voxel-grid volume renderer).
 *
 * The reference (DanJbk/Plenoxels) has no FFI: its boundary is a set of Python free functions on torch
 * tensors.  Each entry point below names the reference function(s) it stands under
 * (`file:line` in the reference tree); the Python modules under `plenoxels_b200/` bind them with ctypes and keep the
 * reference's Python signatures on top (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`.
 *   - Nothing allocates, synchronises or keeps global state; every call is asynchronous on `stream`
 *     (a `cudaStream_t` passed as `void*`; NULL = the legacy default stream).
 *   - Return value: 0 on success, a negative PLX_E_* code for an argument error, or a positive
 *     `cudaError_t` if a launch failed.  `plx_last_error()` returns a thread-local message.  Nothing throws.
 *   - All floating-point data is fp32; cells are 4 floats (R, G, B, opacity) — src/grid_functions.py:214.
 *   - Index arithmetic is IEEE fp32 with every operation rounded separately (no FMA contraction, true
 *     division, round-half-even), so voxel indices match the reference's CPU path bit for bit
 *     (SURVEY.md §7 H2, Appendix A1-A5).
 */
#ifndef PLENOXEL_ABI_H_
#define PLENOXEL_ABI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLX_ABI_VERSION 2

/* error codes */
#define PLX_OK 0
#define PLX_E_NULL (-1)        /* a required pointer is NULL */
#define PLX_E_SHAPE (-2)       /* a size / stride is invalid */
#define PLX_E_UNSUPPORTED (-3) /* valid request the library does not implement */
#define PLX_E_ALIGN (-4)       /* pointer not aligned for the vector path */
#define PLX_E_PEER_TIMEOUT (-5) /* a cross-GPU wait gave up (a peer died or stalled): the step must not be trusted */

/* lookup mode */
#define PLX_NEAREST 0   /* get_nearest_voxels, src/grid_functions.py:103-114 */
#define PLX_TRILINEAR 1 /* get_grid_points_indices + trilinear_interpolation, src/grid_functions.py:220-246, :7-44 */

/* flags of PlxMarch.flags */
#define PLX_CLAMP01 1u        /* look the grid up through clip(0,1) (scripts/train.py:146); backward applies the pass-mask */
#define PLX_NO_CLIP 2u        /* visit every sample k = 1..S (disable the conservative ray/box pre-filter) */
#define PLX_NO_EARLY_STOP 4u  /* keep marching after the transmittance reached exactly 0 */
#define PLX_COHERENT_RAYS 8u  /* hint: consecutive rays are neighbouring pixels of one view (inference, even-spread lattice);
                                 plx_render_fwd then marches one ray per THREAD, 32 neighbouring rays per warp, so that the
                                 lanes of a load hit neighbouring cells (sector / L1 reuse) and no warp scan is needed.
                                 When rays_per_origin is a square number side^2 the rays of each origin are taken to be the
                                 side x side lattice (ray = iu * side + iv) and a block marches a 16 x 8 tile of it; the
                                 result of a ray never depends on which thread marched it */

/* Ray-marching geometry shared by the fused kernels. */
typedef struct PlxMarch {
    int32_t nx, ny, nz;       /* grid cells per axis (X, Y, Z) */
    int32_t num_samples;      /* S samples per ray, k = 1..S (src/ray_sampling.py:161) */
    int64_t sx, sy, sz, sc;   /* element strides of the (X,Y,Z,4) grid tensor (pooled grids are channel-planar) */
    float gmin[3];            /* world coordinate of cell (0,0,0) = grid_indices.min(0)[0] (src/ray_sampling.py:13) */
    float points_distance;    /* fp32(pd), the divisor of src/ray_sampling.py:13 */
    float delta_step;         /* fp32(delta), t_k = fl(delta * k) */
    int32_t mode;             /* PLX_NEAREST | PLX_TRILINEAR */
    uint32_t flags;           /* PLX_CLAMP01 | PLX_NO_CLIP | PLX_NO_EARLY_STOP */
} PlxMarch;

/* Rays: component a of the origin of ray r is origins[(r / rays_per_origin) * origin_stride + a * origin_comp_stride],
 * i.e. one origin per camera exactly as `camera_positions` is repeated in src/ray_sampling.py:164.  A packed (C,3) tensor
 * has strides (3,1); the view `transform_matrices[:, :3, 3]` of src/ray_sampling.py:159 has (16,4) and is read in place.
 * Pass rays_per_origin = 1 for per-ray origins. */
typedef struct PlxRays {
    const float* origins;
    const float* dirs;        /* (n_rays, 3) contiguous */
    int64_t n_rays;
    int64_t rays_per_origin;
    int64_t origin_stride;
    int64_t origin_comp_stride;
} PlxRays;

int plx_version(void);
const char* plx_last_error(void);

/* Number of 32-sample chunks the fused kernels use per ray; `tcarry` buffers are (n_rays, plx_num_chunks(S)) fp32. */
int32_t plx_num_chunks(int32_t num_samples);

/*
 * K1 — fused forward: sample placement + normalisation + lookup + mask + compositing, one warp per ray.
 * Stands under the sequence scripts/train.py:130-151 / src/visualization.py:125-146:
 *   sample_camera_rays_batched (src/ray_sampling.py:161-167), normalize_samples_for_indecies (:13),
 *   get_nearest_voxels on grid.clip(0,1) (src/grid_functions.py:103-114) * mask, compute_alpha_weighted_pixels (:172-192).
 * Outputs (any of depth/count/sample_index/tcarry may be NULL):
 *   rgba (n_rays,4); depth (n_rays) = sum_k w_k t_k; count (n_rays) int32 = in-bounds samples of the ray;
 *   sample_index (n_rays,S) int32 linear cell index nx-major ((ix*ny+iy)*nz+iz) or -1 when out of bounds
 *   (debug/parity dump; forces PLX_NO_CLIP | PLX_NO_EARLY_STOP); tcarry (n_rays, plx_num_chunks(S)) transmittance at
 *   the start of each 32-sample chunk of the clipped range, consumed by plx_render_bwd.
 * Optional MSE epilogue (scripts/train.py:156): if `targets` != NULL, also writes
 *   grad_rgba = (rgba - targets) * grad_scale   and atomically adds  sum((rgba-targets)^2) * loss_scale  to loss[0].
 */
typedef struct PlxRenderFwd {
    PlxMarch march;
    PlxRays rays;
    const float* grid;
    float* rgba;
    float* depth;
    int32_t* count;
    int32_t* sample_index;
    float* tcarry;
    const float* targets;   /* (n_rays,4) or NULL */
    float* grad_rgba;       /* (n_rays,4), required when targets != NULL */
    float* loss;            /* 1 float accumulator (caller zeroes it), or NULL */
    float grad_scale;       /* 2 / (4 * N_global) for mean-MSE */
    float loss_scale;       /* 1 / (4 * N_global) */
    /* Optional image epilogue of visulize_3d_in_2d (src/visualization.py:150-154): when image_u8 != NULL the rays are the
     * u-major even-spread lattice of ONE view (n_rays == image_side^2) and ray (iu, iv) also stores its pixel as uint8
     * clip(rint(255 * rgba), 0, 255) at image_u8[(iv * image_side + iu) * 4 ..], i.e. the (res,res,4) uint8 array the
     * reference returns after its reshape + transpose.  `rgba` may then be NULL. */
    uint8_t* image_u8;
    int32_t image_side;
} PlxRenderFwd;
int plx_render_fwd(const PlxRenderFwd* args, void* stream);

/*
 * K2 — fused backward: recomputes the march, runs the division-free reverse recurrence
 *   d alpha_k = T_k (v_k - S_k),  S_k = alpha_{k+1} v_{k+1} + (1 - alpha_{k+1}) S_{k+1},  d c_k = alpha_k T_k g_rgb
 * and scatter-ADDS into grad_grid (contiguous (X,Y,Z,4)) with warp-aggregated 16-byte vector reductions,
 * gated per channel by the clip pass-mask 0 <= raw <= 1 when PLX_CLAMP01 is set.
 * This is the autograd of scripts/train.py:146-151 triggered at :181 (cumprod/index/clamp backward, index_put accumulate).
 * `beta` adds the sparsity-loss term of scripts/train.py:170-177 with weight beta_over_m = beta / M_global (0 = off).
 * `tcarry` may be NULL (the kernel then recomputes the chunk transmittances in a first pass).
 */
typedef struct PlxRenderBwd {
    PlxMarch march;
    PlxRays rays;
    const float* grid;
    const float* grad_rgba;   /* (n_rays,4) */
    const float* tcarry;      /* from plx_render_fwd, or NULL */
    float* grad_grid;         /* (nx,ny,nz,4) contiguous, accumulated into */
    float beta_over_m;
} PlxRenderBwd;
int plx_render_bwd(const PlxRenderBwd* args, void* stream);

/*
 * K12 — fused training march (nearest lookup + mean-MSE): per ray, in one kernel,
 *   [generate the ray from (pose, uv)] -> forward march -> MSE gradient -> reverse march -> scatter-add into grad_grid.
 * Equivalent to plx_generate_rays + plx_render_fwd (MSE epilogue) + plx_render_bwd, i.e. scripts/train.py:130-157 + :181,
 * without the intermediate (N,.) buffers.  Rays come either from `rays` + `targets` (gen.uv == NULL) or are generated
 * in the kernel from gen.{imgs,poses,fov,uv} (src/ray_sampling.py:212-264; origins = poses[:, :3, 3]); then
 * rays.n_rays must equal gen.n_cams * gen.rays_per_cam.  `rgba` (n_rays,4) is optional.  loss[0] is accumulated into.
 * Returns PLX_E_UNSUPPORTED for trilinear mode or when num_samples exceeds the shared-memory index cache
 * (use K1 + K2 then; plx_train_step does that automatically).
 */
#define PLX_IMG_F32 0   /* imgs = (C,H,W,4) fp32 in [0,1], what load_image_data_from_path returns (src/data_processing.py:51-60) */
#define PLX_IMG_U8 1    /* imgs = (C,H,W,4) uint8 as decoded from the PNGs; the kernels evaluate fp32(u8) / 255 (IEEE division)
                           when they fetch a target pixel, bit-equal to the reference's conversion, at a quarter of the bytes */
typedef struct PlxRayGen {
    const void* imgs; int32_t n_cams, img_h, img_w;
    const float* poses; float fov;
    const float* uv; int32_t rays_per_cam;
    int32_t img_format;       /* PLX_IMG_F32 | PLX_IMG_U8 */
} PlxRayGen;
/*
 * Cross-GPU ordering fused into a kernel (multi-GPU only; all-zero = none).  `flags[r]` is rank r's int32 flag array
 * (PLX_BARRIER_CHANNELS x PLX_MAX_PEERS, symmetric / peer-mapped, zero before the first step; the array plx_peer_barrier
 * uses).  A kernel given a PlxPeerSync
 *   - at its START, if wait_epoch > 0: every block spins (bounded) until its OWN flags[rank][wait_channel][r] >= wait_epoch
 *     for every r < world, i.e. until every peer has signalled that epoch on that channel;
 *   - at its END, if signal_epoch > 0: the last block to finish (counted in `block_counter`, one device int32 that is 0
 *     before the first launch and wraps back to 0 by itself) stores signal_epoch into flags[r][signal_channel][rank] of
 *     every peer r with release semantics at system scope, after all of the grid's writes.
 * This replaces the two stand-alone barrier launches around plx_adam_step_peer: the march signals "my partial gradient
 * is complete", the exchange kernel waits for that from everyone and signals "my slab is stored in every replica", and the
 * next march waits for that.
 */
#define PLX_MAX_PEERS 8
/*
 * Failure record of the cross-GPU waits.  Every device-side wait is BOUNDED (`timeout_ns` of %globaltimer; 0 = 10 s): a peer
 * that died or stalled must not hang this GPU.  A wait that gives up stores a non-zero code
 *   (channel + 1) | (epoch << 8)   into *device_word (device memory) and into *host_word (pinned host memory, may be NULL)
 * and every kernel that is handed the same `device_word` and finds it non-zero at its start SKIPS its parameter / state
 * stores, so a step after a failed wait never writes stale gradients into any replica.  The host checks *host_word (or
 * copies *device_word back) in flush / wait_result / checkpoint and raises; the words are sticky until the caller clears
 * them.
 */
typedef struct PlxPeerError {
    int32_t* device_word;
    int32_t* host_word;
    uint64_t timeout_ns;
} PlxPeerError;

/*
 * Multi-GPU "push" exchange: where the fused march adds the gradient of cell `lin`.  Every rank r owns
 * a contiguous slab of cells, owner(lin) = umulhi(lin, owner_mul) (plx_slab_partition gives owner_mul and the slab
 * bounds), and the march issues its 16-byte vector reduction straight into the OWNER's gradient buffer
 * grads[owner] (peer-mapped, full grid size, only the owner's slab of each buffer is ever written).  After a cross-rank
 * barrier every rank holds the complete gradient sum of its slab, and only the touched cells crossed NVLink
 * (16 B x merged in-bounds samples instead of 16 B x cells).  world == 0: not used, the march adds into `grad_grid`.
 */
typedef struct PlxPeerGrad {
    float* grads[PLX_MAX_PEERS];
    uint32_t owner_mul;
    int32_t world;
} PlxPeerGrad;
/* Slab partition used by PlxPeerGrad / plx_adam_step_slab: owner_mul = floor(2^32 * world / n_cells) and rank r owns the
 * cells [begin, end) = [ceil(r * 2^32 / owner_mul), ceil((r+1) * 2^32 / owner_mul)) clipped to n_cells.  Needs
 * n_cells >= world.  Any output pointer may be NULL. */
int plx_slab_partition(int64_t n_cells, int32_t world, int32_t rank, uint32_t* owner_mul, int64_t* begin_cell, int64_t* end_cell);

typedef struct PlxPeerSync {
    int32_t* flags[PLX_MAX_PEERS];
    int32_t rank, world;
    int32_t wait_channel, wait_epoch;
    int32_t signal_channel, signal_epoch;
    int32_t* block_counter;
    PlxPeerError err;        /* where a wait that gave up is recorded (see PlxPeerError); all-zero = ~10 s bound, no record */
} PlxPeerSync;

/*
 * Device-resident step state: lets ONE captured CUDA graph replay every training step (scripts/train.py:104-191 has nothing
 * that changes from step to step except the uv draw, the loss slot parity and Adam's two bias-corrected scalars).  When
 * PlxTrainStep.replay != NULL the kernels take the step number from device memory instead of from `step`:
 *   this step  = *step_dev + 1                         (the march picks loss[step & 1]; the optimiser its walk direction)
 *   Adam scalars = table[step - table_base - 1]        = { sqrt(1 - beta2^step), -lr / (1 - beta1^step) } as fp32, formed on the
 *                                                        host in double exactly as plx_adam_step forms them (plx_adam_table)
 * and the LAST block of the optimiser kernel to finish stores the new step number into *step_dev (block_counter: one device
 * int32, zero before the first step, returns to zero by itself).  table_len entries; the caller refills the table (outside
 * the graph) before step - table_base exceeds it — the optimiser kernel records error code 0x7ab1e in *step_dev's sign bit
 * otherwise: it leaves *step_dev = -1 and applies nothing.
 */
typedef struct PlxReplayState {
    int32_t* step_dev;
    const float* table;      /* 2 floats per step */
    int64_t table_base;
    int32_t table_len;
    int32_t* block_counter;
} PlxReplayState;
/* fills `table_host` (2 * n floats, HOST memory) for the steps first_step .. first_step + n - 1 */
int plx_adam_table(double lr, double beta1, double beta2, int64_t first_step, int32_t n, float* table_host);

typedef struct PlxRenderTrain {
    PlxMarch march;
    PlxRays rays;
    const float* targets;
    PlxRayGen gen;
    const float* grid;
    float* grad_grid;
    float* rgba;
    float* loss;
    float grad_scale, loss_scale, beta_over_m;
    PlxPeerSync sync;        /* optional cross-GPU wait at the start / signal at the end (all-zero = none) */
    PlxPeerGrad peer_grad;   /* optional multi-GPU push exchange (world == 0: gradient goes to grad_grid) */
    const int32_t* step_dev; /* optional (graph replay): `loss` then points at TWO slots and the march adds into
                                loss[(*step_dev + 1) & 1] */
} PlxRenderTrain;
int plx_render_train(const PlxRenderTrain* args, void* stream);

/*
 * K3 — optimiser: one torch.optim.Adam step over n fp32 values (torch/optim/adam.py `_single_tensor_adam`, as used
 * at scripts/train.py:89,:180-182) fused with `grid_grad += |grad|` (:184) and, with zero_grad != 0, clearing `g` for
 * the next step (:180).  Scalars are formed in double like the Python code (bias corrections 1 - beta^step).
 * `gabs` may be NULL.  `step` counts from 1.
 */
int plx_adam_step(float* p, float* g, float* m, float* v, float* gabs, int64_t n, double lr, double beta1, double beta2,
                  double eps, int64_t step, int32_t zero_grad, void* stream);

/*
 * K3p — the optimiser step fused with the gradient exchange over NVLink peer memory (multi-GPU).
 * Every rank r holds a grid replica grids[r] and a local gradient buffer grads[r], all mapped into this process
 * (symmetric / peer memory).  The rank owns the element range [begin, end) (multiples of 4).  For every owned element
 * the kernel  (1) loads and sums the `world` partial gradients straight from the peers' buffers (reduce-scatter),
 * (2) applies the Adam update of plx_adam_step with the local exp_avg / exp_avg_sq / grad_abs_sum (full-size arrays,
 * only the owned range is touched), (3) stores the new parameter into ALL replicas (all-gather) — one pass, loads and
 * stores in flight together, no staging copy and no atomics on the fabric.  The caller orders it between two
 * cross-rank barriers (gradients complete before / parameters visible after) and clears its own gradient buffer.
 */
typedef struct PlxAdamPeer {
    int32_t world, rank;
    float* grids[PLX_MAX_PEERS];
    const float* grads[PLX_MAX_PEERS];
    float* exp_avg; float* exp_avg_sq; float* grad_abs_sum;
    int64_t begin, end;
    double lr, beta1, beta2, eps;
    int64_t step;
    /* Optional NVLS multicast addresses of the same two symmetric buffers (NULL = use the per-peer pointers above).
     * With them the partial gradients are summed INSIDE the NVSwitch (`multimem.ld_reduce.add.v4.f32`: one 16-byte
     * response per element instead of world-1) and the new parameters are written once and replicated by the switch
     * (`multimem.st.v4.f32`), which halves the bytes every GPU moves over its NVLink ports. */
    float* grid_mc;
    const float* grad_mc;
    /* optional step tail (same contract as plx_train_step / plx_train_step_host): publish the step's loss to result_host
     * { float loss; int32_t step; } and clear *loss_clear.  The loss is *loss_src (this rank's partial) or, when
     * loss_peers[0] != NULL, the sum over r < world of *loss_peers[r] in rank order (the global loss, identical bits on
     * every rank; also stored to *loss_out).  Any pointer may be NULL. */
    const float* loss_src;
    float* loss_clear;
    void* result_host;
    const float* loss_peers[PLX_MAX_PEERS];
    float* loss_out;
    PlxPeerSync sync;        /* optional fused barriers (all-zero = the caller orders the kernel with plx_peer_barrier) */
} PlxAdamPeer;
int plx_adam_step_peer(const PlxAdamPeer* args, void* stream);

/*
 * K3s — the optimiser step of the multi-GPU "push" exchange (PlxPeerGrad): after the cross-rank barrier that follows the
 * march, `grad` (this rank's own gradient buffer) holds the COMPLETE gradient sum of the owned slab [begin, end) (element
 * offsets, multiples of 4), because every rank's march reduced its contributions straight into it.  The kernel is
 * plx_adam_step over that slab with its single-GPU refinements (no |grad| store / gradient clear for untouched cells,
 * alternating walk) whose parameter store goes to ALL replicas: one `multimem.st.v4.f32` through the NVSwitch when
 * `grid_mc` is given, else one 16-byte store per peer pointer in `grids` (grids[rank] = the local replica).  The gradient
 * slab is cleared in place (the next march may only start after the closing barrier, so one buffer suffices).
 * Step tail (any pointer may be NULL): thread 0 sums loss_peers[r][0] over r < world IN RANK ORDER (every rank obtains the
 * same bits) = the global loss of scripts/train.py:156, stores it to *loss_out and to result_host { float loss; int32
 * step; }, and clears *loss_clear.  `err`: see PlxPeerError (a recorded failure skips every store of this kernel).
 */
typedef struct PlxAdamSlab {
    int32_t world, rank;
    float* grids[PLX_MAX_PEERS];
    float* grid_mc;
    float* grad;
    float* exp_avg; float* exp_avg_sq; float* grad_abs_sum;
    int64_t begin, end;
    double lr, beta1, beta2, eps;
    int64_t step;
    const float* loss_peers[PLX_MAX_PEERS];
    float* loss_out;
    float* loss_clear;
    void* result_host;
    PlxPeerError err;
} PlxAdamSlab;
int plx_adam_step_slab(const PlxAdamSlab* args, void* stream);

/*
 * Cross-GPU barrier on the launching stream (one tiny kernel): rank `rank` stores `epoch` into slot [channel][rank] of
 * every peer's flag array (release, system scope) and then waits until all `world` slots of its OWN array hold a value
 * >= epoch (acquire).  flags[r] = peer-mapped pointer to rank r's int32[PLX_BARRIER_CHANNELS * PLX_MAX_PEERS] array in
 * symmetric memory, zero-initialised; `epoch` must increase by one per call and channel.  Orders everything enqueued
 * before it on every rank's stream before everything enqueued after it on this rank's stream.  The wait is bounded and a
 * give-up is recorded in `err` (may be NULL: 10 s, unrecorded) — see PlxPeerError.
 */
#define PLX_BARRIER_CHANNELS 4
int plx_peer_barrier(int32_t* const* flags, int32_t rank, int32_t world, int32_t channel, int32_t epoch,
                     const PlxPeerError* err, void* stream);

/*
 * Ray generation — generate_rays_batched, src/ray_sampling.py:195-264.
 * imgs (C,H,W,4); poses (C,4,4) row-major camera-to-world; uv (C,R,2) in [0,1] (the `torch.rand` draw of :227) or NULL
 * for the even-spread lattice of :220-223 with R = n_side^2 rays (u-major, linspace(0,1,n_side)).
 * Writes dirs (C*R,3) and, when imgs/targets are non-NULL, targets (C*R,4) = imgs[cam, v_pix, u_pix] (:238-248).
 */
int plx_generate_rays(const float* imgs, int32_t n_cams, int32_t img_h, int32_t img_w, const float* poses, float fov,
                      const float* uv, int32_t rays_per_cam, int32_t n_side, float* dirs, float* targets, void* stream);
/* the same with the image set described by a PlxRayGen (gen->uv / gen->rays_per_cam as above; uint8 images allowed) */
int plx_generate_rays_gen(const PlxRayGen* gen, int32_t n_side, float* dirs, float* targets, void* stream);

/* Eager per-function kernels (reference semantics, materialised tensors) ------------------------------------------- */

/* samples (n_rays*S,3) = o + d * (delta*k) — sample_camera_rays_batched, src/ray_sampling.py:161-167 */
int plx_sample_points(const PlxRays* rays, int32_t num_samples, float delta_step, float* samples, void* stream);

/* out (M,3) = (samples - gmin) / pd — normalize_samples_for_indecies, src/ray_sampling.py:12-13 */
int plx_normalize_points(const float* samples, int64_t m, const float gmin[3], float points_distance, float* out,
                         void* stream);

/* get_nearest_voxels, src/grid_functions.py:103-114: vals (M,4) at periodically wrapped indices (unmasked),
 * inbounds (M) uint8; idx_out (M,3) int64 wrapped indices or NULL.  Grid given with explicit strides. */
int plx_gather_nearest(const float* ns, int64_t m, const float* grid, const int32_t dims[3], const int64_t strides[4],
                       float* vals, uint8_t* inbounds, int64_t* idx_out, void* stream);
/* its autograd: grad_grid[(wrapped idx)] += grad_vals, grad_grid contiguous (X,Y,Z,4) */
int plx_gather_nearest_bwd(const float* ns, int64_t m, const float* grad_vals, const int32_t dims[3], float* grad_grid,
                           void* stream);

/* trilinear lookup: corners of get_grid_points_indices (src/grid_functions.py:220-246) wrapped periodically (:66-79),
 * interpolated as trilinear_interpolation (:7-44).  `masked` != 0 multiplies by the float in-bounds test. */
int plx_trilinear_fwd(const float* ns, int64_t m, const float* grid, const int32_t dims[3], const int64_t strides[4],
                      int32_t masked, float* vals, uint8_t* inbounds, void* stream);
int plx_trilinear_bwd(const float* ns, int64_t m, const float* grad_vals, const int32_t dims[3], int32_t masked,
                      float* grad_grid, void* stream);

/* compute_alpha_weighted_pixels, src/ray_sampling.py:172-192: samples (n_rays,S,4) -> out (n_rays,4) */
int plx_composite_fwd(const float* samples, int64_t n_rays, int32_t num_samples, float* out, void* stream);
/* its autograd: grad_samples (n_rays,S,4) from grad_out (n_rays,4) */
int plx_composite_bwd(const float* samples, int64_t n_rays, int32_t num_samples, const float* grad_out,
                      float* grad_samples, void* stream);

/*
 * average_pool3d_grid — src/grid_functions.py:173-181 (F.avg_pool3d, cubic window `kernel`, stride `stride`, no padding)
 * on a contiguous (X,Y,Z,4) grid, as three separable box passes; `out` is contiguous (Ox,Oy,Oz,4) with
 * O = (dim - kernel) / stride + 1.  Scratch: tmp1 (X*Y*Oz*4 floats), tmp2 (X*Oy*Oz*4 floats).
 * plx_avgpool3d_bwd is its autograd (grad_out (Ox,Oy,Oz,4) -> grad_in (X,Y,Z,4), overwritten), same scratch sizes.
 */
int plx_avgpool3d_fwd(const float* in, const int32_t dims[3], int32_t kernel, int32_t stride, float* tmp1, float* tmp2,
                      float* out, void* stream);
int plx_avgpool3d_bwd(const float* grad_out, const int32_t dims[3], int32_t kernel, int32_t stride, float* tmp2, float* tmp1,
                      float* grad_in, void* stream);

/*
 * tv_loss — scripts/train.py:44-65 with its weight (:163-168): loss_out[0] = tv * sqrt(sum of squared neighbour
 * differences of the contiguous (X,Y,Z,4) grid along the three axes) and, when `grad` != NULL,
 * grad += d(loss)/d(grid) (6-neighbour stencil).  `scratch` = 1 double of device memory.  Where the reference's gradient
 * is 0/0 (a constant grid) nothing is added.
 */
int plx_tv_loss(const float* grid, const int32_t dims[3], float tv, float* grad, double* scratch, float* loss_out, void* stream);
/* the same with the gradient added only for the cells [cell_begin, cell_end) (linear cell indices): the share of a rank that
 * owns that slab of the gradient in the multi-GPU exchange; the loss value is still that of the whole grid.  `atomic` != 0
 * adds with 16-byte reductions instead of read-modify-write (other GPUs may be reducing into the same buffer meanwhile).
 * Multi-GPU callers must run it while the replica is stable: after the march, BEFORE the barrier that releases the peers'
 * optimiser kernels (those store new parameters into this replica). */
int plx_tv_loss_range(const float* grid, const int32_t dims[3], float tv, float* grad, int64_t cell_begin, int64_t cell_end,
                      int32_t atomic, double* scratch, float* loss_out, void* stream);

/*
 * visulize_3d_in_2d_fast — src/visualization.py:157-232, the point-splat preview scripts/compare_inference_to_image.py:58
 * calls: every cell of the contiguous (X,Y,Z,4) grid with alpha > 0.1 is projected through the camera (`pose_host`: 16 floats
 * of HOST memory, the row-major 4x4 camera-to-world matrix) onto an (xs, ys) image and the point NEAREST to the camera wins
 * each pixel (what the reference's far-to-near assignment order leaves behind); untouched pixels are 1.  Two passes, no sort:
 * a 64-bit atomicMin per point on (distance bits, cell index), then a resolve of the winners' RGB.
 * zbuf: xs * ys uint64 of device scratch; image: (xs, ys, 3) fp32, xs = int(ys * aspect) as the reference computes it.
 */
int plx_splat_view(const float* grid, const int32_t dims[3], float points_distance, const float* pose_host, float fov,
                   int32_t xs, int32_t ys, uint64_t* zbuf, float* image, void* stream);

/*
 * A/B switches for measurements (tools/ and tests; process-wide, results are identical either way — they only move launch
 * shapes and skip redundant stores): "adam_skip_same" (-1 = auto), "adam_blocks_per_sm" (4), "train_wpb" (4), "pdl" (1).
 */
int plx_tune(const char* name, int32_t value);

/*
 * Self-test of the library's exact fp32 helpers: evaluates the hoisted-reciprocal quotient x / y (the march's divisor
 * path), the inlined square root and the per-element quotient (Adam) on `n` pseudo-random inputs and counts results
 * that differ in any bit from CUDA's IEEE intrinsics (__fdiv_rn / __fsqrt_rn).  `mismatches` = 3 device uint64 counters
 * (caller zeroes them): [0] x / y, [1] sqrt, [2] x / d.  Parity infrastructure; not used by the product path.
 */
int plx_selftest_arith(float y, uint64_t n, uint64_t seed, uint64_t* mismatches, void* stream);

/*
 * One whole training step of scripts/train.py:130-184 (tv = 0) in a single host call:
 *   plx_generate_rays -> plx_render_fwd (+MSE epilogue) -> plx_render_bwd -> [caller's collective] -> plx_adam_step.
 * `phase` selects which part runs so a multi-GPU caller can put the gradient all-reduce between the two halves:
 *   PLX_STEP_RENDER = rays + forward + loss + backward;  PLX_STEP_OPTIM = Adam;  PLX_STEP_ALL = both.
 * Scratch (dirs, targets, rgba, grad_rgba, tcarry) is caller-provided.
 * `loss` points to TWO floats that the caller zeroes once: step s accumulates its loss into loss[s & 1] and the optimiser
 * phase clears loss[(s + 1) & 1] for the next step, so no per-step memset is needed and loss[s & 1] stays readable
 * until step s + 2 runs.
 */
#define PLX_STEP_RENDER 1
#define PLX_STEP_OPTIM 2
#define PLX_STEP_ALL 3
#define PLX_STEP_UNFUSED 4  /* OR-ed into `phase`: render with plx_generate_rays + plx_render_fwd + plx_render_bwd (three kernels and
                               their (N,.) buffers) instead of the fused march — the path the fused kernel is checked against */
typedef struct PlxTrainStep {
    PlxMarch march;
    /* scene, resident on the device (imgs: fp32 or uint8 RGBA, see img_format at the end of the struct) */
    const void* imgs; int32_t n_cams, img_h, img_w;
    const float* poses; float fov;
    /* this step's rays */
    const float* uv; int32_t rays_per_cam;
    int64_t n_rays_global;    /* N over all ranks: mean-MSE scale 1/(4*N_global) so a plain SUM of grads is exact */
    /* state */
    float* grid; float* grad; float* exp_avg; float* exp_avg_sq; float* grad_abs_sum;
    double lr, beta1, beta2, eps; int64_t step;
    float beta_over_m;
    /* scratch */
    float* dirs; float* targets; float* rgba; float* grad_rgba; float* tcarry;
    /* result */
    float* loss;
    /* optional (multi-GPU): cross-GPU wait / signal fused into the march of the render phase (see PlxPeerSync) */
    const PlxPeerSync* render_sync;
    /* optional (multi-GPU): push exchange — the march reduces into the slab owners' buffers instead of `grad` (PlxPeerGrad) */
    const PlxPeerGrad* peer_grad;
    int32_t img_format;       /* PLX_IMG_F32 | PLX_IMG_U8 */
    /* optional: device-resident step state for CUDA-graph replay (see PlxReplayState); `step` is then ignored */
    const PlxReplayState* replay;
} PlxTrainStep;
int plx_train_step(const PlxTrainStep* args, int32_t phase, void* stream);

/*
 * The same step driven from HOST buffers (the end-to-end path), still exactly two kernel launches:
 *   - `uv_host` (C*R*2 floats) must be PINNED host memory (cudaHostAlloc / cudaHostRegister / torch pin_memory): the march
 *     kernel reads this step's uv straight out of it over PCIe (zero-copy, 8 bytes per ray), no staging copy;
 *   - `result_host` (pinned, 8 bytes) receives { float loss; int32_t step; }: the optimiser kernel stores the step's loss
 *     and then, after a system-scope fence, the step number.  The host may poll result_host[1] == step (or synchronise
 *     the stream) and then read the loss.
 * Asynchronous on `stream`.  With phase = PLX_STEP_RENDER only, nothing is published (the caller's own optimiser
 * phase does that: plx_train_step_host(..., PLX_STEP_OPTIM) or plx_adam_step_peer).
 */
int plx_train_step_host(const PlxTrainStep* args, const float* uv_host, void* result_host, int32_t phase, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLENOXEL_ABI_H_ */
