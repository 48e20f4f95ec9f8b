// plx_launch.h — internal launcher prototypes (one per kernel family); the extern "C" layer in plx_abi.cu validates
// arguments and calls these.  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/plenoxel_abi.h"

namespace plx {

constexpr int MAX_WARPS_PER_BLOCK = 8;

// Tuning switches for A/B measurements (plx_tune in the ABI; tools/ and tests only).  Defaults = the measured optimum.
struct Tuning {
    int adam_skip_same = -1;     // Adam kernels: do not store lines of parameters / moments that did not change (-1 = auto: always in
                                 // the multi-GPU slab kernel, where they would cross NVLink; on one GPU only for grids beyond the L2 —
                                 // 256^3: 399 -> 387 us per step, 128^3 where every line changes: 89.5 -> 90.6)
    int adam_blocks_per_sm = 4;  // resident 256-thread blocks per SM of the Adam kernels
    int train_wpb = 4;           // warps per block of the fused training march
    int train_cache_it = 0;      // trilinear fused march: iterations the per-warp value cache holds (0 = sized from the grid's diagonal;
                                 // tests force 1 so that every ray takes the uncached path)
    int packet_tile = 1;         // ray-packet inference kernel marches 16 x 8 lattice tiles per block (0 = 128 consecutive rays)
    int pdl = 1;                 // programmatic dependent launch of the march / optimiser kernels (launch latency behind the previous kernel's tail)
};
Tuning& tuning();

// Launch `kernel` so that it may be scheduled while the previous kernel of the stream drains (programmatic stream
// serialisation).  The kernel must execute grid_dependency_wait() (plx_device.cuh) before it touches anything the previous
// kernel wrote; with that, results are identical to a plain launch.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned blocks, unsigned threads, size_t smem, cudaStream_t st, Args... args) {
    if (!tuning().pdl) {
        kernel<<<blocks, threads, smem, st>>>(args...);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

cudaError_t launch_render_fwd(const PlxRenderFwd& a, cudaStream_t st);
cudaError_t launch_render_bwd(const PlxRenderBwd& a, cudaStream_t st);
bool render_train_supported(const PlxRenderTrain& a);
cudaError_t launch_render_train(const PlxRenderTrain& a, cudaStream_t st);

struct AdamScalars {
    float one_minus_beta1;   // lerp weight
    float beta2;
    float one_minus_beta2;
    float bc2_sqrt;          // sqrt(1 - beta2^step)
    float eps;
    float neg_step_size;     // -lr / (1 - beta1^step)
    bool keep_p, keep_g;     // L2 evict_last tags for parameters / gradient (set by launch_adam)
    bool reverse;            // walk the arrays from the top down (odd steps): the tail the previous step left in L2 is read first
    bool skip_same;          // do not store values that did not change (set by the launchers from tuning())
};
// what one thread of the optimiser kernel does besides Adam: publish this step's loss to pinned host memory
// { float loss; int32 step } and clear the other loss slot for the next step
struct StepTail {
    const float* src;        // this rank's loss accumulator of the step ...
    float* clear;
    float* dst_host;
    int32_t step;
    // ... or (multi-GPU) every rank's accumulator, peer-mapped: summed in rank order = the global loss, same bits on every rank;
    // the sum is also stored to *global_out (device) so that `loss` reads the same on the device path
    const float* src_peers[PLX_MAX_PEERS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int32_t n_peers = 0;
    float* global_out = nullptr;
};
// graph replay (PlxReplayState): step number, loss slots and Adam scalars come from device memory
struct ReplayArgs {
    int32_t* step_dev = nullptr;
    const float* table = nullptr;
    int64_t table_base = 0;
    int32_t table_len = 0;
    int32_t* block_counter = nullptr;
    float* loss2 = nullptr;          // the two loss slots
};
cudaError_t launch_adam(float* p, float* g, float* m, float* v, float* gabs, int64_t n, const AdamScalars& s,
                        bool zero_grad, const StepTail& tail, cudaStream_t st, const ReplayArgs& replay = ReplayArgs());

cudaError_t launch_adam_peer(const PlxAdamPeer& a, const AdamScalars& s, const StepTail& tail, cudaStream_t st);
cudaError_t launch_adam_slab(const PlxAdamSlab& a, const AdamScalars& s, cudaStream_t st);

cudaError_t launch_avgpool3d_fwd(const float* in, const int32_t* dims, int k, int s, float* tmp1, float* tmp2, float* out,
                                 cudaStream_t st);
cudaError_t launch_avgpool3d_bwd(const float* gout, const int32_t* dims, int k, int s, float* tmp2, float* tmp1, float* gin,
                                 cudaStream_t st);
cudaError_t launch_tv_loss(const float* grid, const int32_t* dims, float tv, float* grad, int64_t cell_begin, int64_t cell_end,
                           bool atomic, double* scratch, float* loss_out, cudaStream_t st);
cudaError_t launch_splat_view(const float* grid, const int32_t* dims, float pd, const float* pose16, float fov, int xs, int ys,
                              unsigned long long* zbuf, float* image, cudaStream_t st);
cudaError_t launch_peer_barrier(int32_t* const* flags, int rank, int world, int channel, int epoch, const PlxPeerError& err,
                                cudaStream_t st);

cudaError_t launch_generate_rays(const PlxRayGen& gen, int n_side, float* dirs, float* targets, cudaStream_t st);
cudaError_t launch_sample_points(const PlxRays& rays, int S, float delta, float* out, cudaStream_t st);
cudaError_t launch_normalize_points(const float* in, int64_t m, float gx, float gy, float gz, float pd, float* out,
                                    cudaStream_t st);
cudaError_t launch_gather_nearest(const float* ns, int64_t m, const float* grid, const int32_t* dims,
                                  const int64_t* strides, float* vals, uint8_t* inb, int64_t* idx_out, cudaStream_t st);
cudaError_t launch_gather_nearest_bwd(const float* ns, int64_t m, const float* grad_vals, const int32_t* dims,
                                      float* grad_grid, cudaStream_t st);
cudaError_t launch_trilinear_fwd(const float* ns, int64_t m, const float* grid, const int32_t* dims,
                                 const int64_t* strides, int masked, float* vals, uint8_t* inb, cudaStream_t st);
cudaError_t launch_trilinear_bwd(const float* ns, int64_t m, const float* grad_vals, const int32_t* dims, int masked,
                                 float* grad_grid, cudaStream_t st);
cudaError_t launch_composite_fwd(const float* samples, int64_t n_rays, int S, float* out, cudaStream_t st);
cudaError_t launch_selftest(float y, uint64_t n, uint64_t seed, unsigned long long* bad, cudaStream_t st);
cudaError_t launch_composite_bwd(const float* samples, int64_t n_rays, int S, const float* grad_out, float* grad_samples,
                                 cudaStream_t st);

}  // namespace plx
