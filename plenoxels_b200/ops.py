"""torch.autograd wrappers over the C ABI (include/plenoxel_abi.h).

Tensors in, tensors out; every function launches hand-written sm_100a kernels through `libplenoxel_b200.so`
on the current CUDA stream.  No ATen arithmetic on the hot path: torch is used to own device memory.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L


# --------------------------------------------------------------------------------------------- geometry helpers
def grid_origin(dims, points_distance: float, start_index=0):
    """World coordinate of cell `start_index` (an int or one int per axis) as three fp32 values, computed on the host exactly as
    `generate_grid` builds them (src/grid_functions.py:205-209): fl32(i - ceil(s/2) + 1) * fl32(pd).
    Equals `grid_indices.min(0)[0]` of src/ray_sampling.py:13 for pd > 0 without the (G^3,3) reduction."""
    pd32 = np.float32(points_distance)
    starts = (start_index,) * 3 if isinstance(start_index, int) else tuple(start_index)
    return tuple(float(np.float32(int(i) - math.ceil(int(s) / 2) + 1) * pd32) for i, s in zip(starts, dims[:3]))


# --------------------------------------------------------------------------------------------- fused march (K1 / K2)
class _RenderRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grid, origins, dirs, rays_per_origin, num_samples, delta_step, gmin, points_distance, mode, clamp,
                want_depth, want_count, beta_over_m, coherent):
        dev = L.require_cuda(grid, origins, dirs)
        lib = L.load()
        n = dirs.shape[0]
        need_grad = grid.requires_grad
        march = L.make_march(grid, num_samples, delta_step, gmin, points_distance, mode, clamp,
                             coherent=coherent and not need_grad and not want_count)
        rays = L.make_rays(origins, dirs, rays_per_origin)
        rgba = torch.empty((n, 4), dtype=torch.float32, device=dev)
        depth = torch.empty((n,), dtype=torch.float32, device=dev) if want_depth else None
        count = torch.empty((n,), dtype=torch.int32, device=dev) if want_count else None
        tcarry = torch.empty((n, lib.plx_num_chunks(num_samples)), dtype=torch.float32, device=dev) if need_grad else None
        a = L.PlxRenderFwd()
        a.march, a.rays = march, rays
        a.grid, a.rgba, a.depth, a.count = grid.data_ptr(), rgba.data_ptr(), L.ptr(depth), L.ptr(count)
        a.sample_index, a.tcarry, a.targets, a.grad_rgba, a.loss = None, L.ptr(tcarry), None, None, None
        with torch.cuda.device(dev):
            L.check(lib.plx_render_fwd(C.byref(a), L.stream_ptr(dev)), "plx_render_fwd")
        ctx.march_args = (rays_per_origin, num_samples, delta_step, gmin, points_distance, mode, clamp, beta_over_m)
        ctx.save_for_backward(grid, origins, dirs, tcarry)
        outs = (rgba,)
        if want_depth:
            ctx.mark_non_differentiable(depth)
            outs += (depth,)
        if want_count:
            ctx.mark_non_differentiable(count)
            outs += (count,)
        return outs if len(outs) > 1 else rgba

    @staticmethod
    def backward(ctx, grad_rgba, *unused):
        grid, origins, dirs, tcarry = ctx.saved_tensors
        rays_per_origin, num_samples, delta_step, gmin, points_distance, mode, clamp, beta_over_m = ctx.march_args
        dev = grid.device
        lib = L.load()
        grad_grid = torch.zeros(grid.shape, dtype=torch.float32, device=dev)       # contiguous (X,Y,Z,4)
        b = L.PlxRenderBwd()
        b.march = L.make_march(grid, num_samples, delta_step, gmin, points_distance, mode, clamp)
        b.rays = L.make_rays(origins, dirs, rays_per_origin)
        g = grad_rgba.contiguous().float()
        b.grid, b.grad_rgba, b.tcarry, b.grad_grid = grid.data_ptr(), g.data_ptr(), L.ptr(tcarry), grad_grid.data_ptr()
        b.beta_over_m = float(beta_over_m)
        with torch.cuda.device(dev):
            L.check(lib.plx_render_bwd(C.byref(b), L.stream_ptr(dev)), "plx_render_bwd")
        return (grad_grid,) + (None,) * 13


def render_rays(grid, origins, dirs, num_samples, delta_step, gmin, points_distance, mode="nearest", clamp=True,
                rays_per_origin=1, return_depth=False, return_count=False, beta_over_m=0.0, coherent=False):
    """Fused march: rgba (N,4) [, depth (N,), count (N,) int32] of rays through `grid` (X,Y,Z,4).

    One kernel for the sequence scripts/train.py:130-151: sample_camera_rays_batched (src/ray_sampling.py:161-167),
    normalize_samples_for_indecies (:13), get_nearest_voxels on grid.clip(0,1) (src/grid_functions.py:103-114) or the
    trilinear lookup (:7-44,:220-246), mask multiply, compute_alpha_weighted_pixels (src/ray_sampling.py:172-192).
    Differentiable w.r.t. `grid` (K2).  `beta_over_m` = beta / M adds the gradient of the sparsity loss of
    scripts/train.py:170-177 in the backward pass (its value is not part of the returned pixels).
    `coherent=True` (inference only: ignored when a gradient or the count is requested) tells the library that consecutive
    rays are neighbouring pixels of a view; it then marches one ray per thread in packets of 32 (PLX_COHERENT_RAYS).
    """
    return _RenderRays.apply(grid, origins, dirs, int(rays_per_origin), int(num_samples), float(delta_step),
                             tuple(float(x) for x in gmin), float(points_distance), mode, bool(clamp),
                             bool(return_depth), bool(return_count), float(beta_over_m), bool(coherent))


def render_image_u8(grid, pose, fov, side, num_samples, delta_step, gmin, points_distance, mode="nearest", clamp=False):
    """The (side, side, 4) uint8 image `visulize_3d_in_2d` returns (src/visualization.py:111-154) for ONE camera `pose` (1,4,4):
    even-spread lattice rays (ray generation kernel), coherent ray-packet march, and the x255 / round-half-even / clip / uint8 /
    transpose epilogue written by the march kernel itself — two launches, no (N,4) float image, nothing on the host."""
    dev = L.require_cuda(grid, pose)
    lib = L.load()
    pose = pose.reshape(-1, 4, 4)[:1].contiguous().float()
    n = int(side) * int(side)
    dirs, _ = generate_rays(None, pose, fov, uv=None, rays_per_cam=n, want_targets=False)
    img = torch.empty((int(side), int(side), 4), dtype=torch.uint8, device=dev)
    a = L.PlxRenderFwd()
    a.march = L.make_march(grid, num_samples, delta_step, gmin, points_distance, mode, clamp, coherent=True)
    a.rays = L.make_rays(pose[:, :3, 3], dirs, n)
    a.grid, a.image_u8, a.image_side = grid.data_ptr(), img.data_ptr(), int(side)
    with torch.cuda.device(dev):
        L.check(lib.plx_render_fwd(C.byref(a), L.stream_ptr(dev)), "plx_render_fwd(image)")
    return img


def sample_indices(grid, origins, dirs, num_samples, delta_step, gmin, points_distance, mode="nearest",
                   rays_per_origin=1):
    """Parity/debug dump: (N,S) int32 linear cell index ((ix*Y+iy)*Z+iz) of every sample, -1 when out of bounds,
    plus the per-ray in-bounds count (N,) int32 — the exact index arithmetic of K1 without clipping."""
    dev = L.require_cuda(grid, origins, dirs)
    lib = L.load()
    n = dirs.shape[0]
    idx = torch.empty((n, num_samples), dtype=torch.int32, device=dev)
    rgba = torch.empty((n, 4), dtype=torch.float32, device=dev)
    count = torch.empty((n,), dtype=torch.int32, device=dev)
    a = L.PlxRenderFwd()
    a.march = L.make_march(grid, num_samples, delta_step, gmin, points_distance, mode, True)
    a.rays = L.make_rays(origins, dirs, rays_per_origin)
    a.grid, a.rgba, a.count, a.sample_index = grid.data_ptr(), rgba.data_ptr(), count.data_ptr(), idx.data_ptr()
    with torch.cuda.device(dev):
        L.check(lib.plx_render_fwd(C.byref(a), L.stream_ptr(dev)), "plx_render_fwd")
    return idx, count


def render_train(grid, grad_grid, num_samples, delta_step, gmin, points_distance, *, origins=None, dirs=None,
                 targets=None, rays_per_origin=1, imgs=None, poses=None, fov=None, uv=None, n_rays_global=None,
                 beta_over_m=0.0, clamp=True, mode="nearest", peer_grads=None):
    """K12, the fused training march (nearest or trilinear lookup): forward + mean-MSE + backward in one kernel; the gradient is
    ACCUMULATED into `grad_grid` (contiguous (X,Y,Z,4)).  Rays are either given (`origins`, `dirs`, `targets`) or
    generated in the kernel from (`imgs`, `poses`, `fov`, `uv` (C,R,2)).  Returns (rgba (N,4), loss (1,) device tensor)
    = the pixels and mean((rgba - targets)^2) of scripts/train.py:151-156.  `imgs` may be fp32 in [0,1] or uint8 (the
    target pixel is then converted in the kernel, fp32(u8) / 255).  `peer_grads`: list of W grid-shaped buffers (push exchange,
    PlxPeerGrad): the gradient of cell `lin` is added to peer_grads[owner(lin)] instead of `grad_grid`, owner = the slab
    partition of plx_slab_partition — on a multi-GPU run these are the ranks' peer-mapped buffers."""
    dev = L.require_cuda(grid, grad_grid, origins, dirs, targets, imgs, poses, uv)
    lib = L.load()
    a = L.PlxRenderTrain()
    a.march = L.make_march(grid, num_samples, delta_step, gmin, points_distance, mode, clamp)
    keep = []
    if uv is not None:
        uv = uv.contiguous().float()
        poses = poses.contiguous().float()
        imgs = imgs.contiguous() if imgs.dtype == torch.uint8 else imgs.contiguous().float()
        keep += [uv, poses, imgs]
        n = uv.shape[0] * uv.shape[1]
        a.rays.n_rays = n
        a.gen.imgs, a.gen.n_cams, a.gen.img_h, a.gen.img_w = imgs.data_ptr(), imgs.shape[0], imgs.shape[1], imgs.shape[2]
        a.gen.img_format = L.PLX_IMG_U8 if imgs.dtype == torch.uint8 else L.PLX_IMG_F32
        a.gen.poses, a.gen.fov, a.gen.uv, a.gen.rays_per_cam = poses.data_ptr(), float(fov), uv.data_ptr(), uv.shape[1]
    else:
        a.rays = L.make_rays(origins, dirs, rays_per_origin)
        targets = targets.contiguous().float()
        keep.append(targets)
        n = dirs.shape[0]
        a.targets = targets.data_ptr()
    if not grad_grid.is_contiguous() or grad_grid.shape != grid.shape or grad_grid.dtype != torch.float32:
        raise L.PlxError("grad_grid must be a contiguous float32 tensor of the grid's shape")
    if peer_grads:
        mul = C.c_uint32()
        L.check(lib.plx_slab_partition(grid.numel() // 4, len(peer_grads), 0, C.byref(mul), None, None), "plx_slab_partition")
        for r, buf in enumerate(peer_grads):
            if not buf.is_contiguous() or buf.shape != grid.shape or buf.dtype != torch.float32:
                raise L.PlxError("peer_grads must be contiguous float32 tensors of the grid's shape")
            a.peer_grad.grads[r] = buf.data_ptr()
        a.peer_grad.owner_mul, a.peer_grad.world = mul.value, len(peer_grads)
    n_glob = n if n_rays_global is None else int(n_rays_global)
    rgba = torch.empty((n, 4), dtype=torch.float32, device=dev)
    loss = torch.zeros((1,), dtype=torch.float32, device=dev)
    a.grid, a.grad_grid, a.rgba, a.loss = grid.data_ptr(), grad_grid.data_ptr(), rgba.data_ptr(), loss.data_ptr()
    a.grad_scale, a.loss_scale, a.beta_over_m = 2.0 / (4.0 * n_glob), 1.0 / (4.0 * n_glob), float(beta_over_m)
    with torch.cuda.device(dev):
        L.check(lib.plx_render_train(C.byref(a), L.stream_ptr(dev)), "plx_render_train")
    return rgba, loss


# --------------------------------------------------------------------------------------------- optimiser (K3)
def adam_step(p, g, m, v, gabs, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, zero_grad=True):
    """In-place Adam step + `gabs += |g|` + optional `g = 0` (scripts/train.py:180-184) over contiguous fp32 tensors."""
    dev = L.require_cuda(p, g, m, v, gabs)
    for t in (p, g, m, v, gabs):
        if t is not None and (not t.is_contiguous() or t.dtype != torch.float32 or t.numel() != p.numel()):
            raise L.PlxError("adam_step needs contiguous float32 tensors of equal size")
    with torch.cuda.device(dev):
        L.check(L.load().plx_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), L.ptr(gabs), p.numel(),
                                       float(lr), float(beta1), float(beta2), float(eps), int(step), int(zero_grad),
                                       L.stream_ptr(dev)), "plx_adam_step")


# --------------------------------------------------------------------------------------------- ray generation
def generate_rays(imgs, poses, fov, uv=None, rays_per_cam=None, want_targets=True):
    """dirs (C*R,3), targets (C*R,4) — generate_rays_batched (src/ray_sampling.py:195-264) for given uv (C,R,2),
    or the even-spread lattice (uv=None, rays_per_cam = n_side^2)."""
    dev = L.require_cuda(imgs, poses, uv)
    poses = poses.contiguous().float()
    C_ = poses.shape[0]
    if uv is not None:
        uv = uv.contiguous().float()
        R, n_side = uv.shape[1], 0
    else:
        R = int(rays_per_cam)
        n_side = int(round(math.sqrt(R)))
    dirs = torch.empty((C_ * R, 3), dtype=torch.float32, device=dev)
    targets = None
    H = W = 0
    gen = L.PlxRayGen()
    if want_targets:
        if imgs.dtype != torch.uint8 and (imgs.dtype != torch.float32 or not imgs.is_contiguous()):
            imgs = imgs.contiguous().float()
        imgs = imgs.contiguous()
        H, W = imgs.shape[1], imgs.shape[2]
        targets = torch.empty((C_ * R, 4), dtype=torch.float32, device=dev)
        gen.imgs = imgs.data_ptr()
        gen.img_format = L.PLX_IMG_U8 if imgs.dtype == torch.uint8 else L.PLX_IMG_F32
    gen.n_cams, gen.img_h, gen.img_w, gen.poses, gen.fov, gen.uv, gen.rays_per_cam = C_, H, W, poses.data_ptr(), float(fov), L.ptr(uv), R
    with torch.cuda.device(dev):
        L.check(L.load().plx_generate_rays_gen(C.byref(gen), n_side, dirs.data_ptr(), L.ptr(targets), L.stream_ptr(dev)),
                "plx_generate_rays")
    return dirs, targets


# --------------------------------------------------------------------------------------------- eager kernels
def sample_points(origins, dirs, num_samples, delta_step, rays_per_origin=1):
    """(N*S,3) sample positions — src/ray_sampling.py:161-167."""
    dev = L.require_cuda(origins, dirs)
    n = dirs.shape[0]
    out = torch.empty((n * num_samples, 3), dtype=torch.float32, device=dev)
    rays = L.make_rays(origins, dirs, rays_per_origin)
    with torch.cuda.device(dev):
        L.check(L.load().plx_sample_points(C.byref(rays), int(num_samples), float(delta_step), out.data_ptr(),
                                           L.stream_ptr(dev)), "plx_sample_points")
    return out


def normalize_points(samples, gmin, points_distance):
    """(samples - gmin) / pd — src/ray_sampling.py:13."""
    dev = L.require_cuda(samples)
    s = samples.contiguous().float()
    out = torch.empty_like(s)
    g = (C.c_float * 3)(*[float(x) for x in gmin])
    with torch.cuda.device(dev):
        L.check(L.load().plx_normalize_points(s.data_ptr(), s.numel() // 3, g, float(points_distance), out.data_ptr(),
                                              L.stream_ptr(dev)), "plx_normalize_points")
    return out


def _dims_strides(grid):
    dims = (C.c_int32 * 3)(*[int(s) for s in grid.shape[:3]])
    strides = (C.c_int64 * 4)(*[int(s) for s in grid.stride()])
    return dims, strides


class _GatherNearest(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ns, grid):
        dev = L.require_cuda(ns, grid)
        ns = ns.contiguous().float()
        m = ns.shape[0]
        vals = torch.empty((m, 4), dtype=torch.float32, device=dev)
        inb = torch.empty((m,), dtype=torch.uint8, device=dev)
        dims, strides = _dims_strides(grid)
        with torch.cuda.device(dev):
            L.check(L.load().plx_gather_nearest(ns.data_ptr(), m, grid.data_ptr(), dims, strides, vals.data_ptr(),
                                                inb.data_ptr(), None, L.stream_ptr(dev)), "plx_gather_nearest")
        ctx.save_for_backward(ns)
        ctx.grid_shape = tuple(grid.shape)
        mask = inb.view(torch.bool)
        ctx.mark_non_differentiable(mask)
        return vals, mask

    @staticmethod
    def backward(ctx, grad_vals, _):
        (ns,) = ctx.saved_tensors
        dev = ns.device
        gg = torch.zeros(ctx.grid_shape, dtype=torch.float32, device=dev)
        gv = grad_vals.contiguous().float()
        dims = (C.c_int32 * 3)(*ctx.grid_shape[:3])
        with torch.cuda.device(dev):
            L.check(L.load().plx_gather_nearest_bwd(ns.data_ptr(), ns.shape[0], gv.data_ptr(), dims, gg.data_ptr(),
                                                    L.stream_ptr(dev)), "plx_gather_nearest_bwd")
        return None, gg


def gather_nearest(ns, grid):
    """`get_nearest_voxels` (src/grid_functions.py:103-114): values at wrapped indices (M,4) + in-bounds mask (M,) bool."""
    if grid.dtype != torch.float32 or grid.dim() != 4 or grid.shape[3] != 4:
        raise L.PlxError(f"grid must be float32 (X,Y,Z,4), got {grid.dtype} {tuple(grid.shape)}")
    return _GatherNearest.apply(ns, grid)


class _Trilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ns, grid, masked):
        dev = L.require_cuda(ns, grid)
        ns = ns.contiguous().float()
        m = ns.shape[0]
        vals = torch.empty((m, 4), dtype=torch.float32, device=dev)
        inb = torch.empty((m,), dtype=torch.uint8, device=dev)
        dims, strides = _dims_strides(grid)
        with torch.cuda.device(dev):
            L.check(L.load().plx_trilinear_fwd(ns.data_ptr(), m, grid.data_ptr(), dims, strides, int(masked),
                                               vals.data_ptr(), inb.data_ptr(), L.stream_ptr(dev)), "plx_trilinear_fwd")
        ctx.save_for_backward(ns)
        ctx.grid_shape, ctx.masked = tuple(grid.shape), int(masked)
        mask = inb.view(torch.bool)
        ctx.mark_non_differentiable(mask)
        return vals, mask

    @staticmethod
    def backward(ctx, grad_vals, _):
        (ns,) = ctx.saved_tensors
        dev = ns.device
        gg = torch.zeros(ctx.grid_shape, dtype=torch.float32, device=dev)
        gv = grad_vals.contiguous().float()
        dims = (C.c_int32 * 3)(*ctx.grid_shape[:3])
        with torch.cuda.device(dev):
            L.check(L.load().plx_trilinear_bwd(ns.data_ptr(), ns.shape[0], gv.data_ptr(), dims, ctx.masked, gg.data_ptr(),
                                               L.stream_ptr(dev)), "plx_trilinear_bwd")
        return None, gg, None


def trilinear_lookup(ns, grid, masked=True):
    """Trilinear lookup at normalised coordinates (M,3): 8 periodically wrapped corners, nested lerps
    (src/grid_functions.py:7-44, :220-246, :66-79); `masked` multiplies by the float in-bounds test."""
    if grid.dtype != torch.float32 or grid.dim() != 4 or grid.shape[3] != 4:
        raise L.PlxError(f"grid must be float32 (X,Y,Z,4), got {grid.dtype} {tuple(grid.shape)}")
    return _Trilinear.apply(ns, grid, masked)


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, samples):
        dev = L.require_cuda(samples)
        s = samples.contiguous().float()
        lead, S = s.shape[:-2], s.shape[-2]
        n = int(np.prod(lead)) if len(lead) else 1
        out = torch.empty(lead + (4,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.load().plx_composite_fwd(s.data_ptr(), n, S, out.data_ptr(), L.stream_ptr(dev)), "plx_composite_fwd")
        ctx.save_for_backward(s)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (s,) = ctx.saved_tensors
        dev = s.device
        lead, S = s.shape[:-2], s.shape[-2]
        n = int(np.prod(lead)) if len(lead) else 1
        go = grad_out.contiguous().float()
        gs = torch.zeros_like(s)
        with torch.cuda.device(dev):
            L.check(L.load().plx_composite_bwd(s.data_ptr(), n, S, go.data_ptr(), gs.data_ptr(), L.stream_ptr(dev)),
                    "plx_composite_bwd")
        return gs


class _AvgPool3dGrid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grid, kernel, stride):
        dev = L.require_cuda(grid)
        g = grid.contiguous().float()
        X, Y, Z, ch = g.shape
        if ch != 4:
            raise L.PlxError("average pooling kernel handles (X,Y,Z,4) grids")
        O = [(d - kernel) // stride + 1 for d in (X, Y, Z)]
        if min(O) < 1:
            raise RuntimeError(f"pooling window {kernel} larger than the grid {(X, Y, Z)}")      # F.avg_pool3d raises too
        out = torch.empty((O[0], O[1], O[2], 4), dtype=torch.float32, device=dev)
        t1 = torch.empty((X * Y * O[2] * 4,), dtype=torch.float32, device=dev)
        t2 = torch.empty((X * O[1] * O[2] * 4,), dtype=torch.float32, device=dev)
        dims = (C.c_int32 * 3)(X, Y, Z)
        with torch.cuda.device(dev):
            L.check(L.load().plx_avgpool3d_fwd(g.data_ptr(), dims, kernel, stride, t1.data_ptr(), t2.data_ptr(), out.data_ptr(),
                                               L.stream_ptr(dev)), "plx_avgpool3d_fwd")
        ctx.meta = (X, Y, Z, kernel, stride, O)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        X, Y, Z, kernel, stride, O = ctx.meta
        dev = grad_out.device
        go = grad_out.contiguous().float()
        gin = torch.empty((X, Y, Z, 4), dtype=torch.float32, device=dev)
        t1 = torch.empty((X * Y * O[2] * 4,), dtype=torch.float32, device=dev)
        t2 = torch.empty((X * O[1] * O[2] * 4,), dtype=torch.float32, device=dev)
        dims = (C.c_int32 * 3)(X, Y, Z)
        with torch.cuda.device(dev):
            L.check(L.load().plx_avgpool3d_bwd(go.data_ptr(), dims, kernel, stride, t2.data_ptr(), t1.data_ptr(), gin.data_ptr(),
                                               L.stream_ptr(dev)), "plx_avgpool3d_bwd")
        return gin, None, None


def avgpool3d_grid_backward_into(grad_out, dims, kernel, stride, grad_in):
    """Autograd of `avgpool3d_grid` written straight into an existing full-resolution gradient buffer (overwritten)."""
    dev = L.require_cuda(grad_out, grad_in)
    X, Y, Z = (int(d) for d in dims[:3])
    O = [(d - kernel) // stride + 1 for d in (X, Y, Z)]
    go = grad_out.contiguous().float()
    if tuple(go.shape) != (O[0], O[1], O[2], 4) or tuple(grad_in.shape) != (X, Y, Z, 4) or not grad_in.is_contiguous():
        raise L.PlxError("avgpool3d_grid_backward_into: shape mismatch")
    t1 = torch.empty((X * Y * O[2] * 4,), dtype=torch.float32, device=dev)
    t2 = torch.empty((X * O[1] * O[2] * 4,), dtype=torch.float32, device=dev)
    cd = (C.c_int32 * 3)(X, Y, Z)
    with torch.cuda.device(dev):
        L.check(L.load().plx_avgpool3d_bwd(go.data_ptr(), cd, int(kernel), int(stride), t2.data_ptr(), t1.data_ptr(),
                                           grad_in.data_ptr(), L.stream_ptr(dev)), "plx_avgpool3d_bwd")
    return grad_in


def avgpool3d_grid(grid, kernel, stride=None):
    """`average_pool3d_grid` (src/grid_functions.py:173-181): cubic window, stride (default = window), no padding."""
    return _AvgPool3dGrid.apply(grid, int(kernel), int(stride if stride is not None else kernel))


def tv_loss_(grid, tv, grad=None):
    """tv * tv_loss(grid) (scripts/train.py:44-65, :163-168) as a (1,) device tensor; with `grad` (contiguous, grid-shaped)
    its gradient is ADDED to it in place — two dense stencil passes instead of ~15 ATen ops and their autograd."""
    dev = L.require_cuda(grid, grad)
    if not grid.is_contiguous() or grid.dtype != torch.float32 or grid.dim() != 4 or grid.shape[3] != 4:
        raise L.PlxError("tv_loss_ needs a contiguous float32 (X,Y,Z,4) grid")
    if grad is not None and (not grad.is_contiguous() or grad.shape != grid.shape or grad.dtype != torch.float32):
        raise L.PlxError("grad must be a contiguous float32 tensor of the grid's shape")
    out = torch.zeros((1,), dtype=torch.float32, device=dev)
    scratch = torch.zeros((1,), dtype=torch.float64, device=dev)
    dims = (C.c_int32 * 3)(*[int(s) for s in grid.shape[:3]])
    with torch.cuda.device(dev):
        L.check(L.load().plx_tv_loss(grid.data_ptr(), dims, float(tv), L.ptr(grad), scratch.data_ptr(), out.data_ptr(),
                                     L.stream_ptr(dev)), "plx_tv_loss")
    return out


def composite(samples):
    """`compute_alpha_weighted_pixels` (src/ray_sampling.py:172-192): (..., S, 4) -> (..., 4)."""
    if samples.shape[-1] != 4:
        raise L.PlxError(f"samples must end in 4 channels, got {tuple(samples.shape)}")
    return _Composite.apply(samples)
