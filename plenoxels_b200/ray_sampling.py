"""Drop-in for the reference's `src/ray_sampling.py`: same names, arguments and return layouts, CUDA kernels underneath.

`render_cameras` is the fused extra (one kernel for the whole sample -> normalise -> lookup -> composite sequence);
the reference-named functions below keep the materialised-tensor semantics so the unmodified scripts run.
"""
from __future__ import annotations

import numpy as np
import torch

import os

from . import _lib as L
from . import lazy, ops
from .grid_functions import coords_origin


def _lazy_enabled() -> bool:
    """PLX_LAZY=0 turns the lazy-fusion bridge off (every function then materialises its result)."""
    return os.environ.get("PLX_LAZY", "1") != "0"


def normalize_samples_for_indecies(grid_indices, samples_interval, points_distance):
    """(samples - grid_indices.min(0)[0]) / points_distance — src/ray_sampling.py:12-13.
    The minimum is taken from the grid's metadata when `grid_indices` came from `generate_grid` (no G^3 reduction)."""
    L.require_cuda(samples_interval)
    if isinstance(samples_interval, lazy.LazyTensor) and samples_interval._kind == "samples" and samples_interval._real is None:
        return lazy.lazy_normalized(samples_interval, coords_origin(grid_indices), points_distance)
    if isinstance(samples_interval, lazy.LazyTensor):
        samples_interval = samples_interval.materialize()
    return ops.normalize_points(samples_interval, coords_origin(grid_indices), points_distance)


def generate_rays_batched(imgs, number_of_rays, transform_matricies, camera_angle_x, even_spread=False, device='cuda'):
    """Ray directions (C*R,3) and their target pixels (C*R,4) — src/ray_sampling.py:195-264.
    Random uv are drawn with `torch.rand` on `device` exactly where the reference draws them (:227)."""
    L.require_cuda(imgs, transform_matricies)
    if even_spread:
        n_side = int(np.round(np.sqrt(number_of_rays)))
        return ops.generate_rays(imgs, transform_matricies, camera_angle_x, uv=None, rays_per_cam=n_side * n_side)
    uv = torch.rand(transform_matricies.shape[0], number_of_rays, 2, device=device)
    return ops.generate_rays(imgs, transform_matricies, camera_angle_x, uv=uv)


def sample_camera_rays_batched(transform_matrices, camera_angle_x, imgs, number_of_rays, num_samples, delta_step,
                               even_spread, camera_ray, device='cuda'):
    """samples (C*R*S,3), pixels_to_rays (C*R,4), camera_positions (C,3), ray directions (C*R,3) —
    src/ray_sampling.py:128-169."""
    L.require_cuda(transform_matrices, imgs)
    if even_spread:
        number_of_rays = int(np.round(np.sqrt(number_of_rays)) ** 2)
    if camera_ray:
        # the reference's camera_ray branch builds a (1,C,3) direction tensor that cannot broadcast against the
        # (C*R*S,3) positions (src/ray_sampling.py:151 vs :166-167): it raises there, so it raises here
        raise RuntimeError("camera_ray=True is shape-inconsistent in the reference (src/ray_sampling.py:151) and unsupported")
    dirs, pixels_to_rays = generate_rays_batched(imgs, number_of_rays, transform_matrices, camera_angle_x,
                                                 even_spread=even_spread, device=device)
    camera_positions = transform_matrices[:, :3, 3]
    if _lazy_enabled():
        # a handle, not (C*R*S,3) floats: the positions only exist if someone other than the fused march asks for them
        samples_interval = lazy.lazy_samples(camera_positions.float(), dirs, number_of_rays, num_samples, delta_step)
    else:
        samples_interval = ops.sample_points(camera_positions.float(), dirs, num_samples, delta_step,
                                             rays_per_origin=number_of_rays)
    return samples_interval, pixels_to_rays, camera_positions, dirs


def compute_alpha_weighted_pixels(samples):
    """(C,R,S,4) -> (C,R,4) front-to-back compositing — src/ray_sampling.py:172-192 (warp-scan kernel, differentiable)."""
    L.require_cuda(samples)
    if isinstance(samples, lazy.LazyTensor):
        fused = lazy.fused_composite(samples)          # masked (C,R,S,4) lookup of lazy samples -> one fused march (K1/K2)
        if fused is not None:
            return fused
        samples = samples.materialize()
    return ops.composite(samples)


def render_cameras(grid, grid_indices, points_distance, transform_matrices, camera_angle_x, imgs, number_of_rays,
                   num_samples, delta_step, even_spread=False, mode="nearest", clamp=True, uv=None, return_depth=False,
                   device=None):
    """Fused equivalent of scripts/train.py:130-153 / src/visualization.py:125-146: returns (pixels (C*R,4), targets
    (C*R,4)[, depth]) without materialising samples.  Differentiable w.r.t. `grid`."""
    L.require_cuda(grid, transform_matrices, imgs)
    if even_spread:
        n_side = int(np.round(np.sqrt(number_of_rays)))
        number_of_rays = n_side * n_side
        dirs, targets = ops.generate_rays(imgs, transform_matrices, camera_angle_x, uv=None, rays_per_cam=number_of_rays)
    else:
        if uv is None:
            uv = torch.rand(transform_matrices.shape[0], number_of_rays, 2, device=grid.device)
        dirs, targets = ops.generate_rays(imgs, transform_matrices, camera_angle_x, uv=uv)
    gmin = coords_origin(grid_indices) if isinstance(grid_indices, torch.Tensor) else tuple(grid_indices)
    out = ops.render_rays(grid, transform_matrices[:, :3, 3].float(), dirs, num_samples, delta_step, gmin, points_distance,
                          mode=mode, clamp=clamp, rays_per_origin=number_of_rays, return_depth=return_depth)
    if return_depth:
        return out[0], targets, out[1]
    return out, targets
