"""Drop-in for the hot-path consumers in the reference's `src/visualization.py`.

`visulize_3d_in_2d` (ray-marched inference, :111-154) runs on the fused forward kernel.  `visulize_3d_in_2d_fast`
(:157-232) is the reference's CPU point-splat visual aid, kept on the CPU (out of scope).  The
matplotlib / plotly viewers are GUI-only; they import their toolkit lazily so this module loads without them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import ops
from .grid_functions import coords_origin


def visulize_3d_in_2d(grid_cells_data, transform_matrices, camera_angle_x, imgs, grid_indices, do_threshold,
                      transparency_threshold, number_of_rays, num_samples, device="cpu"):
    """Render one camera by ray marching -> (res,res,4) uint8 — src/visualization.py:111-154.  Same signature and default
    as the reference; the march runs on sm_100a only, so a non-CUDA `device` raises (there is no CPU fallback).  The whole
    post-processing (x255, round, clip, uint8, transpose, :150-154) is the march kernel's epilogue; the host only copies the
    finished image."""
    if torch.device(device).type != "cuda":
        raise L.PlxError(f"visulize_3d_in_2d runs on CUDA (sm_100a) only: pass device='cuda' (got {device!r}); there is no CPU fallback")
    grid = grid_cells_data["grid"].detach().to(device).clip(0.0, 1.0)
    if do_threshold:
        alphas = grid[..., -1]
        grid[..., -1][alphas < transparency_threshold] = 0.0
    pd = grid_cells_data["param"]["points_distance"]
    delta = grid_cells_data["param"]["delta_step"]
    n_side = int(np.round(np.sqrt(number_of_rays)))
    res = int(np.sqrt(number_of_rays))
    poses = transform_matrices.to(device).float()
    if n_side == res and poses.shape[0] == 1:
        return ops.render_image_u8(grid, poses, camera_angle_x, res, num_samples, delta, coords_origin(grid_indices), pd).cpu().numpy()
    # a ray count that is not a perfect square (the reference's reshape only works when it is) or several cameras: float pixels
    dirs, _ = ops.generate_rays(None, poses, camera_angle_x, uv=None, rays_per_cam=n_side * n_side, want_targets=False)
    pix = ops.render_rays(grid, poses[:, :3, 3], dirs, num_samples, delta, coords_origin(grid_indices), pd, clamp=False,
                          rays_per_origin=n_side * n_side, coherent=True)
    img = (pix.cpu().numpy() * 255).round().clip(0, 255).astype(np.uint8).reshape(res, res, 4)
    return np.transpose(img, (1, 0, 2))


def visulize_3d_in_2d_fast(grid, points_distance, transform_matrix, camera_angle_x, size_y):
    """Painter's-order splat of the voxels with alpha > 0.1 -> (xs, ys, 3) float64 image — src/visualization.py:157-232, the
    preview scripts/compare_inference_to_image.py:58 calls.  Two kernels on the device that holds `grid` (plx_splat_view: a
    64-bit atomicMin per voxel on (distance, cell), then a resolve pass) instead of argwhere + argsort + fancy indexing on the
    CPU; the nearest voxel wins each pixel exactly as the reference's far-to-near assignment order leaves it."""
    dev = L.require_cuda(grid)
    g = grid.detach()
    if g.dim() != 4 or g.shape[3] != 4 or g.dtype != torch.float32:
        raise L.PlxError(f"grid must be float32 (X,Y,Z,4), got {g.dtype} {tuple(g.shape)}")
    g = g.contiguous()
    T = transform_matrix.detach().float().cpu().contiguous()                       # 16 floats; the reference moves them too (:161)
    aspect = T[:3, 0].norm(dim=0) / T[:3, 1].norm(dim=0)                           # :171
    ys = int(size_y)
    xs = int(ys * aspect)                                                          # :216
    zbuf = torch.empty((xs * ys,), dtype=torch.int64, device=dev)
    img = torch.empty((xs, ys, 3), dtype=torch.float32, device=dev)
    dims = (C.c_int32 * 3)(*[int(d) for d in g.shape[:3]])
    pose = (C.c_float * 16)(*[float(v) for v in T.reshape(-1).tolist()])
    with torch.cuda.device(dev):
        L.check(L.load().plx_splat_view(g.data_ptr(), dims, float(points_distance), pose, float(camera_angle_x), xs, ys,
                                        zbuf.data_ptr(), img.data_ptr(), L.stream_ptr(dev)), "plx_splat_view")
    return img.cpu().numpy().astype(np.float64)


def visualize_rays_3d(ray_directions, camera_positions, red=None, green=None, orange=None):
    """plotly quiver of rays with optional marker clouds (GUI only) — src/visualization.py:70-107."""
    import plotly.graph_objects as go
    import plotly.io as pio
    fig = go.Figure()
    for i in range(ray_directions.shape[0]):
        p, d = camera_positions[i], ray_directions[i]
        fig.add_trace(go.Scatter3d(x=[float(p[0]), float(p[0] + d[0])], y=[float(p[1]), float(p[1] + d[1])],
                                   z=[float(p[2]), float(p[2] + d[2])], mode="lines"))
    for dots, colour in ((red, "red"), (green, "green"), (orange, "orange")):
        if dots is not None:
            fig.add_trace(go.Scatter3d(x=dots[:, 0], y=dots[:, 1], z=dots[:, 2], mode="markers", marker=dict(size=2, color=colour)))
    pio.write_html(fig, file="visualize_rays_3d.html", auto_open=False)
