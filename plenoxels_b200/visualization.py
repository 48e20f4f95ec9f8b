"""Drop-in for the hot-path consumers in the reference's `src/visualization.py`.

`visulize_3d_in_2d` (ray-marched inference, :111-154) runs on the fused forward kernel.  `visulize_3d_in_2d_fast`
(:157-232) is the reference's CPU point-splat visual aid, kept on the CPU (out of scope).  The
matplotlib / plotly viewers are GUI-only; they import their toolkit lazily so this module loads without them.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .grid_functions import coords_origin


def visulize_3d_in_2d(grid_cells_data, transform_matrices, camera_angle_x, imgs, grid_indices, do_threshold,
                      transparency_threshold, number_of_rays, num_samples, device="cuda"):
    """Render one camera by ray marching -> (res,res,4) uint8 — src/visualization.py:111-154."""
    grid = grid_cells_data["grid"].detach().to(device).clip(0.0, 1.0)
    if do_threshold:
        alphas = grid[..., -1]
        grid[..., -1][alphas < transparency_threshold] = 0.0
    pd = grid_cells_data["param"]["points_distance"]
    delta = grid_cells_data["param"]["delta_step"]
    n_side = int(np.round(np.sqrt(number_of_rays)))
    poses = transform_matrices.to(device).float()
    dirs, _ = ops.generate_rays(None, poses, camera_angle_x, uv=None, rays_per_cam=n_side * n_side, want_targets=False)
    pix = ops.render_rays(grid, poses[:, :3, 3], dirs, num_samples, delta, coords_origin(grid_indices), pd, clamp=False,
                          rays_per_origin=n_side * n_side, coherent=True)
    res = int(np.sqrt(number_of_rays))
    img = (pix.cpu().numpy() * 255).round().clip(0, 255).astype(np.uint8).reshape(res, res, 4)
    return np.transpose(img, (1, 0, 2))


def visulize_3d_in_2d_fast(grid, points_distance, transform_matrix, camera_angle_x, size_y):
    """CPU painter's-algorithm splat of voxels with alpha > 0.1 -> (xs, ys, 3) float image — src/visualization.py:157-232."""
    grid = grid.detach().cpu()
    T = transform_matrix.detach().cpu()
    pos, ax_x, ax_y = T[:3, 3], T[:3, 0], T[:3, 1]
    aspect = ax_x.norm(dim=0) / ax_y.norm(dim=0)
    keep = grid[..., 3] > 0.1
    pts = np.argwhere(keep)
    colors = grid[keep]
    half = (torch.tensor(grid.shape[:3]) / 2).unsqueeze(1).ceil()
    pts = (((pts - half) + 1) * points_distance).T
    rel = pts - pos
    norms = rel.norm(dim=1) * ax_x.norm()
    ang_x = torch.matmul(rel, ax_x) / norms
    ang_y = torch.matmul(rel, ax_y) / norms
    xn = 0.5 + ang_y / -camera_angle_x
    yn = 0.5 + ang_x / (camera_angle_x / aspect)
    ok = (xn < 1.0).logical_and(xn >= 0.0).logical_and(yn < 1.0).logical_and(yn >= 0.0)
    xn, yn, colors, rel = xn[ok], yn[ok], colors[ok], rel[ok]
    ys = size_y
    xs = int(ys * aspect)
    xx = (xs * xn).round().clamp(min=0, max=xs - 1).type(torch.long)
    yy = (ys * yn).round().clamp(min=0, max=ys - 1).type(torch.long)
    order = torch.argsort(-rel.norm(dim=1))
    img = np.ones([xs, ys, 3])
    img[xx[order], yy[order]] = colors[order][:, :3]
    return img


def visualize_rays_3d(ray_directions, camera_positions, red_dots=None):
    """plotly quiver of rays (GUI only) — src/visualization.py:70-107."""
    import plotly.graph_objects as go
    import plotly.io as pio
    fig = go.Figure()
    for i in range(ray_directions.shape[0]):
        p, d = camera_positions[i], ray_directions[i]
        fig.add_trace(go.Scatter3d(x=[float(p[0]), float(p[0] + d[0])], y=[float(p[1]), float(p[1] + d[1])],
                                   z=[float(p[2]), float(p[2] + d[2])], mode="lines"))
    if red_dots is not None:
        fig.add_trace(go.Scatter3d(x=red_dots[:, 0], y=red_dots[:, 1], z=red_dots[:, 2], mode="markers",
                                   marker=dict(size=2, color="red")))
    pio.write_html(fig, file="visualize_rays_3d.html", auto_open=False)
