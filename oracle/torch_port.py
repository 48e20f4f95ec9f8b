"""torch-CPU port of the reference's step — TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's hot path *is* a sequence of eager ATen ops plus autograd and
`torch.optim.Adam`, multi-threaded through ATen's intra-op
pool.  This file restates that op sequence (same ops, same order, same
materialised temporaries) so that `bench.py`'s `cpu_baseline` and
`--impl reference` legs can time "the reference's CPU path" on a GPU box where
`/root/reference` does not exist, and so that tests can cross-check the numpy
oracle's hand-derived backward against autograd.  It is validated against the
real reference in `oracle/validate_against_reference.py` (bit-exact forward
tensors, identical autograd graph => identical gradients).

Nothing under `plenoxels_b200/` imports this.  Citations are reference `file:line`.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def cell_centres(dims, points_distance, device="cpu"):
    """(X,Y,Z,3) fp32 world coordinates of the cell centres — src/grid_functions.py:198-212."""
    axes = [torch.arange(s, device=device) - np.ceil(s / 2) + 1 for s in dims]
    cx, cy, cz = torch.meshgrid(*axes, indexing="ij")
    return torch.stack([cx * points_distance, cy * points_distance, cz * points_distance], dim=-1)


def rays_from_uv(imgs, poses, fov, uv):
    """dirs (N,3), targets (N,4) for given uv (C,R,2) — src/ray_sampling.py:212-264 with the RNG lifted out."""
    C, R = uv.shape[:2]
    ax, ay, az = poses[:, :3, 0], poses[:, :3, 1], -poses[:, :3, 2]
    aspect = ax.norm(dim=1) / ay.norm(dim=1)
    pix = uv.clone()
    ang = uv.clone()
    ang[:, :, 0] = fov * (ang[:, :, 0] - 0.5)
    ang[:, :, 1] = -((fov * (1 / aspect)).unsqueeze(1) * (ang[:, :, 1] - 0.5))
    pix[:, :, 0] = (imgs.shape[1] * pix[:, :, 0]).round().clamp(max=imgs.shape[1] - 1)
    pix[:, :, 1] = (imgs.shape[2] * pix[:, :, 1]).round().clamp(max=imgs.shape[2] - 1)
    pix = pix.to(torch.long).reshape(C * R, 2)
    cam = torch.repeat_interleave(torch.arange(C, device=uv.device), R, 0)
    targets = imgs[cam, pix[:, 1], pix[:, 0]]
    u = ang[:, :, 0:1].expand(C, R, 3)
    v = ang[:, :, 1:2].expand(C, R, 3)
    d = u * ax.unsqueeze(1).expand(C, R, 3) + v * ay.unsqueeze(1).expand(C, R, 3) + az.unsqueeze(1).expand(C, R, 3)
    d = d / d.norm(dim=2).unsqueeze(-1)
    return d.reshape(C * R, 3), targets


def place_samples(cam_pos, dirs, rays_per_cam, num_samples, delta_step):
    """(M,3) sample positions, camera-major — src/ray_sampling.py:161-167."""
    n_rays = dirs.shape[0]
    t = delta_step * torch.arange(num_samples + 1, device=dirs.device)[1:].repeat(n_rays).unsqueeze(1)
    o = torch.repeat_interleave(cam_pos, num_samples * rays_per_cam, 0)
    d = torch.repeat_interleave(dirs, num_samples, 0)
    return o + d * t


def nearest_lookup(ns, grid):
    """`get_nearest_voxels` — src/grid_functions.py:103-114 (round, in-bounds mask, in-place periodic wrap, gather)."""
    idx = torch.round(ns).to(torch.long)
    X, Y, Z, _ = grid.shape
    i0, i1, i2 = idx[:, 0], idx[:, 1], idx[:, 2]
    inb = ((i0 < X) & (i0 >= 0)) & ((i1 < Y) & (i1 >= 0)) & ((i2 < Z) & (i2 >= 0))
    i0 %= X
    i1 %= Y
    i2 %= Z
    return grid[i0, i1, i2], inb


def trilinear_lookup(ns, grid):
    """Composition of the reference's trilinear pieces: float-coordinate mask,
    8 wrapped corners (src/grid_functions.py:230-244, :75-77), nested lerps (:29-42), mask multiply."""
    X, Y, Z, _ = grid.shape
    inb = ((ns[:, 0] < X) & (ns[:, 0] >= 0)) & ((ns[:, 1] < Y) & (ns[:, 1] >= 0)) & ((ns[:, 2] < Z) & (ns[:, 2] >= 0))
    hi, lo = torch.ceil(ns).to(torch.long), torch.floor(ns).to(torch.long)
    dims = torch.tensor([X, Y, Z], device=ns.device)
    hi, lo = hi % dims, lo % dims
    f = torch.frac(ns)

    def corner(cx, cy, cz):
        return grid[(hi if cx else lo)[:, 0], (hi if cy else lo)[:, 1], (hi if cz else lo)[:, 2]]

    fx, fy, fz = f[:, 0:1], f[:, 1:2], f[:, 2:3]
    x = {(cy, cz): corner(1, cy, cz) * fx + corner(0, cy, cz) * (1 - fx) for cy in (0, 1) for cz in (0, 1)}
    y = {cz: x[(1, cz)] * fy + x[(0, cz)] * (1 - fy) for cz in (0, 1)}
    out = y[1] * fz + y[0] * (1 - fz)
    return out * inb.unsqueeze(-1), inb


def composite(samples):
    """(C,R,S,4) -> (C,R,4) — src/ray_sampling.py:181-191 (cat, rsub, cumprod, mul, two sums, cat)."""
    alpha = samples[..., -1]
    shifted = torch.cat([torch.zeros_like(alpha[..., :1]), alpha], dim=2)
    trans = (1 - shifted[..., :-1]).cumprod(dim=2)
    w = (alpha * trans).unsqueeze(-1)
    return torch.cat([(samples[..., :-1] * w).sum(2), w.sum(dim=2)], dim=2)


class ReferenceStep:
    """The step of scripts/train.py:130-184 (tv = beta = 0, full resolution) on torch-CPU."""

    def __init__(self, grid, points_distance, poses, fov, imgs, rays_per_cam, num_samples, delta_step, lr,
                 mode="nearest", device="cpu"):
        self.device = device
        self.grid = grid.detach().clone().to(device).requires_grad_(True)
        self.grid_abs_grad = torch.zeros_like(self.grid)
        self.opt = torch.optim.Adam([self.grid], lr=lr)                       # scripts/train.py:89
        self.pd = points_distance
        self.centres = cell_centres(self.grid.shape[:3], points_distance, device).reshape(-1, 3)
        self.poses, self.fov, self.imgs = poses.to(device), fov, imgs.to(device)
        self.R, self.S, self.delta = rays_per_cam, num_samples, delta_step
        self.mode = mode

    def forward(self, uv, cams=None):
        poses = self.poses if cams is None else self.poses[cams]
        imgs = self.imgs if cams is None else self.imgs[cams]
        dirs, targets = rays_from_uv(imgs, poses, self.fov, uv.clone())
        pos = place_samples(poses[:, :3, 3], dirs, uv.shape[1], self.S, self.delta)
        ns = (pos - self.centres.min(0)[0]) / self.pd                          # src/ray_sampling.py:13
        clipped = self.grid.clip(0, 1)                                         # scripts/train.py:146
        if self.mode == "nearest":
            vals, inb = nearest_lookup(ns, clipped)
            vals = vals * inb.unsqueeze(-1)                                    # :147
        else:
            vals, inb = trilinear_lookup(ns, clipped)
        pix = composite(vals.reshape(poses.shape[0], uv.shape[1], self.S, 4))
        return pix.reshape(-1, 4), targets

    def step(self, uv, cams=None):
        pix, targets = self.forward(uv, cams)
        loss = F.mse_loss(pix, targets)                                        # :156
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()                                                        # :180-182
        self.grid_abs_grad += torch.abs(self.grid.grad)                        # :184
        return loss.detach()
