"""Synthetic scenes for parity tests and benchmarks.

Everything here is host-side (numpy / torch-CPU) and seeded, so the GPU path,
the oracle and the golden fixtures all see the same bytes.  Nothing in this
module touches the CUDA library or the oracle.

The shapes follow the NeRF-synthetic data the reference trains on
(`scripts/train.py:73-75`, `src/data_processing.py:51-60`): camera-to-world
4x4 poses whose columns are (right, up, backward, position), one horizontal
field of view for all cameras, RGBA images in [0, 1].
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

CAMERA_ANGLE_X = 0.6911112070083618      # NeRF-synthetic "camera_angle_x"
CAMERA_RADIUS = 4.031128874              # NeRF-synthetic camera distance
GRID_EXTENT = 3.2                        # world span of the grid (256 * 0.0125, scripts/main.py:24-25)


def lookat_poses(n_cams: int, radius: float = CAMERA_RADIUS, dtype=torch.float32) -> torch.Tensor:
    """(C,4,4) camera-to-world matrices on the upper hemisphere.

    Azimuth follows a golden-ratio spiral, elevation is spread uniformly over
    [0 deg, 80 deg].  The rotation part is orthonormal, so the reference's
    aspect ratio ||X|| / ||Y|| (`src/ray_sampling.py:218`) is 1 up to rounding.
    """
    inv_phi = 2.0 / (1.0 + math.sqrt(5.0))
    poses = np.zeros((n_cams, 4, 4), dtype=np.float64)
    for i in range(n_cams):
        az = 2.0 * math.pi * i * inv_phi
        el = math.radians(80.0) * ((i + 0.5) / n_cams)
        pos = radius * np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])
        backward = pos / np.linalg.norm(pos)            # camera looks along -Z at the origin
        world_up = np.array([0.0, 0.0, 1.0])
        right = np.cross(world_up, backward)
        right /= np.linalg.norm(right)
        up = np.cross(backward, right)
        poses[i, :3, 0] = right
        poses[i, :3, 1] = up
        poses[i, :3, 2] = backward
        poses[i, :3, 3] = pos
        poses[i, 3, 3] = 1.0
    return torch.from_numpy(poses).to(dtype)


def ball_grid(G: int, seed: int = 0, occupancy_radius: float = 0.288) -> torch.Tensor:
    """(G,G,G,4) fp32 grid, ~10 % of the cells occupied by one centred ball.

    Occupied cells: RGB ~ U(0,1), alpha ~ U(0.02,0.6); 5 % of them get a raw
    value outside [0,1] in one channel (exercises the clip pass-mask,
    `scripts/train.py:146`), 1 % get alpha >= 1 (exercises T == 0).
    Empty cells are exactly 0.
    """
    g = torch.Generator().manual_seed(seed)
    ax = torch.arange(G, dtype=torch.float32) - (G - 1) / 2.0
    r2 = ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2
    occ = r2 <= (occupancy_radius * G) ** 2
    n_occ = int(occ.sum())
    vals = torch.empty(n_occ, 4)
    vals[:, :3] = torch.rand(n_occ, 3, generator=g)
    vals[:, 3] = 0.02 + 0.58 * torch.rand(n_occ, generator=g)
    # 5 %: one channel pushed outside [0,1]
    sel = torch.rand(n_occ, generator=g) < 0.05
    ch = torch.randint(0, 4, (n_occ,), generator=g)
    hi = torch.rand(n_occ, generator=g) < 0.5
    out = torch.where(hi, 1.0 + 0.2 * torch.rand(n_occ, generator=g) + 1e-3,
                      -0.2 * torch.rand(n_occ, generator=g) - 1e-3)
    rows = torch.nonzero(sel).squeeze(1)
    vals[rows, ch[rows]] = out[rows]
    # 1 %: opaque cells (alpha >= 1 before the clip)
    opq = torch.rand(n_occ, generator=g) < 0.01
    vals[opq, 3] = 1.0 + 0.1 * torch.rand(int(opq.sum()), generator=g)
    grid = torch.zeros(G, G, G, 4)
    grid[occ] = vals
    return grid


def dense_grid(G: int, seed: int = 0) -> torch.Tensor:
    """(G,G,G,4) fp32, every value ~ U(-0.2, 1.2) ("dense" configs of §8d)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(G, G, G, 4, generator=g) * 1.4 - 0.2


def soft_grid(G: int, seed: int = 0) -> torch.Tensor:
    """(G,G,G,4) fp32 translucent everywhere: alpha ~ U(0,0.05), a few cells outside [0,1].

    Rays never saturate, so every in-bounds sample carries gradient — the
    regime of the first training steps (the reference starts from zeros,
    `scripts/train.py:85-87`).
    """
    g = torch.Generator().manual_seed(seed)
    grid = torch.rand(G, G, G, 4, generator=g)
    grid[..., 3] *= 0.05
    flip = torch.rand(G, G, G, generator=g) < 0.02
    grid[..., 3][flip] = -0.01
    return grid


def random_images(C: int, H: int, W: int, seed: int = 1) -> torch.Tensor:
    """(C,H,W,4) fp32 ~ U(0,1) targets (content is irrelevant for throughput / gradient parity)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(C, H, W, 4, generator=g)


def random_uv(C: int, R: int, seed: int = 2) -> torch.Tensor:
    """(C,R,2) fp32 ~ U[0,1): the per-step random ray coordinates of `src/ray_sampling.py:227`."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(C, R, 2, generator=g)


@dataclass
class Scene:
    """One workload: grid + cameras + images + marching parameters."""
    name: str
    G: int
    points_distance: float
    grid: torch.Tensor            # (G,G,G,4) fp32
    poses: torch.Tensor           # (C,4,4) fp32
    fov: float
    imgs: torch.Tensor            # (C,H,W,4) fp32
    rays_per_cam: int
    num_samples: int
    delta_step: float
    lr: float = 0.0075            # scripts/train.py:226

    @property
    def n_rays(self) -> int:
        return self.poses.shape[0] * self.rays_per_cam


def make_scene(name: str, *, H: int | None = None, n_cams: int | None = None, kind: str | None = None) -> Scene:
    """The BASELINE.json configs as `Scene`s.

    c1: G=64,  1 view 64x64,    R=4096 (even spread), S=64,  delta=6/S, dense grid
    c2: G=128, 100 views 800^2, R=128,               S=600, delta=0.0125, pd=0.025
    c3: G=256, 32 views 800^2,  R=128 (4096 rays),   S=256, delta=6/S, 10 % ball
    c4: G=512, 1 view 800x800,  R=640000 (even),     S=600, delta=0.01
    `H`/`n_cams` shrink the image set for CPU tests (ray geometry unchanged).
    """
    name = name.lower()
    if name == "c1":
        G, C, Hh, R, S = 64, 1, 64, 4096, 64
        delta, kind_d = 6.0 / S, "dense"
    elif name == "c2":
        G, C, Hh, R, S = 128, 100, 800, 128, 600
        delta, kind_d = 0.0125, "ball"
    elif name == "c3":
        G, C, Hh, R, S = 256, 32, 800, 128, 256
        delta, kind_d = 6.0 / S, "ball"
    elif name == "c4":
        G, C, Hh, R, S = 512, 1, 800, 640000, 600
        delta, kind_d = 0.01, "ball"
    else:
        raise ValueError(f"unknown scene {name!r}")
    if H is not None:
        Hh = H
    if n_cams is not None:
        C = n_cams
    kind = kind or kind_d
    pd = GRID_EXTENT / G
    grid = {"ball": ball_grid, "dense": dense_grid, "soft": soft_grid}[kind](G)
    return Scene(name=name, G=G, points_distance=pd, grid=grid, poses=lookat_poses(C), fov=CAMERA_ANGLE_X,
                 imgs=random_images(C, Hh, Hh), rays_per_cam=R, num_samples=S, delta_step=delta)
