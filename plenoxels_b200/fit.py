"""`fit()` with the reference's signature and semantics (scripts/train.py:68-210) on the fused kernels.

The reference's own `scripts/train.py` runs unchanged against this package (src/ shim + lazy bridge), but its loop is
~40 eager torch calls per step.  This is the same training procedure — same arguments, same coarse-to-fine schedule
(:91-123), same losses (:156-177), same Adam (:89, :180-182), same `.pth` payload (:194-210) — issued as a handful of
kernel launches per step:

    [separable average pooling of the full grid]      plx_avgpool3d_fwd        (while the receptive field > 1)
    ray generation + march + MSE (+ beta) + backward   plx_render_train (K12)   on the (pooled) grid
    [TV loss + its gradient]                           plx_tv_loss              (tv > 0 and receptive field < 19)
    [pooling backward to the full grid]                plx_avgpool3d_bwd
    Adam + |grad| accumulation + gradient clear        plx_adam_step (K3)

Nothing is copied to the host per step (the reference syncs on `.cpu()` for its progress bar every step, :159); the
losses are read back every `log_every` steps.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import ops
from .data_processing import load_data, load_image_bytes_from_path

STEPS_PER_FIELD = 5                                       # scripts/train.py:92


def receptive_field_schedule():
    """[93, 91, ..., 3] — scripts/train.py:91."""
    return [i for i in range(93, 2, -2) if i % 2 != 0]


def receptive_field_at(step: int, schedule=None) -> int:
    """Pooling window of training step `step` (0 = full resolution) — scripts/train.py:108."""
    schedule = receptive_field_schedule() if schedule is None else schedule
    k = step // STEPS_PER_FIELD
    return schedule[k] if k < len(schedule) else 0


class GridFitter:
    """State of one fit: full-resolution grid, Adam moments, running |grad|; `step(i)` is one iteration of :104-191."""

    def __init__(self, gridsize, points_distance, poses, fov, imgs, number_of_rays, num_samples, delta_step, lr, tv=0.0,
                 beta=0.0, even_spread=False, progressive_growing=True, schedule=None, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.PlxError("plenoxels_b200.fit runs on CUDA (sm_100a) only; there is no CPU fallback")
        L.load()
        self.device = dev
        self.gridsize = [int(g) for g in gridsize]
        self.pd0 = points_distance
        self.poses = poses.to(dev).float().contiguous()
        self.imgs = (imgs if imgs.dtype == torch.uint8 else imgs.float()).to(dev).contiguous()    # uint8: converted per fetch
        self.fov = float(fov)
        self.S, self.delta, self.lr = int(num_samples), delta_step, lr
        self.tv, self.beta = float(tv), float(beta)
        self.even_spread = bool(even_spread)
        self.R = int(np.round(np.sqrt(number_of_rays)) ** 2) if even_spread else int(number_of_rays)   # :144-145
        self.progressive = bool(progressive_growing)
        self.schedule = receptive_field_schedule() if schedule is None else list(schedule)
        X, Y, Z = self.gridsize
        self.grid = torch.zeros((X, Y, Z, 4), dtype=torch.float32, device=dev)                      # :79, :85-87
        self.grad = torch.zeros_like(self.grid)
        self.exp_avg = torch.zeros_like(self.grid)
        self.exp_avg_sq = torch.zeros_like(self.grid)
        self.grad_abs_sum = torch.zeros_like(self.grid)                                              # :83
        self.steps_done = 0
        self.last = {}
        self._lattice = None

    def _uv(self):
        C_ = self.poses.shape[0]
        if self.even_spread:                                                                          # :220-223
            if self._lattice is None:
                n = int(np.round(np.sqrt(self.R)))
                line = torch.linspace(0, end=1, steps=n, device=self.device)
                self._lattice = torch.cartesian_prod(line, line).repeat([C_, 1, 1]).contiguous()
            return self._lattice
        return torch.rand(C_, self.R, 2, device=self.device)                                         # :227

    def step(self, i: int | None = None):
        """One training iteration; returns (mse, tv_loss, beta_loss) as device tensors / None (no host sync)."""
        i = self.steps_done if i is None else i
        rf = receptive_field_at(i, self.schedule)
        pooled = rf > 1 and self.progressive
        if pooled:                                                                                    # :110-118
            stride = max(1, rf // 4)
            start = int(rf / 2)
            cells = ops.avgpool3d_grid(self.grid, rf, stride)
            gmin = ops.grid_origin(self.gridsize, self.pd0, start)
            pd = self.pd0 * stride
            grad_cells = torch.zeros_like(cells)
        else:                                                                                         # :120-123
            cells, gmin, pd, grad_cells = self.grid, ops.grid_origin(self.gridsize, self.pd0), self.pd0, self.grad
        n_samples_total = self.poses.shape[0] * self.R * self.S
        bom = self.beta / n_samples_total if self.beta > 0 else 0.0
        rgba, mse = ops.render_train(cells, grad_cells, self.S, self.delta, gmin, pd, imgs=self.imgs, poses=self.poses,
                                     fov=self.fov, uv=self._uv(), beta_over_m=bom)                  # :130-157, :170-177
        tvl = None
        if self.tv > 0 and rf < 19:                                                                   # :163-168
            tvl = ops.tv_loss_(cells, self.tv, grad_cells)
        if pooled:
            ops.avgpool3d_grid_backward_into(grad_cells, self.gridsize, rf, stride, self.grad)
        self.steps_done += 1
        ops.adam_step(self.grid, self.grad, self.exp_avg, self.exp_avg_sq, self.grad_abs_sum, self.steps_done, self.lr)  # :180-184
        self.last = {"mse": mse, "tv": tvl, "rf": rf, "shape": tuple(cells.shape)}
        return mse, tvl

    def checkpoint(self, number_of_rays, even_spread):
        """scripts/train.py:194-210."""
        return {"grid": self.grid.detach().cpu(), "grid_grad": self.grad_abs_sum.detach().cpu(),
                "param": {"device": str(self.device), "number_of_rays": number_of_rays, "num_samples": self.S,
                          "delta_step": self.delta, "even_spread": even_spread, "camera_ray": False,
                          "points_distance": self.pd0, "gridsize": self.gridsize}}


def fit(gridsize, points_distance_original, number_of_rays, num_samples, delta_step, lr, tv, beta, steps, even_spread, path,
        transform_path, save_path, device, progressive_growing=True, log_every=50):
    """Same arguments and effect as the reference's `fit` (scripts/train.py:68-210); `log_every` is the only addition."""
    data, imgs = load_image_bytes_from_path(path, transform_path)        # :73, pixels stay uint8 (converted in the kernel)
    poses, _, fov = load_data(data)                                                                   # :74
    fitter = GridFitter(gridsize, points_distance_original, poses, fov, imgs, number_of_rays, num_samples, delta_step, lr,
                        tv=tv, beta=beta, even_spread=even_spread, progressive_growing=progressive_growing, device=device)
    if steps < STEPS_PER_FIELD * len(fitter.schedule):                                                # :101-102
        print(f"not enough steps to for frequency regularization {steps=} < {STEPS_PER_FIELD * len(fitter.schedule)=}")
    for i in range(steps):
        mse, tvl = fitter.step(i)
        if log_every and (i % log_every == 0 or i == steps - 1):
            msg = f"step {i}: grid size: {fitter.last['shape']}, kernel size: {fitter.last['rf']} closs:{float(mse):10.6f}"
            if tvl is not None:
                msg += f" tvloss:{float(tvl):10.6f}"
            print(msg, flush=True)
    torch.save(fitter.checkpoint(number_of_rays, even_spread), save_path)                             # :197-210
    return fitter
