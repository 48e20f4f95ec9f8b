"""Edge cases on the GPU path, each against the oracle: empty inputs, S = 0, rays that miss the grid, the all-zero grid
fit() starts from, cameras inside the grid, non-cubic grids, delta = 0, saturated grids, checkpoint format."""
import numpy as np
import pytest
import torch

from oracle import plenoxel_oracle as po
from plenoxels_b200 import ops, synth
from plenoxels_b200.trainer import VoxelTrainer
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _oracle(grid, o, d, S, delta, gmin, pd, targets, mode="nearest"):
    rgba, depth, count, lin = po.render_forward(grid, o, d, S, delta, gmin, pd, mode)
    loss, gpix = po.mse_loss(rgba, targets)
    grad = po.render_backward(grid, o, d, S, delta, gmin, pd, gpix, mode)
    return rgba, depth, count, loss, grad


def _gpu(grid, o, d, S, delta, gmin, pd, targets, mode="nearest"):
    g = torch.from_numpy(grid).to(DEV).requires_grad_(True)
    to = torch.from_numpy(o).to(DEV)
    td = torch.from_numpy(d).to(DEV)
    rgba, depth, count = ops.render_rays(g, to, td, S, delta, gmin, pd, mode=mode, return_depth=True, return_count=True)
    loss = torch.nn.functional.mse_loss(rgba, torch.from_numpy(targets).to(DEV))
    loss.backward()
    gg = torch.zeros_like(g.detach())
    rgba_f, loss_f = ops.render_train(g.detach(), gg, S, delta, gmin, pd, origins=to, dirs=td,
                                      targets=torch.from_numpy(targets).to(DEV)) if mode == "nearest" else (rgba, loss)
    n = lambda t: t.detach().cpu().numpy()
    return (n(rgba), n(depth), n(count), float(loss), n(g.grad), n(rgba_f), float(loss_f),
            n(gg) if mode == "nearest" else n(g.grad))


def _check(grid, o, d, S, delta, pd, mode="nearest", seed=0):
    rng = np.random.default_rng(seed)
    targets = rng.random((d.shape[0], 4)).astype(np.float32)
    gmin = po.grid_origin(grid.shape[:3], pd)
    orgba, odepth, ocount, oloss, ograd = _oracle(grid, o, d, S, delta, gmin, pd, targets, mode)
    rgba, depth, count, loss, grad, rgba_f, loss_f, grad_f = _gpu(grid, o, d, S, delta, gmin, pd, targets, mode)
    assert np.array_equal(count, ocount)
    scale = max(np.abs(orgba).max(), 1e-12)
    assert np.abs(rgba - orgba).max() <= 1e-5 * scale and np.abs(rgba_f - orgba).max() <= 1e-5 * scale
    assert np.abs(depth - odepth).max() <= 1e-5 * max(np.abs(odepth).max(), 1e-12)
    assert abs(loss - oloss) <= 1e-5 * oloss and abs(loss_f - oloss) <= 1e-5 * oloss
    gs = max(np.abs(ograd).max(), 1e-20)
    assert np.abs(grad - ograd).max() <= 1e-5 * gs and np.abs(grad_f - ograd).max() <= 1e-5 * gs
    return ocount


def _dirs(n, seed=1):
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def test_zero_rays_and_zero_samples(plx_lib):
    grid = synth.dense_grid(8).to(DEV)
    o, d = torch.zeros(0, 3, device=DEV), torch.zeros(0, 3, device=DEV)
    assert tuple(ops.render_rays(grid, torch.zeros(1, 3, device=DEV), d, 16, 0.1, (0, 0, 0), 0.4).shape) == (0, 4)
    o = torch.tensor([[0.0, 0.0, 3.0]], device=DEV)
    d = torch.tensor([[0.0, 0.0, -1.0], [0.0, 0.6, -0.8]], device=DEV)
    rgba, depth, count = ops.render_rays(grid, o, d, 0, 0.0, (0, 0, 0), 0.4, rays_per_origin=2, return_depth=True, return_count=True)
    assert float(rgba.abs().max()) == 0.0 and float(depth.abs().max()) == 0.0 and count.tolist() == [0, 0]
    # S = 0 through the eager functions too (scripts/visulize_camera_and_grid.py:36-46 calls with num_samples=0, delta_step=0)
    assert tuple(ops.sample_points(o, d, 0, 0.0, rays_per_origin=2).shape) == (0, 3)
    assert tuple(ops.composite(torch.zeros(1, 2, 0, 4, device=DEV)).shape) == (1, 2, 4)


def test_rays_that_miss_the_grid(plx_lib):
    grid = synth.dense_grid(16).numpy()
    o = np.array([[0.0, 0.0, 5.0]] * 64, np.float32)
    d = _dirs(64)
    d[:, 2] = np.abs(d[:, 2])                          # all pointing away from the grid
    counts = _check(grid, o, d, 64, 0.1, 0.2)
    assert counts.sum() == 0


def test_zero_initialised_grid_like_fit_start(plx_lib):
    """scripts/train.py:85-87: the grid starts at exactly 0 — every in-bounds sample is transparent but carries gradient."""
    grid = np.zeros((24, 24, 24, 4), np.float32)
    poses = synth.lookat_poses(4).numpy()
    o = np.repeat(poses[:, :3, 3], 16, axis=0)
    d = (-o / np.linalg.norm(o, axis=1, keepdims=True) + 0.05 * _dirs(64, 3)).astype(np.float32)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    counts = _check(grid, o, d, 96, 6.0 / 96, synth.GRID_EXTENT / 24)
    assert counts.sum() > 0


def test_camera_inside_the_grid_and_non_cubic_grid(plx_lib):
    rng = np.random.default_rng(5)
    grid = (rng.random((20, 12, 28, 4)) * 1.3 - 0.15).astype(np.float32)
    grid[..., 3] *= 0.2
    o = (rng.random((96, 3)).astype(np.float32) - 0.5) * 0.8          # origins inside the box
    _check(grid, o, _dirs(96, 6), 80, 0.05, 0.11)


@pytest.mark.parametrize("mode", ["nearest", "trilinear"])
def test_saturated_grid_every_ray_terminates(plx_lib, mode):
    grid = np.full((16, 16, 16, 4), 1.7, np.float32)                   # clips to (1,1,1,1): first in-bounds sample is opaque
    grid[::2, :, :, :3] = -0.3
    poses = synth.lookat_poses(3).numpy()
    o = np.repeat(poses[:, :3, 3], 32, axis=0)
    d = (-o / np.linalg.norm(o, axis=1, keepdims=True) + 0.08 * _dirs(96, 7)).astype(np.float32)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    _check(grid, o, d, 128, 6.0 / 128, synth.GRID_EXTENT / 16, mode=mode)


def test_delta_zero_puts_every_sample_at_the_origin(plx_lib):
    rng = np.random.default_rng(8)
    grid = (rng.random((8, 8, 8, 4)) * 0.5).astype(np.float32)
    o = (rng.random((16, 3)).astype(np.float32) - 0.5)
    _check(grid, o, _dirs(16, 9), 12, 0.0, 0.4)


def test_large_sample_count_uses_unfused_fallback_in_train_step(plx_lib):
    """num_samples beyond the fused kernel's shared-memory index cache: plx_train_step falls back to K1 + K2."""
    G, C, H, R, S = 16, 2, 8, 8, 30000
    pd, delta = synth.GRID_EXTENT / G, 6.0 / S
    grid, poses, imgs, uv = synth.ball_grid(G), synth.lookat_poses(C), synth.random_images(C, H, H), synth.random_uv(C, R)
    tr = VoxelTrainer(grid.to(DEV), pd, poses.to(DEV), synth.CAMERA_ANGLE_X, imgs.to(DEV), R, S, delta, lr=0.0075)
    loss = float(tr.step(uv.to(DEV)))
    gmin = po.grid_origin(grid.shape[:3], pd)
    dirs, targets, _ = po.generate_rays(imgs.numpy(), poses.numpy(), synth.CAMERA_ANGLE_X, uv.numpy())
    o = np.repeat(poses[:, :3, 3].numpy(), R, axis=0)
    rgba, _, _, _ = po.render_forward(grid.numpy(), o, dirs, S, delta, gmin, pd)
    oloss, gpix = po.mse_loss(rgba, targets)
    ograd = po.render_backward(grid.numpy(), o, dirs, S, delta, gmin, pd, gpix)
    assert abs(loss - oloss) <= 1e-5 * oloss
    assert rel_err(tr.grad_abs_sum.cpu().numpy(), np.abs(ograd)) <= 1e-5


def test_checkpoint_has_the_reference_format(plx_lib, tmp_path):
    """scripts/train.py:194-210: {grid, grid_grad, param{...}} on the CPU, float32, no grad — loadable by torch.load as
    scripts/compare_inference_to_image.py:41-49 and scripts/visulize_grid.py:23-30 do."""
    sc = synth.make_scene("c1", H=16)
    tr = VoxelTrainer(sc.grid.to(DEV), sc.points_distance, sc.poses.to(DEV), sc.fov, sc.imgs.to(DEV), 64, 32, 6.0 / 32, lr=0.01)
    tr.step(synth.random_uv(1, 64).to(DEV))
    path = tmp_path / "grid_cells_trained.pth"
    torch.save(tr.checkpoint(), path)
    ck = torch.load(path)
    assert set(ck) == {"grid", "grid_grad", "param"}
    assert set(ck["param"]) == {"device", "number_of_rays", "num_samples", "delta_step", "even_spread", "camera_ray",
                                "points_distance", "gridsize"}
    for k in ("grid", "grid_grad"):
        assert ck[k].shape == (64, 64, 64, 4) and ck[k].dtype == torch.float32 and not ck[k].is_cuda and not ck[k].requires_grad
    assert ck["param"]["gridsize"] == [64, 64, 64] and isinstance(ck["param"]["points_distance"], float)
    assert float(ck["grid_grad"].sum()) > 0


def test_resume_continues_bit_exactly(plx_lib):
    """checkpoint() + resume_state() -> load_state(): a resumed trainer takes the same next steps as the original
    (deterministic here: 1 ray per cell at most would be needed for bit-exactness, so compare to atomics' noise level)."""
    sc = synth.make_scene("c1", H=16)
    mk = lambda: VoxelTrainer(sc.grid.to(DEV), sc.points_distance, sc.poses.to(DEV), sc.fov, sc.imgs.to(DEV), 256, 48, 6.0 / 48, lr=0.01)
    a = mk()
    uvs = [synth.random_uv(1, 256, seed=20 + i).to(DEV) for i in range(4)]
    for u in uvs[:2]:
        a.step(u)
    ck = {**a.checkpoint(), **a.resume_state()}
    b = mk()
    b.load_state(ck)
    assert b.step_count == 2 and torch.equal(b.grid, a.grid) and torch.equal(b.exp_avg_sq, a.exp_avg_sq)
    for u in uvs[2:]:
        la, lb = float(a.step(u)), float(b.step(u))
        assert abs(la - lb) <= 1e-6 * abs(la)
    assert float((a.grid - b.grid).abs().max()) <= 1e-5
