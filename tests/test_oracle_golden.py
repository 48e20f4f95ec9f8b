"""The oracle against golden vectors frozen from the UNMODIFIED reference (tests/golden/make_golden.py).

Bit-exact: voxel indices, in-bounds masks, gathered values, ray directions, target pixels.  1e-6: pixels, loss, gradient.
"""
import glob
import os

import numpy as np
import pytest

from oracle import plenoxel_oracle as po

GOLD = os.path.join(os.path.dirname(__file__), "golden")
STEP_CASES = ["nn_dense_g24", "nn_ball_g32", "tri_ball_g24", "tri_dense_g16"]


def load(name):
    return np.load(os.path.join(GOLD, f"{name}.npz"))


def test_fixtures_present():
    names = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz"))}
    assert set(STEP_CASES) | {"even_spread", "adam3"} <= names


@pytest.mark.parametrize("name", STEP_CASES)
def test_oracle_step_matches_reference(name):
    z = load(name)
    mode, S, R = str(z["mode"]), int(z["S"]), int(z["R"])
    grid, poses, imgs, uv = z["grid"], z["poses"], z["imgs"], z["uv"]
    pd, delta, fov = float(z["pd"]), float(z["delta"]), float(z["fov"])
    gmin = po.grid_origin(grid.shape[:3], pd)
    assert np.array_equal(gmin, z["gmin"])
    dirs, targets, _ = po.generate_rays(imgs, poses, fov, uv)
    assert np.array_equal(dirs, z["dirs"])
    assert np.array_equal(targets, z["targets"])
    o = np.repeat(poses[:, :3, 3], R, axis=0)
    rgba, depth, count, lin = po.render_forward(grid, o, dirs, S, delta, gmin, pd, mode)
    assert np.array_equal(lin >= 0, z["inb"])
    assert np.array_equal(count, z["inb"].sum(1).astype(np.int32))
    if mode == "nearest":
        assert np.array_equal(lin.astype(np.int32), z["lin"]), "nearest-neighbour linear indices must be bit-exact"
    ns = po.normalize_positions(po.sample_positions(o, dirs, S, delta), gmin, pd)
    g01 = np.clip(grid, 0, 1).astype(np.float32)
    if mode == "nearest":
        vals, inb = po.gather_nearest(ns, g01)
        vals = vals * inb[..., None]
    else:
        vals, inb = po.trilinear_lookup(ns, g01)
    assert np.array_equal(vals.reshape(-1, 4), z["vals"]), "looked-up sample values must be bit-exact"
    assert np.abs(rgba - z["pix"]).max() <= 1e-6 * np.abs(z["pix"]).max()
    loss, gpix = po.mse_loss(rgba, targets)
    assert abs(loss - float(z["loss"])) <= 1e-6 * float(z["loss"])
    grad = po.render_backward(grid, o, dirs, S, delta, gmin, pd, gpix, mode)
    assert np.abs(grad - z["grad"]).max() <= 1e-6 * np.abs(z["grad"]).max()
    # depth is not a reference quantity: check its definition sum_k w_k t_k against an independent fp64 evaluation
    t = po.sample_steps(S, delta).astype(np.float64)
    v = vals.reshape(o.shape[0], S, 4).astype(np.float64)
    T = np.cumprod(np.concatenate([np.ones((o.shape[0], 1)), 1 - v[:, :-1, 3]], axis=1), axis=1)
    assert np.abs(depth - (v[..., 3] * T * t).sum(1)).max() <= 1e-5 * max(np.abs(depth).max(), 1e-9)


def test_oracle_even_spread_matches_reference():
    z = load("even_spread")
    uv = po.even_spread_uv(2, 100)
    dirs, targets, _ = po.generate_rays(z["imgs"], z["poses"], float(z["fov"]), uv)
    assert np.array_equal(dirs, z["dirs"])
    assert np.array_equal(targets, z["targets"])
    pos = po.sample_positions(np.repeat(z["cam_pos"], 100, axis=0), dirs, 5, 0.3)
    assert np.array_equal(pos.reshape(-1, 3), z["samples"])


def test_oracle_adam_matches_reference():
    z = load("adam3")
    p, n = z["p0"].copy(), z["p0"].shape[0]
    m, v, ga = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for t in range(3):
        p, m, v, ga = po.adam_step(p, z["grads"][t], m, v, ga, float(z["lr"]), t + 1)
        # torch-CPU's sqrt (MKL VML) is not correctly rounded: <= 1 ulp-of-update differences on ~0.03 % of elements
        assert np.abs(p - z["params"][t]).max() <= 2.4e-7
        assert (p == z["params"][t]).mean() > 0.995
    assert np.array_equal(m, z["exp_avg"])
    assert np.array_equal(v, z["exp_avg_sq"])
    assert np.array_equal(ga, z["gabs"])


def test_composite_backward_matches_finite_differences():
    rng = np.random.default_rng(0)
    s = rng.random((3, 7, 4))
    s[1, 3, 3] = 1.0                      # an opaque sample: T becomes exactly 0 behind it
    g = rng.standard_normal((3, 4))
    ana = po.composite_backward(s, g)
    num = np.zeros_like(s)
    eps = 1e-6
    for idx in np.ndindex(*s.shape):
        sp, sm = s.copy(), s.copy()
        sp[idx] += eps
        sm[idx] -= eps
        num[idx] = ((po.composite(sp, dtype=np.float64) - po.composite(sm, dtype=np.float64)) * g).sum() / (2 * eps)
    assert np.abs(ana - num).max() <= 1e-8


def test_round_half_even_and_bounds_edges():
    """A5: n = -0.5 -> index 0 (inside); n = X - 0.5 rounds to X (outside) for even X; modulo is non-negative."""
    ns = np.array([[-0.5, 0.0, 0.0], [63.5, 0.0, 0.0], [62.5, 0.0, 0.0], [-0.50001, 0.0, 0.0], [-1.0, 64.0, -65.0]], np.float32)
    idx, inb = po.nearest_indices(ns, (64, 64, 64))
    assert idx[:, 0].tolist() == [0, 64, 62, -1, -1]
    assert inb.tolist() == [True, False, True, False, False]
    grid = np.arange(64 * 64 * 64 * 4, dtype=np.float32).reshape(64, 64, 64, 4)
    vals, _ = po.gather_nearest(ns[4:], grid)
    assert np.array_equal(vals[0], grid[63, 0, 63])


def test_oracle_tv_loss_matches_reference():
    z = load("tv_g12")
    loss, grad = po.tv_loss(z["grid"])
    assert abs(loss - float(z["loss"])) <= 1e-6 * float(z["loss"])
    assert np.abs(grad - z["grad"]).max() <= 1e-6 * np.abs(z["grad"]).max()
    assert po.tv_loss(np.full((3, 3, 3, 4), 0.25))[1].max() == 0.0        # constant grid: the reference gives NaN, we give 0


def _oracle_inference_image(z):
    """visulize_3d_in_2d (src/visualization.py:111-154) restated on the oracle: clip, alpha threshold, even-spread rays of
    one camera, nearest lookup without further clamping, composite, x255 round clip uint8, transpose."""
    grid = np.clip(z["grid"], 0.0, 1.0).astype(np.float32)
    grid[..., 3][grid[..., 3] < float(z["threshold"])] = 0.0
    res, S = int(z["res"]), int(z["S"])
    uv = po.even_spread_uv(1, res * res)
    dirs, _, _ = po.generate_rays(z["imgs"], z["poses"], float(z["fov"]), uv)
    o = np.repeat(z["poses"][:, :3, 3], res * res, axis=0)
    gmin = po.grid_origin(grid.shape[:3], float(z["pd"]))
    rgba, _, _, _ = po.render_forward(grid, o, dirs, S, float(z["delta"]), gmin, float(z["pd"]), clamp=False)
    img = (rgba * 255).round().clip(0, 255).astype(np.uint8).reshape(res, res, 4)
    return np.transpose(img, (1, 0, 2))


def test_oracle_inference_image_matches_reference():
    z = load("inference_g24")
    img = _oracle_inference_image(z)
    assert img.shape == z["image"].shape and img.dtype == np.uint8
    assert np.abs(img.astype(int) - z["image"].astype(int)).max() <= 1          # a value on a .5 boundary may round either way
    assert (img == z["image"]).mean() > 0.99
