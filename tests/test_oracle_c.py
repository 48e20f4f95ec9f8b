"""The plain-C restatement (oracle/plenoxel_oracle.c) against the reference's golden vectors and against the numpy oracle.

Three statements of the same algorithm now have to agree: the reference (frozen in tests/golden/*.npz), the numpy oracle
and the C oracle.  Bit-exact: ray directions, target pixels, linear indices, masks, counts, pixels and depth (C vs numpy),
Adam state.  Summation order only (1e-12): the float64 gradients.  1e-6: pixels / loss / gradient against the reference.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import plenoxel_oracle as po
from plenoxels_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
STEP_CASES = ["nn_dense_g24", "nn_ball_g32", "tri_ball_g24", "tri_dense_g16"]


def load(name):
    return np.load(os.path.join(GOLD, f"{name}.npz"))


def test_library_builds_and_loads():
    assert os.path.exists(co.build())
    assert co.load().plxo_version() == 1


@pytest.mark.parametrize("name", STEP_CASES)
def test_c_oracle_step_matches_reference_golden(name):
    z = load(name)
    mode, S, R = str(z["mode"]), int(z["S"]), int(z["R"])
    grid, poses, imgs, uv = z["grid"], z["poses"], z["imgs"], z["uv"]
    pd, delta, fov = float(z["pd"]), float(z["delta"]), float(z["fov"])
    gmin = z["gmin"]
    dirs, targets, _ = co.generate_rays(imgs, poses, fov, uv)
    assert np.array_equal(dirs, z["dirs"]), "ray directions must be bit-exact against the reference"
    assert np.array_equal(targets, z["targets"])
    o = np.repeat(poses[:, :3, 3], R, axis=0)
    rgba, depth, count, lin = co.render_forward(grid, o, dirs, S, delta, gmin, pd, mode)
    assert np.array_equal(lin >= 0, z["inb"])
    assert np.array_equal(count, z["inb"].sum(1).astype(np.int32))
    if mode == "nearest":
        assert np.array_equal(lin.astype(np.int32), z["lin"]), "nearest-neighbour linear indices must be bit-exact"
    assert np.abs(rgba - z["pix"]).max() <= 1e-6 * np.abs(z["pix"]).max()
    loss, gpix = co.mse_loss(rgba, targets)
    assert abs(loss - float(z["loss"])) <= 1e-6 * float(z["loss"])
    grad = co.render_backward(grid, o, dirs, S, delta, gmin, pd, gpix, mode)
    assert np.abs(grad - z["grad"]).max() <= 1e-6 * np.abs(z["grad"]).max()


@pytest.mark.parametrize("mode", ["nearest", "trilinear"])
@pytest.mark.parametrize("clamp", [True, False])
def test_c_oracle_equals_numpy_oracle(mode, clamp):
    G, C_, R, S = 20, 3, 40, 56
    pd, delta = synth.GRID_EXTENT / G, 6.0 / S
    grid = synth.dense_grid(G, seed=3)[:, :19, :17].numpy().copy()           # non-cubic, values in [-0.2, 1.2]
    poses, imgs, uv = synth.lookat_poses(C_).numpy(), synth.random_images(C_, 12, 12, seed=4).numpy(), synth.random_uv(C_, R, seed=5).numpy()
    gmin = po.grid_origin(grid.shape[:3], pd)
    dn, tn, pn = po.generate_rays(imgs, poses, synth.CAMERA_ANGLE_X, uv)
    dc, tc, pc = co.generate_rays(imgs, poses, synth.CAMERA_ANGLE_X, uv)
    assert np.array_equal(dn, dc) and np.array_equal(tn, tc) and np.array_equal(pn, pc)
    o = np.repeat(poses[:, :3, 3], R, axis=0)
    rn, depn, cn, ln = po.render_forward(grid, o, dn, S, delta, gmin, pd, mode, clamp)
    rc, depc, cc, lc = co.render_forward(grid, o, dn, S, delta, gmin, pd, mode, clamp)
    assert np.array_equal(ln, lc) and np.array_equal(cn, cc)
    assert np.array_equal(rn, rc), "pixels: same fp32 operations in the same order"
    assert np.array_equal(depn, depc)
    lossn, gn = po.mse_loss(rn, tn, n_global=2 * o.shape[0])
    lossc, gc = co.mse_loss(rc, tc, n_global=2 * o.shape[0])
    assert abs(lossn - lossc) <= 1e-12 * lossn and np.array_equal(gn, gc)      # loss: summation order only
    for beta in (0.0, 5e-3):
        bn = po.render_backward(grid, o, dn, S, delta, gmin, pd, gn, mode, clamp, beta)
        bc = co.render_backward(grid, o, dn, S, delta, gmin, pd, gn, mode, clamp, beta)
        assert np.abs(bn - bc).max() <= 1e-12 * np.abs(bn).max()


def test_c_oracle_adam_is_bit_identical_to_numpy_oracle_and_matches_torch_golden():
    z = load("adam3")
    p, m, v, ga = z["p0"].copy(), np.zeros_like(z["p0"]), np.zeros_like(z["p0"]), np.zeros_like(z["p0"])
    pn, mn, vn, gan = p.copy(), m.copy(), v.copy(), ga.copy()
    for step in range(1, 4):
        g = z["grads"][step - 1]
        p, m, v, ga = co.adam_step(p, g, m, v, ga, float(z["lr"]), step)
        pn, mn, vn, gan = po.adam_step(pn, g, mn, vn, gan, float(z["lr"]), step)
        assert np.array_equal(p, pn) and np.array_equal(m, mn) and np.array_equal(v, vn) and np.array_equal(ga, gan)
        # torch-CPU's sqrt is not correctly rounded (see the numpy oracle): parameters agree to 1 ulp, moments exactly
        assert np.abs(p - z["params"][step - 1]).max() <= 2e-7
    assert np.array_equal(m, z["exp_avg"]) and np.array_equal(v, z["exp_avg_sq"]) and np.array_equal(ga, z["gabs"])


def test_c_oracle_whole_steps_equal_numpy_oracle():
    G, C_, R, S = 16, 2, 32, 48
    pd, delta = synth.GRID_EXTENT / G, 6.0 / S
    grid = synth.ball_grid(G, seed=8).numpy().copy()
    poses, imgs = synth.lookat_poses(C_).numpy(), synth.random_images(C_, 10, 10, seed=9).numpy()
    gmin = po.grid_origin(grid.shape[:3], pd)
    o = np.repeat(poses[:, :3, 3], R, axis=0)
    sn = [grid.copy(), np.zeros_like(grid), np.zeros_like(grid), np.zeros_like(grid)]
    sc = [a.copy() for a in sn]
    for step in range(1, 4):
        uv = synth.random_uv(C_, R, seed=20 + step).numpy()
        dirs, targets, _ = po.generate_rays(imgs, poses, synth.CAMERA_ANGLE_X, uv)
        ln, gn, *sn = po.train_step(*sn, o, dirs, targets, S, delta, gmin, pd, 0.0075, step)
        lc, gc, *sc = co.train_step(*sc, o, dirs, targets, S, delta, gmin, pd, 0.0075, step)
        assert abs(ln - lc) <= 1e-12 * abs(ln)
        assert np.abs(gn - gc).max() <= 1e-12 * np.abs(gn).max()
        # the fp32 cast of the float64 gradient may differ in the last bit between the two summation orders; Adam's first
        # steps move every touched cell by ~lr * sign(g), so compare the state to 1e-6 instead of bitwise
        for a, b in zip(sn, sc):
            assert np.quantile(np.abs(a - b), 0.999) <= 1e-6


def test_c_oracle_full_baseline_size_finishes_in_seconds():
    """BASELINE config #2 geometry at full size (12 800 rays x 600 samples, 128^3): what the numpy oracle only subsamples."""
    sc = synth.make_scene("c2", H=8)
    uv = synth.random_uv(sc.poses.shape[0], sc.rays_per_cam, seed=77).numpy()
    dirs, _, _ = co.generate_rays(sc.imgs.numpy(), sc.poses.numpy(), sc.fov, uv)
    o = np.repeat(sc.poses.numpy()[:, :3, 3], sc.rays_per_cam, axis=0)
    gmin = po.grid_origin(sc.grid.shape[:3], sc.points_distance)
    rgba, depth, count, lin = co.render_forward(sc.grid.numpy(), o, dirs, sc.num_samples, sc.delta_step, gmin, sc.points_distance)
    assert rgba.shape == (12800, 4) and lin.shape == (12800, 600)
    frac = (lin >= 0).mean()
    assert 0.3 < frac < 0.6, f"in-bounds fraction {frac}"
    # every 97th ray against the numpy oracle, bit for bit
    sel = np.arange(0, 12800, 97)
    rn, dn, cn, ln = po.render_forward(sc.grid.numpy(), o[sel], dirs[sel], sc.num_samples, sc.delta_step, gmin, sc.points_distance)
    assert np.array_equal(ln, lin[sel]) and np.array_equal(cn, count[sel]) and np.array_equal(rn, rgba[sel]) and np.array_equal(dn, depth[sel])


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["full_c2", "full_c3"])
def test_c_oracle_matches_the_reference_at_full_baseline_size(name):
    """tests/golden/full_c2.npz / full_c3.npz hold what the UNMODIFIED reference computed for a whole config-#2 / #3 batch
    (make_golden.py full): the inputs are regenerated from their seeds and checked by SHA-256, then every ray direction,
    target pixel and linear sample index (7.68 M / 1.05 M of them, by SHA-256) must be bit-exact, per-ray counts equal,
    pixels / loss / gradient (sums + strided subset) within 1e-6."""
    z = load(name)
    sc = synth.make_scene(str(z["scene"]), H=int(z["H"]))
    C_, R, S = sc.poses.shape[0], sc.rays_per_cam, sc.num_samples
    uv = synth.random_uv(C_, R, seed=int(z["uv_seed"])).numpy()
    grid, poses, imgs = sc.grid.numpy(), sc.poses.numpy(), sc.imgs.numpy()
    assert (_sha(grid), _sha(uv), _sha(poses), _sha(imgs)) == (str(z["sha_grid"]), str(z["sha_uv"]), str(z["sha_poses"]), str(z["sha_imgs"])), \
        "the seeded inputs no longer reproduce the ones the fixture was made from"
    gmin = po.grid_origin(grid.shape[:3], sc.points_distance)
    assert np.array_equal(gmin, z["gmin"])
    dirs, targets, _ = co.generate_rays(imgs, poses, sc.fov, uv)
    assert _sha(dirs) == str(z["sha_dirs"]) and _sha(targets) == str(z["sha_targets"])
    o = np.repeat(poses[:, :3, 3], R, axis=0)
    rgba, _, count, lin = co.render_forward(grid, o, dirs, S, sc.delta_step, gmin, sc.points_distance)
    assert _sha(lin.astype(np.int32)) == str(z["sha_lin"]), "every linear index of the full batch, bit-exact"
    assert np.array_equal(count, z["count"])
    assert np.abs(rgba - z["pix"]).max() <= 1e-6 * np.abs(z["pix"]).max()
    loss, gpix = co.mse_loss(rgba, targets)
    assert abs(loss - float(z["loss"])) <= 1e-6 * float(z["loss"])
    grad = co.render_backward(grid, o, dirs, S, sc.delta_step, gmin, sc.points_distance, gpix)
    gmax = float(z["grad_max"])
    assert abs(np.abs(grad).max() - gmax) <= 1e-6 * gmax
    assert np.abs(grad[::7, ::5, ::3] - z["grad_subset"]).max() <= 1e-6 * gmax
    assert np.abs(grad.reshape(-1, 4).sum(0) - z["grad_sum"]).max() <= 1e-5 * np.abs(z["grad_abs_sum"]).max()
    assert np.abs(np.abs(grad).reshape(-1, 4).sum(0) - z["grad_abs_sum"]).max() <= 1e-5 * np.abs(z["grad_abs_sum"]).max()
    assert int((np.abs(grad).sum(-1) > 0).sum()) == int(z["grad_nonzero_cells"])


def test_c_oracle_handles_non_finite_and_degenerate_rays_like_numpy():
    """NaN / inf / astronomically large coordinates are out of bounds (numpy's float -> int64 cast gives INT64_MIN there; the
    C cast would be undefined behaviour without the guard), a zero direction stays at the origin, S = 0 renders nothing."""
    G, S = 12, 24
    pd, delta = synth.GRID_EXTENT / G, 6.0 / S
    grid = synth.dense_grid(G, seed=1).numpy()
    gmin = po.grid_origin(grid.shape[:3], pd)
    o = np.array([[0, 0, 4], [0, 0, 4], [0, 0, 0], [0, 0, 4], [1e30, 0, 0], [0, 0, 4]], np.float32)
    d = np.array([[0, 0, -1], [np.nan, 0, -1], [0, 0, 0], [np.inf, 0, -1], [0, 0, -1], [1e38, 1e38, -1e38]], np.float32)
    with np.errstate(all="ignore"):
        rn, dn, cn, ln = po.render_forward(grid, o, d, S, delta, gmin, pd)
    rc, dc, cc, lc = co.render_forward(grid, o, d, S, delta, gmin, pd)
    assert np.array_equal(ln, lc) and np.array_equal(cn, cc)
    assert cn[0] > 0 and cn[1] == 0 and cn[2] == S and cn[3] == 0 and cn[4] == 0
    assert np.array_equal(rn, rc, equal_nan=True) and np.array_equal(dn, dc, equal_nan=True)
    r0 = co.render_forward(grid, o, d, 0, delta, gmin, pd)
    assert not r0[0].any() and not r0[2].any() and r0[3].shape == (6, 0)


def test_c_oracle_even_spread_lattice_matches_reference_golden():
    z = load("even_spread")
    uv = co.even_spread_uv(2, 100)
    assert np.array_equal(uv, po.even_spread_uv(2, 100))
    for n in (1, 2, 17, 100, 800):
        assert np.array_equal(co.even_spread_uv(1, n * n), po.even_spread_uv(1, n * n)), n
    dirs, targets, _ = co.generate_rays(z["imgs"], z["poses"], float(z["fov"]), uv)
    assert np.array_equal(dirs, z["dirs"]) and np.array_equal(targets, z["targets"])


def test_c_oracle_tv_loss_matches_reference_golden():
    z = load("tv_g12")
    loss, grad = co.tv_loss(z["grid"])
    assert abs(loss - float(z["loss"])) <= 1e-6 * float(z["loss"])
    assert np.abs(grad - z["grad"]).max() <= 1e-6 * np.abs(z["grad"]).max()
    ln, gn = po.tv_loss(z["grid"])
    assert abs(loss - ln) <= 1e-12 * ln and np.abs(grad - gn).max() <= 1e-12
    l0, g0 = co.tv_loss(np.full((3, 4, 5, 4), 0.25, np.float32))
    assert l0 == 0.0 and not g0.any()


def test_c_oracle_inference_image_matches_reference_golden():
    """visulize_3d_in_2d (src/visualization.py:111-154) restated on the C oracle: clip, alpha threshold, even-spread rays of
    one camera, nearest lookup without further clamping, composite, x255 round clip uint8, transpose — the reference's uint8
    image bit for bit."""
    z = load("inference_g24")
    grid = np.clip(z["grid"], 0.0, 1.0).astype(np.float32)
    grid[..., 3][grid[..., 3] < float(z["threshold"])] = 0.0
    res, S = int(z["res"]), int(z["S"])
    uv = co.even_spread_uv(1, res * res)
    dirs, _, _ = co.generate_rays(z["imgs"], z["poses"], float(z["fov"]), uv)
    o = np.repeat(z["poses"][:, :3, 3], res * res, axis=0)
    gmin = po.grid_origin(grid.shape[:3], float(z["pd"]))
    rgba, _, _, _ = co.render_forward(grid, o, dirs, S, float(z["delta"]), gmin, float(z["pd"]), clamp=False)
    img = np.transpose((rgba * 255).round().clip(0, 255).astype(np.uint8).reshape(res, res, 4), (1, 0, 2))
    assert np.array_equal(img, z["image"])
