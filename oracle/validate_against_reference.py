"""Pin the oracle against the real reference (build container only; reads /root/reference).

Run:  python oracle/validate_against_reference.py
Prints one line per check and exits non-zero on the first mismatch.  The GPU box has no
/root/reference, so nothing in tests/ -m gpu, smoke() or bench.py runs this; its frozen outputs are
the fixtures made by tests/golden/make_golden.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PLX_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(1, REF)
# `src.*` must resolve to the REFERENCE here.  The reference's src/ has no __init__.py (a namespace package), so this
# repo's src/ shim (a regular package) would always win the import; pin the name to the reference directory instead.
import types  # noqa: E402
_ref_src = types.ModuleType("src")
_ref_src.__path__ = [os.path.join(REF, "src")]
sys.modules["src"] = _ref_src

from oracle import c_oracle as co              # noqa: E402
from oracle import plenoxel_oracle as po      # noqa: E402
from oracle import torch_port as tp           # noqa: E402
from plenoxels_b200 import synth              # noqa: E402

import src.grid_functions as rgf              # noqa: E402  (reference)
import src.ray_sampling as rrs                # noqa: E402  (reference)

assert rgf.__file__.startswith(REF), rgf.__file__

FAILED = []


def check(name, ok, detail=""):
    print(f"[{'ok' if ok else 'FAIL'}] {name} {detail}")
    if not ok:
        FAILED.append(name)


def reference_step_tensors(grid, pd, poses, fov, imgs, R, S, delta, uv, mode="nearest"):
    """Reference functions in the order of scripts/train.py:130-157, RNG replaced by a given uv."""
    dims = grid.shape[:3]
    coords, _, _, _ = rgf.generate_grid(*dims, points_distance=pd, info_size=4, device="cpu")
    orig_rand = torch.rand
    torch.rand = lambda *a, **k: uv.clone()
    try:
        samples, targets, cam_pos, dirs = rrs.sample_camera_rays_batched(
            transform_matrices=poses, camera_angle_x=fov, imgs=imgs, number_of_rays=R, num_samples=S,
            delta_step=delta, even_spread=False, camera_ray=False, device="cpu")
    finally:
        torch.rand = orig_rand
    ns = rrs.normalize_samples_for_indecies(coords, samples, pd)
    g = grid.detach().clone().requires_grad_(True)
    if mode == "nearest":
        idx = torch.round(ns).to(torch.long)
        vals, inb = rgf.get_nearest_voxels(ns, g.clip(0, 1))
        vals = vals * inb.unsqueeze(-1)
    else:
        inb = rgf.find_out_of_bound(ns, g)
        pts = rgf.get_grid_points_indices(ns)
        rgf.fix_out_of_bounds(pts.reshape(-1, 3), g)
        idx = pts
        vals = rgf.trilinear_interpolation(ns, pts, g.clip(0, 1)) * inb.unsqueeze(-1)
    pix = rrs.compute_alpha_weighted_pixels(vals.reshape(poses.shape[0], R, S, 4)).reshape(-1, 4)
    loss = torch.nn.functional.mse_loss(pix, targets)
    loss.backward()
    return dict(coords=coords, samples=samples, targets=targets, dirs=dirs, cam_pos=cam_pos, ns=ns, idx=idx,
                inb=inb, vals=vals.detach(), pix=pix.detach(), loss=loss.detach(), grad=g.grad)


def run_scene(tag, G, C, H, R, S, delta, kind, mode="nearest"):
    pd = synth.GRID_EXTENT / G
    grid = {"ball": synth.ball_grid, "dense": synth.dense_grid, "soft": synth.soft_grid}[kind](G)
    poses = synth.lookat_poses(C)
    imgs = synth.random_images(C, H, H)
    uv = synth.random_uv(C, R)
    fov = synth.CAMERA_ANGLE_X
    ref = reference_step_tensors(grid, pd, poses, fov, imgs, R, S, delta, uv, mode)
    N = C * R

    # ---- numpy oracle
    gmin = po.grid_origin(grid.shape[:3], pd)
    check(f"{tag}: gmin == grid_indices.min(0)", np.array_equal(gmin, ref["coords"].min(0)[0].numpy()))
    dirs, targets, _ = po.generate_rays(imgs.numpy(), poses.numpy(), fov, uv.numpy())
    check(f"{tag}: dirs bit-exact", np.array_equal(dirs, ref["dirs"].numpy()),
          f"max|d|={np.abs(dirs - ref['dirs'].numpy()).max():.2e}")
    check(f"{tag}: targets bit-exact", np.array_equal(targets, ref["targets"].numpy()))
    o = np.repeat(poses[:, :3, 3].numpy(), R, axis=0)
    d_ref = ref["dirs"].numpy()
    pos = po.sample_positions(o, d_ref, S, delta)
    check(f"{tag}: sample positions bit-exact", np.array_equal(pos.reshape(-1, 3), ref["samples"].numpy()))
    ns = po.normalize_positions(pos, gmin, pd)
    check(f"{tag}: normalised coords bit-exact", np.array_equal(ns.reshape(-1, 3), ref["ns"].numpy()))
    rgba, depth, count, lin = po.render_forward(grid.numpy(), o, d_ref, S, delta, gmin, pd, mode)
    if mode == "nearest":
        idx, inb = po.nearest_indices(ns, grid.shape[:3])
        check(f"{tag}: int64 indices bit-exact", np.array_equal(idx.reshape(-1, 3), ref["idx"].numpy()))
    else:
        inb = lin >= 0
    check(f"{tag}: in-bounds mask equal", np.array_equal(inb.reshape(-1), ref["inb"].numpy()),
          f"in-bounds fraction {inb.mean():.3f}")
    check(f"{tag}: per-ray counts equal",
          np.array_equal(count, ref["inb"].reshape(N, S).sum(1).numpy().astype(np.int32)))
    g = np.clip(grid.numpy(), 0, 1).astype(np.float32)
    if mode == "nearest":
        vals, _ = po.gather_nearest(ns, g)
        vals = vals * inb[..., None]
    else:
        vals, _ = po.trilinear_lookup(ns, g)
    check(f"{tag}: gathered values bit-exact", np.array_equal(vals.reshape(-1, 4), ref["vals"].numpy()),
          f"max diff {np.abs(vals.reshape(-1, 4) - ref['vals'].numpy()).max():.2e}")
    scale = np.abs(ref["pix"].numpy()).max()
    err = np.abs(rgba - ref["pix"].numpy()).max() / scale
    check(f"{tag}: pixels <= 1e-6 rel", err <= 1e-6, f"rel err {err:.2e}")
    loss, gpix = po.mse_loss(rgba, ref["targets"].numpy())
    check(f"{tag}: loss <= 1e-6 rel", abs(loss - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"])),
          f"{loss:.8f} vs {float(ref['loss']):.8f}")
    grad = po.render_backward(grid.numpy(), o, d_ref, S, delta, gmin, pd, gpix, mode)
    gref = ref["grad"].numpy()
    gerr = np.abs(grad - gref).max() / np.abs(gref).max()
    check(f"{tag}: grid gradient <= 1e-6 rel (hand-derived reverse recurrence vs reference autograd)",
          gerr <= 1e-6, f"rel err {gerr:.2e}, nonzero cells {int((np.abs(gref).sum(-1) > 0).sum())}")

    # ---- C oracle (oracle/plenoxel_oracle.c): the same quantities straight against the live reference
    dirs_c, targets_c, _ = co.generate_rays(imgs.numpy(), poses.numpy(), fov, uv.numpy())
    check(f"{tag}: C dirs bit-exact", np.array_equal(dirs_c, ref["dirs"].numpy()))
    check(f"{tag}: C targets bit-exact", np.array_equal(targets_c, ref["targets"].numpy()))
    rgba_c, _, count_c, lin_c = co.render_forward(grid.numpy(), o, d_ref, S, delta, gmin, pd, mode)
    check(f"{tag}: C in-bounds mask / counts equal", np.array_equal((lin_c >= 0).reshape(-1), ref["inb"].numpy()) and
          np.array_equal(count_c, ref["inb"].reshape(N, S).sum(1).numpy().astype(np.int32)))
    if mode == "nearest":
        ridx = ref["idx"].numpy().reshape(N, S, 3)
        rlin = (ridx[..., 0] * grid.shape[1] + ridx[..., 1]) * grid.shape[2] + ridx[..., 2]
        check(f"{tag}: C linear indices bit-exact", np.array_equal(lin_c[lin_c >= 0], rlin[lin_c >= 0]))
    cerr = np.abs(rgba_c - ref["pix"].numpy()).max() / scale
    check(f"{tag}: C pixels <= 1e-6 rel (and == numpy oracle: {np.array_equal(rgba_c, rgba)})", cerr <= 1e-6 and np.array_equal(rgba_c, rgba),
          f"rel err {cerr:.2e}")
    loss_c, gpix_c = co.mse_loss(rgba_c, ref["targets"].numpy())
    grad_c = co.render_backward(grid.numpy(), o, d_ref, S, delta, gmin, pd, gpix_c, mode)
    gcerr = np.abs(grad_c - gref).max() / np.abs(gref).max()
    check(f"{tag}: C loss / grid gradient <= 1e-6 rel", abs(loss_c - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"])) and
          gcerr <= 1e-6, f"grad rel err {gcerr:.2e}")

    # ---- torch port (what the cpu_baseline / --impl reference legs time)
    port = tp.ReferenceStep(grid, pd, poses, fov, imgs, R, S, delta, lr=0.0075, mode=mode)
    pix_p, tgt_p = port.forward(uv)
    check(f"{tag}: torch port pixels bit-exact", torch.equal(pix_p.detach(), ref["pix"]))
    check(f"{tag}: torch port targets bit-exact", torch.equal(tgt_p, ref["targets"]))
    loss_p = port.step(uv)
    check(f"{tag}: torch port loss bit-exact", torch.equal(loss_p, ref["loss"]))
    # CPU index_put_(accumulate=True) adds with parallel atomics above a grain size => run-to-run ulp noise
    perr = float((port.grid.grad - ref["grad"]).abs().max() / ref["grad"].abs().max())
    check(f"{tag}: torch port grid gradient <= 1e-6 rel", perr <= 1e-6, f"rel err {perr:.2e}")

    # ---- Adam emulation vs torch.optim.Adam on the reference's gradient, 3 steps
    p = grid.detach().clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=0.0075)
    pn, m, v, ga = grid.numpy().copy(), np.zeros_like(gref), np.zeros_like(gref), np.zeros_like(gref)
    exact = True
    worst = 0.0
    for t in range(1, 4):
        gt = torch.from_numpy(gref * np.float32(1.0 / t)).clone()
        p.grad = gt
        opt.step()
        pn, m, v, ga = po.adam_step(pn, gt.numpy(), m, v, ga, 0.0075, t)
        exact &= np.array_equal(pn, p.detach().numpy())
        worst = max(worst, float(np.abs(pn - p.detach().numpy()).max()))
    st = opt.state[p]
    exact_m = np.array_equal(m, st["exp_avg"].numpy())
    exact_v = np.array_equal(v, st["exp_avg_sq"].numpy())
    check(f"{tag}: Adam params after 3 steps (bit-exact={exact}, m={exact_m}, v={exact_v})",
          worst <= 2.4e-7 and exact_m and exact_v, f"max abs diff {worst:.2e}")


def run_even_spread(tag, G, H, R, S, delta):
    """even_spread=True path (linspace lattice) incl. the S=0 call of scripts/visulize_camera_and_grid.py:36-46."""
    poses = synth.lookat_poses(3)
    imgs = synth.random_images(3, H, H)
    fov = synth.CAMERA_ANGLE_X
    samples, targets, cam_pos, dirs = rrs.sample_camera_rays_batched(
        transform_matrices=poses, camera_angle_x=fov, imgs=imgs, number_of_rays=R, num_samples=S,
        delta_step=delta, even_spread=True, camera_ray=False, device="cpu")
    uv = po.even_spread_uv(3, R)
    d, t, _ = po.generate_rays(imgs.numpy(), poses.numpy(), fov, uv)
    check(f"{tag}: even-spread dirs bit-exact", np.array_equal(d, dirs.numpy()))
    check(f"{tag}: even-spread targets bit-exact", np.array_equal(t, targets.numpy()))
    o = np.repeat(poses[:, :3, 3].numpy(), uv.shape[1], axis=0)
    pos = po.sample_positions(o, d, S, delta)
    check(f"{tag}: even-spread samples bit-exact (S={S})", np.array_equal(pos.reshape(-1, 3), samples.numpy()))


def run_tv():
    """tv_loss (scripts/train.py:44-65) value and autograd vs the oracle's analytic gradient."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_train", os.path.join(REF, "scripts", "train.py"))
    ref_train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_train)
    for G in (6, 17):
        g = (synth.dense_grid(G)[:, : G - 1, : G - 2]).clone().requires_grad_(True)
        loss = ref_train.tv_loss(g)
        loss.backward()
        oloss, ograd = po.tv_loss(g.detach().numpy())
        check(f"tv G={G}: loss", abs(oloss - float(loss)) <= 1e-6 * float(loss), f"{oloss:.7f} vs {float(loss):.7f}")
        err = np.abs(ograd - g.grad.numpy()).max() / np.abs(g.grad.numpy()).max()
        check(f"tv G={G}: gradient <= 1e-6 rel", err <= 1e-6, f"rel err {err:.2e}")


if __name__ == "__main__":
    torch.manual_seed(0)
    run_tv()
    run_scene("c1-dense-nn", 64, 1, 64, 4096, 64, 6.0 / 64, "dense")
    run_scene("c1-ball-nn", 64, 4, 32, 256, 64, 6.0 / 64, "ball")
    run_scene("c2small-ball-nn", 128, 6, 40, 64, 600, 0.0125, "ball")
    run_scene("soft-nn", 32, 3, 16, 128, 96, 6.0 / 96, "soft")
    run_scene("c1-dense-tri", 64, 2, 32, 512, 64, 6.0 / 64, "dense", mode="trilinear")
    run_scene("ball-tri", 48, 3, 16, 128, 128, 6.0 / 128, "ball", mode="trilinear")
    run_even_spread("even", 64, 64, 4096, 64, 6.0 / 64)
    run_even_spread("even-S0", 64, 16, 9, 0, 0)
    print("FAILED:" if FAILED else "all checks passed", FAILED if FAILED else "")
    sys.exit(1 if FAILED else 0)
