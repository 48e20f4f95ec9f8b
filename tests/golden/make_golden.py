"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on small seeded inputs.

Build-container only (the GPU box has no /root/reference):   python tests/golden/make_golden.py
The fixtures hold both the inputs and the reference's outputs, so the checks in tests/ need neither the reference nor
torch's RNG to reproduce them.  The reference ships no golden vectors of its own; these are its outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PLX_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(1, REF)
# `src.*` must resolve to the REFERENCE here.  The reference's src/ has no __init__.py (a namespace package), so this
# repo's src/ shim (a regular package) would always win the import; pin the name to the reference directory instead.
import types  # noqa: E402
_ref_src = types.ModuleType("src")
_ref_src.__path__ = [os.path.join(REF, "src")]
sys.modules["src"] = _ref_src

import src.grid_functions as rgf      # noqa: E402  reference
import src.ray_sampling as rrs        # noqa: E402  reference
from plenoxels_b200 import synth      # noqa: E402

assert rgf.__file__.startswith(REF)

# matplotlib / plotly are GUI-only imports of src/visualization.py and absent from this image: empty stand-ins
GUI_STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d", "plotly",
             "plotly.graph_objects", "plotly.io", "plotly.express", "plotly.graph_objs")


def reference_step(grid, pd, poses, fov, imgs, R, S, delta, uv, mode, even_spread=False):
    """scripts/train.py:130-157 + :181 with the reference's own functions; uv injected in place of torch.rand."""
    dims = grid.shape[:3]
    coords, _, _, _ = rgf.generate_grid(*dims, points_distance=pd, info_size=4, device="cpu")
    real_rand = torch.rand
    torch.rand = lambda *a, **k: uv.clone()
    try:
        samples, targets, cam_pos, dirs = rrs.sample_camera_rays_batched(
            transform_matrices=poses, camera_angle_x=fov, imgs=imgs, number_of_rays=R, num_samples=S, delta_step=delta,
            even_spread=even_spread, camera_ray=False, device="cpu")
    finally:
        torch.rand = real_rand
    R_eff = dirs.shape[0] // poses.shape[0]
    ns = rrs.normalize_samples_for_indecies(coords, samples, pd)
    g = grid.detach().clone().requires_grad_(True)
    if mode == "nearest":
        idx = torch.round(ns).to(torch.long)
        inb = rgf.find_out_of_bound(idx, g)
        lin = torch.where(inb, (idx[:, 0] * dims[1] + idx[:, 1]) * dims[2] + idx[:, 2], torch.full_like(idx[:, 0], -1))
        vals, inb2 = rgf.get_nearest_voxels(ns, g.clip(0, 1))
        assert torch.equal(inb, inb2)
        vals = vals * inb.unsqueeze(-1)
    else:
        inb = rgf.find_out_of_bound(ns, g)
        pts = rgf.get_grid_points_indices(ns)
        rgf.fix_out_of_bounds(pts.reshape(-1, 3), g)
        lin = torch.where(inb, torch.zeros_like(inb, dtype=torch.long), torch.full_like(inb, -1, dtype=torch.long))
        vals = rgf.trilinear_interpolation(ns, pts, g.clip(0, 1)) * inb.unsqueeze(-1)
    pix = rrs.compute_alpha_weighted_pixels(vals.reshape(poses.shape[0], R_eff, S, 4)).reshape(-1, 4)
    loss = torch.nn.functional.mse_loss(pix, targets)
    loss.backward()
    return dict(gmin=coords.min(0)[0].numpy(), dirs=dirs.numpy(), targets=targets.numpy(),
                lin=lin.reshape(-1, S).numpy().astype(np.int32), inb=inb.reshape(-1, S).numpy(),
                vals=vals.detach().numpy().astype(np.float32), pix=pix.detach().numpy(), loss=np.float64(loss.item()),
                grad=g.grad.numpy())


def make_case(name, G, C, H, R, S, delta, kind, mode, seed):
    torch.manual_seed(seed)
    pd = synth.GRID_EXTENT / G
    grid = {"ball": synth.ball_grid, "dense": synth.dense_grid, "soft": synth.soft_grid}[kind](G, seed=seed)
    poses, imgs, uv = synth.lookat_poses(C), synth.random_images(C, H, H, seed=seed + 1), synth.random_uv(C, R, seed=seed + 2)
    out = reference_step(grid, pd, poses, synth.CAMERA_ANGLE_X, imgs, R, S, delta, uv, mode)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), grid=grid.numpy(), poses=poses.numpy(), imgs=imgs.numpy(),
                        uv=uv.numpy(), fov=np.float64(synth.CAMERA_ANGLE_X), pd=np.float64(pd), delta=np.float64(delta),
                        S=np.int64(S), R=np.int64(R), mode=np.array(mode), **out)
    print(name, "in-bounds", float(out["inb"].mean()), "loss", float(out["loss"]))


def make_even_spread(name):
    """even_spread=True ray lattice (src/ray_sampling.py:220-223) incl. S = 0 (scripts/visulize_camera_and_grid.py:36-46)."""
    poses, imgs = synth.lookat_poses(2), synth.random_images(2, 24, 24, seed=7)
    samples, targets, cam_pos, dirs = rrs.sample_camera_rays_batched(
        transform_matrices=poses, camera_angle_x=synth.CAMERA_ANGLE_X, imgs=imgs, number_of_rays=100, num_samples=5,
        delta_step=0.3, even_spread=True, camera_ray=False, device="cpu")
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), poses=poses.numpy(), imgs=imgs.numpy(),
                        fov=np.float64(synth.CAMERA_ANGLE_X), dirs=dirs.numpy(), targets=targets.numpy(),
                        samples=samples.numpy(), cam_pos=cam_pos.numpy())
    print(name, dirs.shape)


def make_adam(name):
    """Three torch.optim.Adam steps (scripts/train.py:89,:182) on a small parameter vector + |grad| accumulation (:184)."""
    torch.manual_seed(5)
    n = 4096
    p0 = torch.rand(n) * 1.4 - 0.2
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=0.0075)
    grads, params = [], []
    gabs = torch.zeros(n)
    for _ in range(3):
        g = torch.randn(n) * 1e-3 * (torch.rand(n) < 0.6)
        p.grad = g.clone()
        opt.step()
        gabs += g.abs()
        grads.append(g.numpy().copy())
        params.append(p.detach().numpy().copy())
    st = opt.state[p]
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), p0=p0.numpy(), grads=np.stack(grads), params=np.stack(params),
                        exp_avg=st["exp_avg"].numpy(), exp_avg_sq=st["exp_avg_sq"].numpy(), gabs=gabs.numpy(),
                        lr=np.float64(0.0075))
    print(name)


def make_tv(name):
    """tv_loss value + autograd gradient (scripts/train.py:44-65) on a small non-cubic grid."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_train", os.path.join(REF, "scripts", "train.py"))
    ref_train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_train)
    g = synth.dense_grid(12, seed=21)[:, :11, :9].clone().requires_grad_(True)
    loss = ref_train.tv_loss(g)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), grid=g.detach().numpy(), loss=np.float64(loss.item()), grad=g.grad.numpy())
    print(name, float(loss))


def make_inference(name):
    """visulize_3d_in_2d (src/visualization.py:111-154): ray-marched uint8 image of one camera, with the alpha threshold.
    matplotlib / plotly are GUI-only imports of that module and absent here: empty stand-ins are injected."""
    for mod in GUI_STUBS:
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import src.visualization as rvz
    G, res, S = 24, 12, 80
    pd = synth.GRID_EXTENT / G
    grid = synth.ball_grid(G, seed=31)
    poses = synth.lookat_poses(5)[2:3]
    imgs = synth.random_images(1, 16, 16, seed=32)
    coords, _, _, _ = rgf.generate_grid(G, G, G, points_distance=pd, info_size=4, device="cpu")
    ck = {"grid": grid, "param": {"points_distance": pd, "delta_step": 6.0 / S}}
    img = rvz.visulize_3d_in_2d(ck, poses, synth.CAMERA_ANGLE_X, imgs, coords, True, 0.2, res * res, S, device="cpu")
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), grid=grid.numpy(), poses=poses.numpy(), imgs=imgs.numpy(),
                        pd=np.float64(pd), delta=np.float64(6.0 / S), S=np.int64(S), res=np.int64(res), threshold=np.float64(0.2),
                        fov=np.float64(synth.CAMERA_ANGLE_X), image=img)
    print(name, img.shape, img.dtype, int(img[..., 3].max()))


def make_splat(name):
    """visulize_3d_in_2d_fast (src/visualization.py:157-232), the CPU point-splat preview scripts/compare_inference_to_image.py:58
    calls: the (xs, ys, 3) float64 image of one camera over a clipped, thresholded grid."""
    for mod in GUI_STUBS:
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import src.visualization as rvz
    G, size_y = 40, 96
    pd = synth.GRID_EXTENT / G
    grid = synth.ball_grid(G, seed=41).clip(0.0, 1.0)
    grid[..., 3][grid[..., 3] < 0.2] = 0.0
    pose = synth.lookat_poses(7)[3]
    img = rvz.visulize_3d_in_2d_fast(grid, pd, pose, synth.CAMERA_ANGLE_X, size_y)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), grid=grid.numpy(), pose=pose.numpy(), pd=np.float64(pd),
                        fov=np.float64(synth.CAMERA_ANGLE_X), size_y=np.int64(size_y), image=img)
    print(name, img.shape, img.dtype, "painted pixels", int((img != 1.0).any(-1).sum()))


def make_signatures(name):
    """The drop-in boundary frozen as text: `inspect.signature` of every reference function the scripts import by
    name, read from the reference's source with ast (importing src.visualization would need the GUI toolkits)."""
    import ast
    import json
    wanted = {
        "src/grid_functions.py": ["generate_grid", "get_nearest_voxels", "average_pool3d_grid", "convolve_grid_to_remove_noise",
                                  "trilinear_interpolation", "get_grid_points_indices", "find_out_of_bound", "fix_out_of_bounds",
                                  "collect_cell_information_via_indices"],
        "src/ray_sampling.py": ["sample_camera_rays_batched", "normalize_samples_for_indecies", "compute_alpha_weighted_pixels",
                                "generate_rays_batched"],
        "src/data_processing.py": ["load_data", "load_image_data_from_path", "load_image_data", "read_data", "get_data_from_index"],
        "src/visualization.py": ["visulize_3d_in_2d", "visulize_3d_in_2d_fast", "visualize_rays_3d"],
        "scripts/train.py": ["fit"],
    }
    out = {}
    for rel, names in wanted.items():
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                a = node.args
                pos = [x.arg for x in a.posonlyargs + a.args]
                defaults = [None] * (len(pos) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
                out[f"{rel[:-3].replace('/', '.')}.{node.name}"] = [[n, d] for n, d in zip(pos, defaults)]
    with open(os.path.join(HERE, f"{name}.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(name, len(out), "functions")


from tests.helpers import FIT_ARGS, write_fit_dataset      # noqa: E402  (shared with the GPU test that replays the fit)


def make_fit(name):
    """The UNMODIFIED scripts/train.py::fit (:68-210) on the CPU: 8 synthetic PNGs, 128^3 grid, 240 steps = the whole
    progressive-growing schedule (46 pooling windows x 5 steps) plus 10 full-resolution steps, train.py's default tv / beta,
    the uv stream of torch.manual_seed(123).  Frozen: a strided subset of the saved grid / grid_grad plus whole-array sums."""
    import tempfile
    for mod in GUI_STUBS + ("tqdm",):
        if mod == "tqdm":
            continue
        sys.modules.setdefault(mod, types.ModuleType(mod))
    sys.path.insert(1, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_train", os.path.join(REF, "scripts", "train.py"))
    ref_train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_train)
    with tempfile.TemporaryDirectory() as tmp:
        path, tpath = write_fit_dataset(tmp)
        save = os.path.join(tmp, "grid.pth")
        torch.manual_seed(123)
        ref_train.fit(path=path, transform_path=tpath, save_path=save, device="cpu", progressive_growing=True, **FIT_ARGS)
        ck = torch.load(save)
    g, gg = ck["grid"].numpy(), ck["grid_grad"].numpy()
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), grid_subset=g[::4, ::4, ::4].copy(), grad_subset=gg[::4, ::4, ::4].copy(),
                        grid_sum=g.astype(np.float64).reshape(-1, 4).sum(0), grid_abs_sum=np.abs(g).astype(np.float64).reshape(-1, 4).sum(0),
                        grad_sum=gg.astype(np.float64).reshape(-1, 4).sum(0), grid_max=np.float64(np.abs(g).max()),
                        grad_max=np.float64(gg.max()), occupied=np.int64((g[..., 3] > 0.1).sum()), seed=np.int64(123),
                        param=np.array(repr({k: v for k, v in ck["param"].items()})))
    print(name, "grid |max|", float(np.abs(g).max()), "cells with alpha > 0.1:", int((g[..., 3] > 0.1).sum()))


def sha(a: np.ndarray) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_fullsize(name, scene):
    """The reference at a BASELINE config's FULL size (config #2: 12 800 rays x 600 samples through a 128^3 grid; #3: 4096 x
    256 through 256^3), frozen compactly: the inputs are regenerated from their seeds (their SHA-256 is stored), the
    7.68 M / 1.05 M linear indices are stored as a SHA-256, the gradient as sums plus a strided subset."""
    sc = synth.make_scene(scene, H=8)
    C_, R, S = sc.poses.shape[0], sc.rays_per_cam, sc.num_samples
    uv = synth.random_uv(C_, R, seed=77)
    out = reference_step(sc.grid, sc.points_distance, sc.poses, sc.fov, sc.imgs, R, S, sc.delta_step, uv, "nearest")
    grad = out["grad"].astype(np.float64)
    np.savez_compressed(
        os.path.join(HERE, f"{name}.npz"), scene=np.array(scene), uv_seed=np.int64(77), H=np.int64(8),
        sha_grid=np.array(sha(sc.grid.numpy())), sha_uv=np.array(sha(uv.numpy())), sha_poses=np.array(sha(sc.poses.numpy())),
        sha_imgs=np.array(sha(sc.imgs.numpy())), gmin=out["gmin"], sha_dirs=np.array(sha(out["dirs"])),
        sha_targets=np.array(sha(out["targets"])), sha_lin=np.array(sha(out["lin"])),
        count=out["inb"].sum(1).astype(np.int32), pix=out["pix"], loss=out["loss"],
        grad_max=np.float64(np.abs(grad).max()), grad_sum=grad.reshape(-1, 4).sum(0), grad_abs_sum=np.abs(grad).reshape(-1, 4).sum(0),
        grad_nonzero_cells=np.int64((np.abs(grad).sum(-1) > 0).sum()), grad_subset=out["grad"][::7, ::5, ::3].copy())
    print(name, "rays", C_ * R, "samples", S, "in-bounds", float(out["inb"].mean()), "loss", float(out["loss"]),
          "nonzero gradient cells", int((np.abs(grad).sum(-1) > 0).sum()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "full":            # only the full-size fixtures (slow: the reference on M = 7.7 M samples)
        make_fullsize("full_c2", "c2")
        make_fullsize("full_c3", "c3")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fit":             # the reference's whole fit() on the CPU (a few minutes)
        make_fit("fit_g128")
        sys.exit(0)
    make_inference("inference_g24")
    make_splat("splat_g40")
    make_signatures("signatures")
    make_tv("tv_g12")
    make_case("nn_dense_g24", 24, 2, 8, 64, 48, 6.0 / 48, "dense", "nearest", seed=11)
    make_case("nn_ball_g32", 32, 3, 12, 48, 96, 6.0 / 96, "ball", "nearest", seed=12)
    make_case("tri_ball_g24", 24, 2, 8, 48, 64, 6.0 / 64, "ball", "trilinear", seed=13)
    make_case("tri_dense_g16", 16, 2, 8, 32, 40, 6.0 / 40, "dense", "trilinear", seed=14)
    make_even_spread("even_spread")
    make_adam("adam3")
