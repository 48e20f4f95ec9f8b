"""The reference-named Python surface (src/*.py -> plenoxels_b200/*.py): host-side logic on CPU, and on the GPU the loop
body of the reference's fit() (scripts/train.py:104-191) written with the drop-in functions + torch.optim.Adam, checked
against the torch-CPU port of the reference."""
import os

import numpy as np
import pytest
import torch

import src.grid_functions as gf
import src.ray_sampling as rs
import src.rays_logic as rl
from oracle import torch_port as tp
from plenoxels_b200 import synth


def test_src_shim_exposes_the_reference_names():
    for name in ("generate_grid", "get_nearest_voxels", "average_pool3d_grid", "convolve_grid_to_remove_noise",
                 "trilinear_interpolation", "get_grid_points_indices", "find_out_of_bound", "fix_out_of_bounds"):
        assert callable(getattr(gf, name))
    for name in ("sample_camera_rays_batched", "normalize_samples_for_indecies", "compute_alpha_weighted_pixels",
                 "generate_rays_batched"):
        assert callable(getattr(rs, name))
    assert rl.compute_alpha_weighted_pixels is rs.compute_alpha_weighted_pixels
    import src.data_processing as dp
    import src.visualization as vz
    assert callable(dp.load_data) and callable(dp.load_image_data_from_path)
    assert callable(vz.visulize_3d_in_2d) and callable(vz.visulize_3d_in_2d_fast)


def test_boundary_signatures_equal_the_references():
    """the drop-in boundary is the set of Python signatures the unmodified scripts import by name.  Parameter
    names, order and defaults of every one of them against tests/golden/signatures.json (read off the reference's source with
    ast by tests/golden/make_golden.py).  `fit` may only ADD trailing parameters that have defaults."""
    import importlib
    import inspect
    import json
    table = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "signatures.json")))
    assert len(table) >= 20
    for qual, want in table.items():
        mod_name, fn_name = qual.rsplit(".", 1)
        mod = importlib.import_module("plenoxels_b200.fit" if mod_name == "scripts.train" else mod_name)
        params = list(inspect.signature(getattr(mod, fn_name)).parameters.values())
        got = [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default)] for p in params]
        norm = lambda d: None if d is None else d.replace('"', "'")
        want_n = [[n, norm(d)] for n, d in want]
        got_n = [[n, norm(d)] for n, d in got]
        if fn_name == "fit":
            assert got_n[:len(want_n)] == want_n and all(d is not None for _, d in got_n[len(want_n):]), qual
        else:
            assert got_n == want_n, (qual, got_n, want_n)


def test_grid_coords_metadata_does_not_survive_in_place_edits():
    """GridCoords caches the grid origin; an in-place edit of the coordinates must drop the cache so that
    coords_origin() falls back to the real reduction instead of returning a stale origin."""
    import src.grid_functions as gf
    coords, _, _, grid_grid = gf.generate_grid(6, 6, 6, points_distance=0.5, info_size=4, device="cpu")
    from plenoxels_b200.grid_functions import coords_origin
    want = tuple(float(x) for x in coords.as_subclass(torch.Tensor).min(0)[0])
    assert coords_origin(coords) == want
    for edit in (lambda t: t.add_(5.0), lambda t: t.mul_(2.0), lambda t: t.__iadd__(1.0), lambda t: t.__setitem__(0, 9.0),
                 lambda t: t.copy_(torch.zeros_like(t.as_subclass(torch.Tensor)))):
        c = coords.clone()
        assert coords_origin(c) == want                       # clone keeps the (still valid) cache
        edit(c)
        real = tuple(float(x) for x in c.as_subclass(torch.Tensor).reshape(-1, 3).min(0)[0])
        assert coords_origin(c) == real, "stale cached origin after an in-place edit"
    sl = grid_grid[1::2]
    sl += 1.0
    assert coords_origin(sl.reshape(-1, 3)) == tuple(float(x) for x in sl.as_subclass(torch.Tensor).reshape(-1, 3).min(0)[0])


@pytest.mark.parametrize("G,pd", [(8, 0.1), (93, 0.0125), (20, 0.05)])
def test_generate_grid_layout_and_free_origin(G, pd):
    """4-tuple of src/grid_functions.py:184-217; the coordinate tensors know their minimum without a reduction,
    also after the slicing + reshape fit() applies while growing the grid (scripts/train.py:112-116)."""
    coords, cells, mesh, grid_grid = gf.generate_grid(G, G, G + 1, points_distance=pd, info_size=4, device="cpu")
    assert coords.shape == (G * G * (G + 1), 3) and coords.dtype == torch.float32
    assert cells.shape == (G, G, G + 1, 4) and cells.requires_grad and float(cells.abs().sum()) == 0.0
    assert mesh.shape == (G, G, G + 1, 3) and mesh.dtype == torch.int64
    assert grid_grid.shape == (G, G, G + 1, 3)
    ref = tp.cell_centres((G, G, G + 1), pd)
    assert torch.equal(grid_grid.as_subclass(torch.Tensor), ref)
    assert gf.coords_origin(coords) == tuple(ref.reshape(-1, 3).min(0)[0].tolist())
    for start, stride in ((1, 2), (3, 5), (0, 1)):
        if start >= G:
            continue
        sl = grid_grid[start::stride, start::stride, start::stride].reshape(-1, 3)
        assert isinstance(sl, gf.GridCoords)
        assert gf.coords_origin(sl) == tuple(sl.as_subclass(torch.Tensor).min(0)[0].tolist())
    # anything else falls back to the real reduction
    assert not isinstance(coords * 2, gf.GridCoords)
    assert gf.coords_origin(coords * 2) == tuple((ref.reshape(-1, 3) * 2).min(0)[0].tolist())


def test_index_helpers_match_reference_semantics():
    ns = torch.tensor([[0.2, 1.5, -0.5], [3.7, -1.2, 2.0], [7.9, 7.0, 8.1]])
    grid = torch.zeros(8, 8, 8, 4)
    pts = gf.get_grid_points_indices(ns)
    assert pts.shape == (3, 8, 3) and pts.dtype == torch.int64
    assert pts[0, 0].tolist() == [1, 2, 0] and pts[0, 7].tolist() == [0, 1, -1]          # [ccc] ... [fff]
    assert gf.find_out_of_bound(ns, grid).tolist() == [False, False, False]
    assert gf.find_out_of_bound(torch.tensor([[0.0, 7.9, 3.0]]), grid).tolist() == [True]
    idx = torch.tensor([[-1, 8, 3], [9, -9, 0]])
    i0, i1, i2 = gf.fix_out_of_bounds(idx, grid)
    assert idx.tolist() == [[7, 0, 3], [1, 7, 0]], "wrap is in place, python-style modulo"


@pytest.mark.gpu
@pytest.mark.parametrize("pooled", [False, True])
def test_fit_loop_body_on_dropin_functions_matches_reference_port(plx_lib, monkeypatch, pooled):
    """scripts/train.py:130-184 with the drop-in functions, autograd and torch.optim.Adam on the GPU vs the torch-CPU port
    of the reference; `pooled` routes the grid through average_pool3d_grid (channel-planar, strided) like :110-118."""
    dev = "cuda:0"
    G, C, H, R, S = 24, 3, 16, 64, 48
    pd, delta, lr = synth.GRID_EXTENT / G, 6.0 / S, 0.0075
    grid0, poses, imgs = synth.soft_grid(G), synth.lookat_poses(C), synth.random_images(C, H, H)
    grid_indices, grid_cells_full, _, grid_grid = gf.generate_grid(G, G, G, points_distance=pd, info_size=4, device=dev)
    with torch.no_grad():
        grid_cells_full.copy_(grid0.to(dev))
    opt = torch.optim.Adam([grid_cells_full], lr=lr)
    port = tp.ReferenceStep(grid0, pd, poses, synth.CAMERA_ANGLE_X, imgs, R, S, delta, lr)
    T, im = poses.to(dev), imgs.to(dev)
    rf, stride = 3, 1
    for step in range(3):
        uv = synth.random_uv(C, R, seed=70 + step)
        monkeypatch.setattr(torch, "rand", lambda *a, **k: uv.clone().to(k.get("device", "cpu")))
        if pooled:
            start = int(rf / 2)
            gi = grid_grid[start::stride, start::stride, start::stride].reshape(-1, 3)
            cells = gf.average_pool3d_grid(grid_cells_full, receptive_field_size=rf, stride=stride)
            cur_pd = pd * stride
        else:
            gi, cells, cur_pd = grid_indices, grid_cells_full, pd
        samples, targets, cam_pos, dirs = rs.sample_camera_rays_batched(
            transform_matrices=T, camera_angle_x=synth.CAMERA_ANGLE_X, imgs=im, number_of_rays=R, num_samples=S,
            delta_step=delta, even_spread=False, camera_ray=False, device=dev)
        ns = rs.normalize_samples_for_indecies(gi, samples, cur_pd)
        nearest, mask = gf.get_nearest_voxels(ns, cells.clip(0, 1))
        nearest = nearest * mask.unsqueeze(-1)
        pix = rs.compute_alpha_weighted_pixels(nearest.reshape(C, R, S, 4)).reshape(-1, 4)
        loss = torch.nn.functional.mse_loss(pix, targets)
        opt.zero_grad()
        loss.backward()
        monkeypatch.undo()
        if pooled:
            # reference side of the same pooled step
            pg = port.grid
            pc = torch.nn.functional.avg_pool3d(pg.permute(3, 0, 1, 2).unsqueeze(0), (rf,) * 3, stride=stride).squeeze().permute(1, 2, 3, 0)
            centres = tp.cell_centres((G, G, G), pd)[start::stride, start::stride, start::stride].reshape(-1, 3)
            d_ref, t_ref = tp.rays_from_uv(imgs, poses, synth.CAMERA_ANGLE_X, uv.clone())
            pos = tp.place_samples(poses[:, :3, 3], d_ref, R, S, delta)
            v_ref, m_ref = tp.nearest_lookup((pos - centres.min(0)[0]) / (pd * stride), pc.clip(0, 1))
            pix_ref = tp.composite((v_ref * m_ref.unsqueeze(-1)).reshape(C, R, S, 4)).reshape(-1, 4)
            loss_ref = torch.nn.functional.mse_loss(pix_ref, t_ref)
            port.opt.zero_grad()
            loss_ref.backward()
            port.opt.step()
        else:
            loss_ref = port.step(uv)
        assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
        g, gr = grid_cells_full.grad.cpu(), port.grid.grad
        assert float((g - gr).abs().max()) <= 1e-5 * float(gr.abs().max())
        opt.step()
    diff = (grid_cells_full.detach().cpu() - port.grid.detach()).abs()
    assert float(torch.quantile(diff.flatten(), 0.999)) <= 1e-5


REFERENCE = "/root/reference"


@pytest.mark.skipif(not __import__("os").path.isdir(REFERENCE), reason="reference tree only exists in the build container")
def test_reference_scripts_import_unmodified_against_this_repo(tmp_path):
    """`scripts/train.py` / `compare_inference_to_image.py` / `main.py` of the UNMODIFIED reference resolve every `src.*` import
    to this repo (PYTHONPATH order), including `src.rays_logic`, which the reference itself lacks.
    matplotlib / plotly are GUI-only and absent from the image, so empty stand-ins are put on the path."""
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stubs = tmp_path / "stubs"
    for mod in ("matplotlib", "plotly"):
        (stubs / mod).mkdir(parents=True)
    (stubs / "matplotlib" / "__init__.py").write_text("")
    (stubs / "matplotlib" / "pyplot.py").write_text("")
    (stubs / "plotly" / "__init__.py").write_text("")
    (stubs / "plotly" / "graph_objects.py").write_text("")
    (stubs / "plotly" / "io.py").write_text("")
    (stubs / "plotly" / "express.py").write_text("")
    (stubs / "plotly" / "graph_objs.py").write_text("")
    code = ("import scripts.train as t, scripts.compare_inference_to_image as c, src.grid_functions as g;"
            "import scripts.main as m, scripts.visulize_grid as vg, scripts.visulize_camera_and_grid as vc;"
            "assert m.fit is t.fit and vg.convolve_grid_to_remove_noise is g.convolve_grid_to_remove_noise;"
            "import inspect, sys;"
            "assert g.__file__.startswith(%r), g.__file__;"
            "assert t.__file__.startswith(%r), t.__file__;"
            "assert t.get_nearest_voxels is g.get_nearest_voxels;"
            "print(list(inspect.signature(t.fit).parameters)[:4])") % (repo, REFERENCE)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([repo, str(stubs), REFERENCE]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    assert "gridsize" in out.stdout


@pytest.mark.gpu
def test_lazy_bridge_fuses_and_falls_back(plx_lib, monkeypatch):
    """The unfused reference call sequence returns lazy handles and ends in ONE fused march; an unrecognised op (the
    beta-loss slice of scripts/train.py:172) materialises through the eager kernels; both equal the fully eager path."""
    from plenoxels_b200 import lazy
    dev = "cuda:0"
    G, C, H, R, S = 20, 3, 16, 48, 40
    pd, delta, beta = synth.GRID_EXTENT / G, 6.0 / S, 5e-3
    poses, imgs = synth.lookat_poses(C).to(dev), synth.random_images(C, H, H).to(dev)
    uv = synth.random_uv(C, R, seed=5)
    grid0 = synth.soft_grid(G).to(dev)
    results = {}
    for mode in ("lazy", "eager"):
        monkeypatch.setenv("PLX_LAZY", "1" if mode == "lazy" else "0")
        monkeypatch.setattr(torch, "rand", lambda *a, **k: uv.clone().to(k.get("device", "cpu")))
        gi, cells, _, _ = gf.generate_grid(G, G, G, points_distance=pd, info_size=4, device=dev)
        with torch.no_grad():
            cells.copy_(grid0)
        samples, targets, cam_pos, dirs = rs.sample_camera_rays_batched(
            transform_matrices=poses, camera_angle_x=synth.CAMERA_ANGLE_X, imgs=imgs, number_of_rays=R, num_samples=S,
            delta_step=delta, even_spread=False, camera_ray=False, device=dev)
        ns = rs.normalize_samples_for_indecies(gi, samples, pd)
        nearest, mask = gf.get_nearest_voxels(ns, cells.clip(0, 1))
        nearest = nearest * mask.unsqueeze(-1)
        nearest = nearest.reshape(C, R, S, 4)
        if mode == "lazy":
            assert all(isinstance(t, lazy.LazyTensor) for t in (samples, ns, nearest, mask))
            assert nearest._spec["masked"] and nearest._real is None and tuple(nearest.shape) == (C, R, S, 4)
        pix = rs.compute_alpha_weighted_pixels(nearest)
        assert not isinstance(pix, lazy.LazyTensor) and tuple(pix.shape) == (C, R, 4)
        loss = torch.nn.functional.mse_loss(pix.reshape(-1, 4), targets)
        eps = 1e-4
        a = nearest[:, :, :, -1]                                       # unrecognised op -> materialises
        assert not isinstance(a, lazy.LazyTensor)
        loss = loss + beta * (torch.log(a + eps) - torch.log(1 - a + eps)).mean()
        loss.backward()
        monkeypatch.undo()
        results[mode] = (pix.detach().cpu(), float(loss), cells.grad.detach().cpu(),
                         samples.materialize().cpu() if mode == "lazy" else samples.cpu())
    assert torch.equal(results["lazy"][3], results["eager"][3]), "materialised sample positions are the eager ones"
    assert float((results["lazy"][0] - results["eager"][0]).abs().max()) <= 2e-6
    assert abs(results["lazy"][1] - results["eager"][1]) <= 1e-6 * abs(results["eager"][1])
    assert float((results["lazy"][2] - results["eager"][2]).abs().max()) <= 2e-6 * float(results["eager"][2].abs().max())


def test_dataset_loader_roundtrip(tmp_path):
    """load_image_data_from_path / load_data (src/data_processing.py:18-60) on a synthetic NeRF-format dataset; compared with
    the reference's loader when the reference tree is present."""
    import json
    import os
    from PIL import Image
    import src.data_processing as dp
    (tmp_path / "train").mkdir()
    poses = synth.lookat_poses(3)
    rng = np.random.default_rng(0)
    frames, pix = [], []
    for i in range(3):
        a = (rng.random((8, 8, 4)) * 255).astype(np.uint8)
        pix.append(a)
        Image.fromarray(a, "RGBA").save(tmp_path / "train" / f"r_{i}.png")
        frames.append({"file_path": f"./train/r_{i}", "rotation": 0.1, "transform_matrix": poses[i].tolist()})
    (tmp_path / "transforms_train.json").write_text(json.dumps({"camera_angle_x": 0.69, "frames": frames}))
    data, imgs = dp.load_image_data_from_path(str(tmp_path / "train"), str(tmp_path / "transforms_train.json"))
    T, paths, fov = dp.load_data(data)
    assert imgs.shape == (3, 8, 8, 4) and imgs.dtype == torch.float32
    assert torch.equal(imgs, torch.tensor(np.stack(pix), dtype=torch.float) / 255)
    assert torch.equal(T, poses) and fov == 0.69 and paths == [f["file_path"] for f in frames]
    if os.path.isdir(REFERENCE):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_dp", os.path.join(REFERENCE, "src", "data_processing.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        rdata, rimgs = ref.load_image_data_from_path(str(tmp_path / "train"), str(tmp_path / "transforms_train.json"))
        rT, rpaths, rfov = ref.load_data(rdata)
        assert torch.equal(imgs, rimgs) and torch.equal(T, rT) and paths == rpaths and fov == rfov


@pytest.mark.gpu
def test_fused_fit_with_progressive_growing_matches_reference_port(plx_lib, monkeypatch):
    """plenoxels_b200.fit.GridFitter (pooling -> fused march with beta -> TV -> pooling backward -> Adam) against the
    reference's loop body with the same schedule restated on torch-CPU (scripts/train.py:104-184): losses, gradients and the
    grid after 12 steps that cross two pooling windows and the full-resolution regime."""
    from plenoxels_b200.fit import GridFitter, receptive_field_at
    G, C, H, R, S = 32, 3, 16, 48, 40
    pd, delta, lr, tv, beta = synth.GRID_EXTENT / G, 6.0 / S, 0.0075, 1e-5, 5e-3
    schedule = [9, 5]
    poses, imgs = synth.lookat_poses(C), synth.random_images(C, H, H)
    init = synth.soft_grid(G)                      # a zero grid would make tv_loss's gradient 0/0 in the reference
    ft = GridFitter([G, G, G], pd, poses, synth.CAMERA_ANGLE_X, imgs, R, S, delta, lr, tv=tv, beta=beta, schedule=schedule,
                    device="cuda:0")
    ft.grid.copy_(init.cuda())
    ref_grid = init.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_grid], lr=lr)
    centres_full = tp.cell_centres((G, G, G), pd)
    eps = 1e-4
    for i in range(12):
        uv = synth.random_uv(C, R, seed=300 + i)
        monkeypatch.setattr(torch, "rand", lambda *a, **k: uv.clone().to(k.get("device", "cpu")))
        mse, tvl = ft.step(i)
        monkeypatch.undo()
        gpu_grad = ft.grad_abs_sum.clone()
        # ---- reference loop body on the CPU
        rf = receptive_field_at(i, schedule)
        if rf > 1:
            stride, start = max(1, rf // 4), int(rf / 2)
            cells = torch.nn.functional.avg_pool3d(ref_grid.permute(3, 0, 1, 2).unsqueeze(0), (rf,) * 3, stride=stride).squeeze().permute(1, 2, 3, 0)
            centres = centres_full[start::stride, start::stride, start::stride].reshape(-1, 3)
            cur_pd = pd * stride
        else:
            cells, centres, cur_pd = ref_grid, centres_full.reshape(-1, 3), pd
        d_ref, t_ref = tp.rays_from_uv(imgs, poses, synth.CAMERA_ANGLE_X, uv.clone())
        pos = tp.place_samples(poses[:, :3, 3], d_ref, R, S, delta)
        v_ref, m_ref = tp.nearest_lookup((pos - centres.min(0)[0]) / cur_pd, cells.clip(0, 1))
        nearest = (v_ref * m_ref.unsqueeze(-1)).reshape(C, R, S, 4)
        pix = tp.composite(nearest).reshape(-1, 4)
        mse_ref = torch.nn.functional.mse_loss(pix, t_ref)
        gx = cells[:, :-1] - cells[:, 1:]
        gy = cells[:, :, :-1] - cells[:, :, 1:]
        gz = cells[:-1] - cells[1:]
        tv_ref = tv * torch.sqrt(gx.pow(2).sum() + gy.pow(2).sum() + gz.pow(2).sum())            # scripts/train.py:44-65
        a = nearest[:, :, :, -1]
        beta_ref = beta * (torch.log(a + eps) - torch.log(1 - a + eps)).mean()                    # :170-177
        opt.zero_grad()
        (mse_ref + tv_ref + beta_ref).backward()
        assert abs(float(mse) - float(mse_ref)) <= 2e-5 * abs(float(mse_ref)), f"step {i}"
        assert abs(float(tvl) - float(tv_ref)) <= 2e-5 * abs(float(tv_ref)), f"step {i}"
        if i == 0:
            g0 = ref_grid.grad.abs()
            assert float((gpu_grad.cpu() - g0).abs().max()) <= 1e-5 * float(g0.max())
        opt.step()
    diff = (ft.grid.cpu() - ref_grid.detach()).abs()
    assert float(torch.quantile(diff.flatten(), 0.999)) <= 2e-5, float(torch.quantile(diff.flatten(), 0.999))


@pytest.mark.gpu
def test_fit_function_end_to_end(plx_lib, tmp_path):
    """fit() with the reference's argument list on a synthetic NeRF-format dataset: runs, logs, saves the reference's .pth."""
    import json
    from PIL import Image
    from plenoxels_b200.fit import fit
    (tmp_path / "train").mkdir()
    poses = synth.lookat_poses(4)
    rng = np.random.default_rng(1)
    frames = []
    for i in range(4):
        Image.fromarray((rng.random((16, 16, 4)) * 255).astype(np.uint8), "RGBA").save(tmp_path / "train" / f"r_{i}.png")
        frames.append({"file_path": f"./train/r_{i}", "rotation": 0.0, "transform_matrix": poses[i].tolist()})
    (tmp_path / "transforms_train.json").write_text(json.dumps({"camera_angle_x": synth.CAMERA_ANGLE_X, "frames": frames}))
    save = tmp_path / "grid_cells_trained.pth"
    ft = fit([96, 96, 96], synth.GRID_EXTENT / 96, 32, 64, 6.0 / 64, 0.0075, 1e-5, 5e-3, 8, False, str(tmp_path / "train"),
             str(tmp_path / "transforms_train.json"), str(save), "cuda:0", progressive_growing=False, log_every=4)
    ck = torch.load(save)
    assert set(ck) == {"grid", "grid_grad", "param"} and ck["grid"].shape == (96, 96, 96, 4)
    assert ck["param"]["gridsize"] == [96, 96, 96] and ck["param"]["number_of_rays"] == 32
    assert float(ck["grid_grad"].sum()) > 0 and ft.steps_done == 8
    # the reference's own fit() raises when the grid is smaller than the first pooling window (93): so do we
    with pytest.raises(RuntimeError):
        fit([64, 64, 64], 0.05, 8, 16, 0.1, 0.01, 0, 0, 1, False, str(tmp_path / "train"), str(tmp_path / "transforms_train.json"),
            str(save), "cuda:0")
