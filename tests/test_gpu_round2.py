"""Round-2 parity: trilinear inside the fused training march, full-size checks of BASELINE configs #2 (trilinear) and #4
(512^3 inference, strided rays) against the C oracle, the uint8 image epilogue and the GPU point splat against what the
UNMODIFIED reference produced (tests/golden), uint8 target images, and the trainer's step() / step_host() agreement with the
TV term.  Everything goes through the C ABI (plenoxels_b200.ops / trainer)."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from plenoxels_b200 import _lib as L, ops, synth
from plenoxels_b200.trainer import VoxelTrainer
from tests.helpers import Case, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5                                      # BASELINE.json north_star: RGB, depth, loss, grid gradients within 1e-5 relative
GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------------------------------------ trilinear in the fused march
@pytest.mark.parametrize("G,C,H,R,S,kind,beta", [(24, 2, 8, 48, 64, "ball", 0.0), (16, 2, 8, 32, 40, "dense", 0.0),
                                                 (20, 3, 8, 64, 200, "soft", 0.0), (24, 2, 8, 48, 160, "ball", 5e-3)])
def test_fused_trilinear_march_matches_oracle_and_unfused_kernels(plx_lib, G, C, H, R, S, kind, beta):
    """K12 in trilinear mode (eight-corner gather, corner pass-mask bits, in-lane merge of samples sharing a floor cell,
    two-sided early termination) against the numpy oracle (src/grid_functions.py:7-44, :220-246 + scripts/train.py:146-157)
    and against the separate K1 + K2 kernels."""
    cs = Case(G, C, H, R, S, 6.0 / S, kind)
    d = cs.cuda()
    gg = torch.zeros_like(d["grid"])
    bom = beta / (cs.N * S)
    rgba, loss = ops.render_train(d["grid"], gg, S, cs.delta, cs.gmin, cs.pd, imgs=d["imgs"], poses=d["poses"], fov=cs.fov,
                                  uv=d["uv"], mode="trilinear", beta_over_m=bom)
    orgba, odepth, ocount, _ = cs.oracle_forward("trilinear")
    from oracle import plenoxel_oracle as po
    oloss, gpix = po.mse_loss(orgba, cs.targets)
    ograd = cs.oracle_backward(gpix, "trilinear", beta=beta)
    assert rel_err(rgba.cpu().numpy(), orgba) <= TOL
    assert abs(float(loss) - oloss) <= TOL * oloss
    assert rel_err(gg.cpu().numpy(), ograd) <= TOL
    # the unfused pair (autograd path) on the same rays
    g2 = d["grid"].clone().requires_grad_(True)
    pix = ops.render_rays(g2, d["origins"], d["dirs"], S, cs.delta, cs.gmin, cs.pd, mode="trilinear", rays_per_origin=R, beta_over_m=bom)
    torch.nn.functional.mse_loss(pix, d["targets"]).backward()
    assert rel_err(rgba.cpu().numpy(), pix.detach().cpu().numpy()) <= 2e-6
    assert rel_err(gg.cpu().numpy(), g2.grad.cpu().numpy()) <= 5e-6


def test_trilinear_training_at_config2_size_matches_the_c_oracle(plx_lib):
    """BASELINE config #2 in full — 12 800 rays x 600 samples through the 128^3 ball grid — with the trilinear lookup: pixels,
    depth, loss and the gradient of the fused march against the plain-C oracle on EVERY ray (1e-5), in-bounds counts bit-exact."""
    sc = synth.make_scene("c2", H=8)
    C_, R, S = sc.poses.shape[0], sc.rays_per_cam, sc.num_samples
    grid, poses = sc.grid.to(DEV), sc.poses.to(DEV)
    gmin = ops.grid_origin(sc.grid.shape, sc.points_distance)
    uv = synth.random_uv(C_, R, seed=78).to(DEV)
    dirs, _ = ops.generate_rays(None, poses, sc.fov, uv=uv, want_targets=False)
    targets = torch.rand(C_ * R, 4, device=DEV, generator=torch.Generator(device=DEV).manual_seed(6))
    o = np.repeat(sc.poses[:, :3, 3].numpy(), R, axis=0)
    rgba_o, depth_o, count_o, _ = co.render_forward(sc.grid.numpy(), o, dirs.cpu().numpy(), S, sc.delta_step, np.float32(gmin),
                                                    sc.points_distance, mode="trilinear", want_lin=False)
    rgba, depth, count = ops.render_rays(grid, poses[:, :3, 3], dirs, S, sc.delta_step, gmin, sc.points_distance, mode="trilinear",
                                         rays_per_origin=R, return_depth=True, return_count=True)
    assert np.array_equal(count.cpu().numpy(), count_o)
    assert rel_err(rgba.cpu().numpy(), rgba_o) <= TOL and rel_err(depth.cpu().numpy(), depth_o) <= TOL
    gg = torch.zeros_like(grid)
    rgba_f, loss_f = ops.render_train(grid, gg, S, sc.delta_step, gmin, sc.points_distance, origins=poses[:, :3, 3], dirs=dirs,
                                      targets=targets, rays_per_origin=R, mode="trilinear")
    loss_o, gpix = co.mse_loss(rgba_o, targets.cpu().numpy())
    grad_o = co.render_backward(sc.grid.numpy(), o, dirs.cpu().numpy(), S, sc.delta_step, np.float32(gmin), sc.points_distance, gpix,
                                mode="trilinear")
    assert rel_err(rgba_f.cpu().numpy(), rgba_o) <= TOL
    assert abs(float(loss_f) - loss_o) <= TOL * loss_o
    assert rel_err(gg.cpu().numpy(), grad_o) <= TOL


@pytest.mark.parametrize("scale", [1.0, 0.25])
def test_trilinear_rays_longer_than_the_value_cache_take_the_uncached_path(plx_lib, scale):
    """The fused trilinear march caches interpolated values per ray in shared memory, sized for unit directions through the
    grid's diagonal; a ray that visits more samples (here: directions scaled to a quarter length, and the cache forced down to
    one iteration with plx_tune) must run the uncached path and produce the same pixels, loss and gradient."""
    G, S = 24, 192
    cs = Case(G, 2, 8, 40, S, 6.0 / S, "ball")
    d = cs.cuda()
    dirs, targets = ops.generate_rays(d["imgs"], d["poses"], cs.fov, uv=d["uv"])
    o = d["poses"][:, :3, 3].repeat_interleave(cs.R, 0).contiguous()
    dirs = (dirs * scale).contiguous()
    delta = cs.delta / scale                  # same world-space sample spacing: same pixels as the unit-length rays when scale = 1
    lib = L.load()
    out = {}
    try:
        for cap in (0, 1):
            L.check(lib.plx_tune(b"train_cache_it", cap))
            gg = torch.zeros_like(d["grid"])
            rgba, loss = ops.render_train(d["grid"], gg, S, delta if scale != 1.0 else cs.delta, cs.gmin, cs.pd, origins=o, dirs=dirs,
                                          targets=targets, rays_per_origin=1, mode="trilinear")
            out[cap] = (rgba, float(loss), gg)
    finally:
        L.check(lib.plx_tune(b"train_cache_it", 0))
    assert float(out[0][2].abs().sum()) > 0
    assert torch.equal(out[0][0], out[1][0])
    assert abs(out[0][1] - out[1][1]) <= 1e-6 * out[0][1]          # summed with atomics: order differs from launch to launch
    assert rel_err(out[1][2].cpu().numpy(), out[0][2].cpu().numpy()) <= 1e-6
    # and both equal the separate forward + backward kernels
    grid = d["grid"].clone().requires_grad_(True)
    px = ops.render_rays(grid, o, dirs, S, delta if scale != 1.0 else cs.delta, cs.gmin, cs.pd, mode="trilinear", rays_per_origin=1)
    ((px - targets) ** 2).mean().backward()
    assert rel_err(out[0][0].cpu().numpy(), px.detach().cpu().numpy()) <= TOL
    assert rel_err(out[0][2].cpu().numpy(), grid.grad.cpu().numpy()) <= TOL


@pytest.mark.parametrize("mode,S,world", [("nearest", 200, 2), ("nearest", 64, 3), ("nearest", 600, 8), ("trilinear", 96, 4)])
def test_push_exchange_march_sends_every_cell_to_its_slab_owner(plx_lib, mode, S, world):
    """PlxPeerGrad on ONE device: `world` gradient buffers stand in for the ranks' peer-mapped ones.  Every buffer may only
    receive cells of its own slab (plx_slab_partition), and the slabs put together are the gradient the plain march produces
    (the cross-lane run merge of the push path only regroups the sums)."""
    import ctypes as C
    from plenoxels_b200.trainer import slab_partition
    cs = Case(32, 3, 8, 96, S, (0.5 if S == 600 else 1.0) * 6.0 / S, "ball")
    d = cs.cuda()
    kw = dict(imgs=d["imgs"], poses=d["poses"], fov=cs.fov, uv=d["uv"], mode=mode)
    ref = torch.zeros_like(d["grid"])
    rgba0, loss0 = ops.render_train(d["grid"], ref, S, cs.delta, cs.gmin, cs.pd, **kw)
    bufs = [torch.zeros_like(d["grid"]) for _ in range(world)]
    unused = torch.zeros_like(d["grid"])
    rgba1, loss1 = ops.render_train(d["grid"], unused, S, cs.delta, cs.gmin, cs.pd, peer_grads=bufs, **kw)
    assert torch.equal(rgba0, rgba1) and float(unused.abs().max()) == 0.0
    assert abs(float(loss0) - float(loss1)) <= 1e-6 * float(loss0)
    n_cells = cs.G ** 3
    total = torch.zeros_like(ref)
    for r, buf in enumerate(bufs):
        mul, b, e = slab_partition(n_cells, r, world)
        flat = buf.view(-1, 4)
        outside = flat.abs().sum(1)
        outside[b:e] = 0
        assert float(outside.max()) == 0.0, f"rank {r} received cells outside its slab"
        total.view(-1, 4)[b:e] = flat[b:e]
    assert float(ref.abs().max()) > 0
    assert rel_err(total.cpu().numpy(), ref.cpu().numpy()) <= 2e-6


# ------------------------------------------------------------------------------------------------ config #4 at full size
@pytest.mark.parametrize("mode", ["nearest", "trilinear"])
def test_config4_inference_frame_matches_the_c_oracle_on_strided_rays(plx_lib, mode):
    """BASELINE config #4: the 512^3 grid (alpha threshold 0.2, scripts/compare_inference_to_image.py:91), one 800x800 view,
    600 samples per ray, rendered by the coherent ray-packet kernel; every 89th ray of the 640 000 (7 192 rays, all image
    regions) is restated by the C oracle: in-bounds counts bit-exact, pixels and depth within 1e-5, and the uint8 image the
    march kernel writes equals the reference's post-processing of those pixels."""
    G, S, delta, side = 512, 600, 0.01, 800
    pd = synth.GRID_EXTENT / G
    grid_h = synth.ball_grid(G).clip_(0.0, 1.0)
    grid_h[..., 3][grid_h[..., 3] < 0.2] = 0.0
    grid = grid_h.to(DEV)
    pose = synth.lookat_poses(4)[1:2]
    gmin = ops.grid_origin(grid.shape, pd)
    n = side * side
    dirs, _ = ops.generate_rays(None, pose.to(DEV), synth.CAMERA_ANGLE_X, uv=None, rays_per_cam=n, want_targets=False)
    o_dev = pose.to(DEV)[:, :3, 3]
    packet, depth = ops.render_rays(grid, o_dev, dirs, S, delta, gmin, pd, mode=mode, clamp=False, rays_per_origin=n,
                                    return_depth=True, coherent=True)
    _, count = ops.render_rays(grid, o_dev, dirs, S, delta, gmin, pd, mode=mode, clamp=False, rays_per_origin=n, return_count=True)
    sel = np.arange(0, n, 89)
    d_sel = dirs.cpu().numpy()[sel]
    o_sel = np.repeat(pose[:, :3, 3].numpy(), len(sel), axis=0)
    rgba_o, depth_o, count_o, _ = co.render_forward(grid_h.numpy(), o_sel, d_sel, S, delta, np.float32(gmin), pd, mode=mode,
                                                    clamp=False, want_lin=False)
    assert np.array_equal(count.cpu().numpy()[sel], count_o)
    assert rel_err(packet.cpu().numpy()[sel], rgba_o) <= TOL
    assert rel_err(depth.cpu().numpy()[sel], depth_o) <= TOL
    img = ops.render_image_u8(grid, pose.to(DEV), synth.CAMERA_ANGLE_X, side, S, delta, gmin, pd, mode=mode).cpu().numpy()
    want = np.transpose((packet.cpu().numpy() * 255).round().clip(0, 255).astype(np.uint8).reshape(side, side, 4), (1, 0, 2))
    assert img.shape == (side, side, 4) and img.dtype == np.uint8
    assert np.array_equal(img, want), "the kernel's epilogue is src/visualization.py:150-154 bit for bit"


@pytest.mark.parametrize("side,views", [(20, 1), (37, 2), (8, 3), (64, 1)])
def test_packet_tiles_render_the_same_pixels_as_lattice_rows(plx_lib, side, views):
    """The ray-packet kernel marches the even-spread lattice in 16 x 8 tiles per block (plx_render.cu); which thread marches a
    ray must not change the ray: pixels and depth bit-equal to the lattice-row order (plx_tune packet_tile = 0) and to the
    warp-per-ray kernel within 1e-5, for sides that are not multiples of the tile and for several views in one launch."""
    G, S = 24, 96
    pd = synth.GRID_EXTENT / G
    grid = synth.ball_grid(G, occupancy_radius=0.4).to(DEV)
    poses = synth.lookat_poses(views + 1)[1:].to(DEV)
    gmin = ops.grid_origin(grid.shape, pd)
    n = side * side
    dirs, _ = ops.generate_rays(None, poses, synth.CAMERA_ANGLE_X, uv=None, rays_per_cam=n, want_targets=False)
    o = poses[:, :3, 3]
    lib = L.load()
    for mode in ("nearest", "trilinear"):
        out = {}
        try:
            for tile in (1, 0):
                L.check(lib.plx_tune(b"packet_tile", tile))
                out[tile] = ops.render_rays(grid, o, dirs, S, 6.0 / S, gmin, pd, mode=mode, rays_per_origin=n, return_depth=True, coherent=True)
        finally:
            L.check(lib.plx_tune(b"packet_tile", 1))
        assert torch.equal(out[1][0], out[0][0]) and torch.equal(out[1][1], out[0][1])
        warp, depth_w = ops.render_rays(grid, o, dirs, S, 6.0 / S, gmin, pd, mode=mode, rays_per_origin=n, return_depth=True)
        assert float(out[1][0].abs().sum()) > 0
        assert rel_err(out[1][0].cpu().numpy(), warp.cpu().numpy()) <= TOL
        assert rel_err(out[1][1].cpu().numpy(), depth_w.cpu().numpy()) <= TOL


# ------------------------------------------------------------------------------------------------ inference epilogue + splat
def test_image_epilogue_is_the_references_postprocessing(plx_lib):
    """(pix * 255).round().clip(0, 255).astype(uint8), reshape, transpose — src/visualization.py:150-154 — done by the march
    kernel: bit-equal to doing it on the host from the same kernel's float pixels, incl. values outside [0, 1] (unclamped
    grid) and exact .5 ties."""
    G, side, S = 24, 20, 64
    pd = synth.GRID_EXTENT / G
    grid = synth.dense_grid(G, seed=5).to(DEV) * 1.5
    grid[0, 0, 0] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255, 1.0])
    pose = synth.lookat_poses(3)[1:2].to(DEV)
    gmin = ops.grid_origin(grid.shape, pd)
    for mode in ("nearest", "trilinear"):
        img = ops.render_image_u8(grid, pose, synth.CAMERA_ANGLE_X, side, S, 6.0 / S, gmin, pd, mode=mode).cpu().numpy()
        dirs, _ = ops.generate_rays(None, pose, synth.CAMERA_ANGLE_X, uv=None, rays_per_cam=side * side, want_targets=False)
        pix = ops.render_rays(grid, pose[:, :3, 3], dirs, S, 6.0 / S, gmin, pd, mode=mode, clamp=False, rays_per_origin=side * side,
                              coherent=True).cpu().numpy()
        want = np.transpose((pix * 255).round().clip(0, 255).astype(np.uint8).reshape(side, side, 4), (1, 0, 2))
        assert np.array_equal(img, want)
        assert img.max() == 255 and img.min() == 0


def test_gpu_splat_matches_the_references_painter_order_image(plx_lib):
    """visulize_3d_in_2d_fast (src/visualization.py:157-232) against the image the UNMODIFIED reference produced
    (tests/golden/splat_g40.npz): same shape and dtype; the same pixels painted; colours equal wherever the pixel assignment
    agrees.  The reference's projection runs through BLAS matmul and an unstable argsort, so a point that lands within an ulp of
    a pixel boundary, or ties in distance, may legitimately resolve differently: at most 0.5 % of the pixels may differ."""
    import src.visualization as vz
    z = np.load(os.path.join(GOLD, "splat_g40.npz"))
    img = vz.visulize_3d_in_2d_fast(torch.from_numpy(z["grid"]).to(DEV), float(z["pd"]), torch.from_numpy(z["pose"]).to(DEV),
                                    float(z["fov"]), int(z["size_y"]))
    ref = z["image"]
    assert img.shape == ref.shape and img.dtype == ref.dtype == np.float64
    same = (img == ref).all(-1)
    assert same.mean() >= 0.995, f"{(~same).sum()} of {same.size} pixels differ"
    painted, painted_ref = (img != 1.0).any(-1), (ref != 1.0).any(-1)
    assert abs(int(painted.sum()) - int(painted_ref.sum())) <= 0.005 * painted_ref.sum()
    with pytest.raises(L.PlxError, match="no CPU fallback"):
        vz.visulize_3d_in_2d_fast(torch.from_numpy(z["grid"]), float(z["pd"]), torch.from_numpy(z["pose"]), float(z["fov"]), 16)


def test_visulize_3d_in_2d_keeps_the_reference_default_and_refuses_the_cpu(plx_lib):
    import inspect
    import src.visualization as vz
    assert inspect.signature(vz.visulize_3d_in_2d).parameters["device"].default == "cpu"       # src/visualization.py:112
    ck = {"grid": torch.zeros(4, 4, 4, 4), "param": {"points_distance": 0.5, "delta_step": 0.1}}
    with pytest.raises(L.PlxError, match="no CPU fallback"):
        vz.visulize_3d_in_2d(ck, torch.eye(4)[None], 0.6, None, torch.zeros(64, 3), False, 0.2, 16, 8)


# ------------------------------------------------------------------------------------------------ uint8 target images
def test_uint8_images_give_the_same_targets_and_pixels_as_the_float_conversion(plx_lib):
    """src/data_processing.py:58 converts the PNG bytes once, fp32(u8) / 255.  With uint8 images resident the kernels do that
    conversion when they fetch a target pixel: targets bit-equal, the fused march's pixels bit-equal and loss / gradient equal
    up to the order of the atomic sums."""
    C_, H, R, S, G = 3, 16, 64, 64, 24
    pd = synth.GRID_EXTENT / G
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (C_, H, H, 4), dtype=torch.uint8, generator=g)
    f32 = torch.tensor(u8.numpy(), dtype=torch.float) / 255                      # the reference's expression
    poses, uv = synth.lookat_poses(C_).to(DEV), synth.random_uv(C_, R).to(DEV)
    grid = synth.ball_grid(G).to(DEV)
    gmin = ops.grid_origin(grid.shape, pd)
    _, t_u8 = ops.generate_rays(u8.to(DEV), poses, synth.CAMERA_ANGLE_X, uv=uv)
    _, t_f32 = ops.generate_rays(f32.to(DEV), poses, synth.CAMERA_ANGLE_X, uv=uv)
    assert torch.equal(t_u8, t_f32)
    out = []
    for imgs in (u8, f32):
        gg = torch.zeros_like(grid)
        rgba, loss = ops.render_train(grid, gg, S, 6.0 / S, gmin, pd, imgs=imgs.to(DEV), poses=poses, fov=synth.CAMERA_ANGLE_X, uv=uv)
        out.append((rgba, float(loss), gg))
    assert torch.equal(out[0][0], out[1][0])
    assert abs(out[0][1] - out[1][1]) <= 1e-6 * out[1][1]
    assert rel_err(out[0][2].cpu().numpy(), out[1][2].cpu().numpy()) <= 1e-6
    # and through the trainer (resident uint8 image set)
    ta = VoxelTrainer(grid, pd, poses, synth.CAMERA_ANGLE_X, u8.to(DEV), R, S, 6.0 / S, lr=0.0075)
    tb = VoxelTrainer(grid, pd, poses, synth.CAMERA_ANGLE_X, f32.to(DEV), R, S, 6.0 / S, lr=0.0075)
    la, lb = float(ta.step(uv)), float(tb.step(uv))
    assert abs(la - lb) <= 1e-6 * lb and ta.imgs.dtype == torch.uint8


# ------------------------------------------------------------------------------------------------ trainer entry points agree
@pytest.mark.parametrize("tv,beta,mode", [(0.0, 0.0, "nearest"), (1e-4, 0.0, "nearest"), (1e-4, 5e-3, "nearest"), (1e-4, 0.0, "trilinear")])
def test_step_and_step_host_optimise_the_same_objective(plx_lib, tv, beta, mode):
    """step() (device uv), step_host() (pinned uv, loss published to the host) and the unfused three-kernel path must apply the
    same losses — MSE + beta term + TV term — whatever the entry point (the TV weight was dropped on step_host)."""
    cs = Case(24, 3, 8, 48, 96, 6.0 / 96, "ball")
    d = cs.cuda()
    mk = lambda: VoxelTrainer(d["grid"], cs.pd, d["poses"], cs.fov, d["imgs"], cs.R, cs.S, cs.delta, lr=0.0075, tv=tv, beta=beta, mode=mode)
    ta, tb, tc = mk(), mk(), mk()
    tc.unfused = True
    for i in range(3):
        uv = synth.random_uv(cs.C, cs.R, seed=20 + i)
        la = float(ta.step(uv.to(DEV)))
        tb.step_host(uv.pin_memory())
        lb = tb.wait_result()
        lc = float(tc.step(uv.to(DEV)))
        assert abs(la - lb) <= 2e-6 * la and abs(la - lc) <= 2e-6 * la
    torch.cuda.synchronize()
    if tv > 0:
        assert float(ta.tv_loss) > 0 and abs(float(ta.tv_loss) - float(tb.tv_loss)) <= 1e-5 * float(ta.tv_loss)
    for other in (tb, tc):
        assert float(torch.quantile((ta.grid - other.grid).abs().flatten(), 0.999)) <= 1e-5
        assert rel_err(other.grad_abs_sum.cpu().numpy(), ta.grad_abs_sum.cpu().numpy()) <= 1e-5


def test_wait_result_raises_instead_of_spinning_forever(plx_lib):
    cs = Case(16, 2, 8, 16, 32, 6.0 / 32, "ball")
    d = cs.cuda()
    tr = VoxelTrainer(d["grid"], cs.pd, d["poses"], cs.fov, d["imgs"], cs.R, cs.S, cs.delta, lr=0.0075)
    tr.step_host(cs.uv.pin_memory())
    assert tr.wait_result() > 0
    tr.wait_timeout_s = 0.2
    with pytest.raises(L.PlxError, match="never published|not published"):
        tr.wait_result(step=tr.step_count + 5)              # a step nobody issued: the stream is idle, nothing will arrive


# ------------------------------------------------------------------------------------------------ the reference's whole fit()
def _cpu_stream_rand(monkeypatch, seed):
    """torch.rand draws from a CPU generator seeded like the reference run (torch.manual_seed(seed) on device='cpu') and moves
    the draw to the requested device, so a CUDA fit sees the uv stream the CPU reference saw."""
    gen = torch.Generator().manual_seed(seed)
    real = torch.rand

    def rand(*a, **k):
        dev = k.pop("device", "cpu")
        k.pop("generator", None)
        return real(*a, generator=gen, **k).to(dev)
    monkeypatch.setattr(torch, "rand", rand)


def test_fused_fit_reproduces_what_the_references_fit_saved(plx_lib, tmp_path, monkeypatch):
    """tests/golden/fit_g128.npz: the UNMODIFIED scripts/train.py::fit (:68-210) run on the CPU in the build container — 8
    synthetic PNGs, 128^3 grid, 240 steps = all 46 progressive-growing windows plus 10 full-resolution steps, train.py's
    default tv / beta.  plenoxels_b200.fit.fit with the same arguments, dataset (rebuilt from seeds) and uv stream must save
    the same `.pth`: grid and grid_grad within 5e-4 of their own scale (observed 4e-5 / 9e-5) after 240 optimiser steps (sum-order noise passes through
    Adam's normalisation 240 times), whole-array sums within 1e-5, identical `param` block."""
    from plenoxels_b200.fit import fit
    from tests.helpers import FIT_ARGS, write_fit_dataset
    z = np.load(os.path.join(GOLD, "fit_g128.npz"))
    path, tpath = write_fit_dataset(str(tmp_path))
    save = str(tmp_path / "grid.pth")
    _cpu_stream_rand(monkeypatch, int(z["seed"]))
    fit(path=path, transform_path=tpath, save_path=save, device="cuda:0", progressive_growing=True, log_every=0, **FIT_ARGS)
    monkeypatch.undo()
    ck = torch.load(save)
    g, gg = ck["grid"].numpy(), ck["grid_grad"].numpy()
    assert g.shape == (128, 128, 128, 4) and set(ck) == {"grid", "grid_grad", "param"}
    want = eval(str(z["param"]))                                   # the reference's own param dict (repr of python scalars)
    assert {k: v for k, v in ck["param"].items() if k != "device"} == {k: v for k, v in want.items() if k != "device"}
    e_grid = np.abs(g[::4, ::4, ::4] - z["grid_subset"]).max() / float(z["grid_max"])
    e_grad = np.abs(gg[::4, ::4, ::4] - z["grad_subset"]).max() / float(z["grad_max"])
    e_sum = np.abs(g.astype(np.float64).reshape(-1, 4).sum(0) - z["grid_sum"]).max() / np.abs(z["grid_abs_sum"]).max()
    e_gsum = np.abs(gg.astype(np.float64).reshape(-1, 4).sum(0) - z["grad_sum"]).max() / np.abs(z["grad_sum"]).max()
    print(f"fit vs reference fit: grid {e_grid:.2e}, grid_grad {e_grad:.2e}, sums {e_sum:.2e} / {e_gsum:.2e}")
    assert e_grid <= 5e-4 and e_grad <= 5e-4 and e_sum <= 1e-5 and e_gsum <= 1e-5, (e_grid, e_grad, e_sum, e_gsum)


REFERENCE = os.environ.get("PLX_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree exists in the build container only")
def test_unmodified_reference_fit_runs_on_the_dropin_functions(plx_lib, tmp_path, monkeypatch):
    """scripts/train.py of the reference, imported from where it lies and executed UNCHANGED: its `from src.... import`
    lines resolve to this repo's src/ shim, so every hot-path call lands in the sm_100a kernels (lazy-fusion bridge: one fused
    march per step) while clip / mse_loss / Adam stay the script's own torch calls.  Its saved `.pth` must equal the golden the
    same script produced on the CPU with the reference's own src/ (fit_g128.npz)."""
    import importlib.util
    import sys
    import types
    from tests.helpers import FIT_ARGS, write_fit_dataset
    if "tqdm" not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except ImportError:
            sys.modules["tqdm"] = types.SimpleNamespace(tqdm=lambda *a, **k: types.SimpleNamespace(update=lambda *a: None, set_description=lambda *a: None))
    spec = importlib.util.spec_from_file_location("ref_train_unmodified", os.path.join(REFERENCE, "scripts", "train.py"))
    ref_train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_train)
    import src.grid_functions as shim
    assert ref_train.generate_grid is shim.generate_grid, "the script must have bound this repo's functions"
    z = np.load(os.path.join(GOLD, "fit_g128.npz"))
    path, tpath = write_fit_dataset(str(tmp_path))
    save = str(tmp_path / "grid_ref.pth")
    _cpu_stream_rand(monkeypatch, int(z["seed"]))
    ref_train.fit(path=path, transform_path=tpath, save_path=save, device="cuda:0", progressive_growing=True, **FIT_ARGS)
    monkeypatch.undo()
    ck = torch.load(save)
    g, gg = ck["grid"].numpy(), ck["grid_grad"].numpy()
    e_grid = np.abs(g[::4, ::4, ::4] - z["grid_subset"]).max() / float(z["grid_max"])
    e_grad = np.abs(gg[::4, ::4, ::4] - z["grad_subset"]).max() / float(z["grad_max"])
    print(f"unmodified fit() on the drop-in functions vs on the reference's own: grid {e_grid:.2e}, grid_grad {e_grad:.2e}")
    assert e_grid <= 5e-4 and e_grad <= 5e-4, (e_grid, e_grad)


# ------------------------------------------------------------------------------------------------ CUDA-graph replay of the step
@pytest.mark.parametrize("host_uv", [False, True])
def test_graph_replay_equals_plain_steps(plx_lib, host_uv):
    """One captured CUDA graph (fused march + optimiser, device-resident step number / loss slot / Adam scalars,
    PlxReplayState) replayed for 7 steps against the same steps issued one by one: Adam's scalars come from the table
    plx_adam_table fills in double like plx_adam_step, so losses agree to sum-order noise and the state to the 1e-5 bar."""
    cs = Case(24, 3, 8, 64, 96, 6.0 / 96, "ball")
    d = cs.cuda()
    mk = lambda: VoxelTrainer(d["grid"], cs.pd, d["poses"], cs.fov, d["imgs"], cs.R, cs.S, cs.delta, lr=0.0075)
    ta, tb = mk(), mk()
    for i in range(2):                                   # the graph is captured mid-run: step numbering must carry on
        u = synth.random_uv(cs.C, cs.R, seed=40 + i).to(DEV)
        ta.step(u), tb.step(u)
    tb.capture_graph(host_uv=host_uv)
    for i in range(2, 9):
        uv = synth.random_uv(cs.C, cs.R, seed=40 + i)
        la = float(ta.step(uv.to(DEV)))
        if host_uv:
            tb.uv_host.copy_(uv)
            tb.step_graph()
            lb = tb.wait_result()
        else:
            tb.uv.copy_(uv.to(DEV))
            lb = float(tb.step_graph())
        assert abs(la - lb) <= 2e-6 * la, (i, la, lb)
    torch.cuda.synchronize()
    assert ta.step_count == tb.step_count == 9 and int(tb._step_dev.item()) == 9
    assert float(torch.quantile((ta.grid - tb.grid).abs().flatten(), 0.999)) <= 1e-5
    assert rel_err(tb.exp_avg_sq.cpu().numpy(), ta.exp_avg_sq.cpu().numpy()) <= 1e-5
    assert rel_err(tb.grad_abs_sum.cpu().numpy(), ta.grad_abs_sum.cpu().numpy()) <= 1e-5
