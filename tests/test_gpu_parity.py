"""GPU parity: the sm_100a kernels (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): nearest-neighbour indices and per-ray sample counts bit-exact; RGB(A), depth, loss and
grid gradients within 1e-5 relative (fp32).
"""
import numpy as np
import pytest
import torch

from oracle import plenoxel_oracle as po
from plenoxels_b200 import ops, synth
from plenoxels_b200.trainer import VoxelTrainer
from tests.helpers import Case, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5

CASES = {
    "c1-dense": dict(G=64, C=1, H=64, R=4096, S=64, delta=6.0 / 64, kind="dense"),       # BASELINE config #1 shape
    "ball-64": dict(G=64, C=4, H=32, R=256, S=64, delta=6.0 / 64, kind="ball"),
    "c2-small": dict(G=128, C=6, H=40, R=64, S=600, delta=0.0125, kind="ball"),          # config #2 geometry, fewer rays
    "soft-32": dict(G=32, C=3, H=16, R=128, S=96, delta=6.0 / 96, kind="soft"),
    "ragged": dict(G=24, C=5, H=8, R=37, S=45, delta=6.0 / 45, kind="dense"),            # N, S not multiples of 32/8
}


@pytest.fixture(scope="module", params=list(CASES))
def case(request, plx_lib):
    return Case(**CASES[request.param])


def test_indices_and_counts_bit_exact(case):
    d = case.cuda()
    idx, count = ops.sample_indices(d["grid"], d["origins"], d["dirs"], case.S, case.delta, case.gmin, case.pd,
                                    rays_per_origin=case.R)
    _, _, ocount, olin = case.oracle_forward()
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), olin), "linear voxel indices differ from the oracle"
    assert np.array_equal(count.cpu().numpy(), ocount)


@pytest.mark.parametrize("mode", ["nearest", "trilinear"])
def test_forward_rgba_depth_count(case, mode):
    d = case.cuda()
    rgba, depth, count = ops.render_rays(d["grid"], d["origins"], d["dirs"], case.S, case.delta, case.gmin, case.pd,
                                         mode=mode, rays_per_origin=case.R, return_depth=True, return_count=True)
    orgba, odepth, ocount, _ = case.oracle_forward(mode)
    assert np.array_equal(count.cpu().numpy(), ocount), "clipped march lost / gained in-bounds samples"
    assert rel_err(rgba.cpu().numpy(), orgba) <= TOL
    assert rel_err(depth.cpu().numpy(), odepth) <= TOL
    # the clipped + early-terminated march (no count requested) gives the same pixels
    rgba2 = ops.render_rays(d["grid"], d["origins"], d["dirs"], case.S, case.delta, case.gmin, case.pd, mode=mode,
                            rays_per_origin=case.R)
    assert rel_err(rgba2.cpu().numpy(), orgba) <= TOL


@pytest.mark.parametrize("mode", ["nearest", "trilinear"])
def test_forward_coherent_ray_packets(case, mode):
    """PLX_COHERENT_RAYS (one ray per thread, packets of 32) renders the same pixels and depth as the oracle."""
    d = case.cuda()
    rgba, depth = ops.render_rays(d["grid"], d["origins"], d["dirs"], case.S, case.delta, case.gmin, case.pd, mode=mode,
                                  rays_per_origin=case.R, return_depth=True, coherent=True)
    orgba, odepth, _, _ = case.oracle_forward(mode)
    assert rel_err(rgba.cpu().numpy(), orgba) <= TOL
    assert rel_err(depth.cpu().numpy(), odepth) <= TOL


@pytest.mark.parametrize("mode", ["nearest", "trilinear"])
@pytest.mark.parametrize("beta", [0.0, 5e-3])
def test_backward_grid_gradient(case, mode, beta):
    d = case.cuda()
    grid = d["grid"].clone().requires_grad_(True)
    bom = beta / (case.N * case.S)
    rgba = ops.render_rays(grid, d["origins"], d["dirs"], case.S, case.delta, case.gmin, case.pd, mode=mode,
                           rays_per_origin=case.R, beta_over_m=bom)
    loss = torch.nn.functional.mse_loss(rgba, d["targets"])
    loss.backward()
    orgba, _, _, _ = case.oracle_forward(mode)
    oloss, gpix = po.mse_loss(orgba, case.targets)
    ograd = case.oracle_backward(gpix, mode, beta=beta)
    assert abs(float(loss) - oloss) <= TOL * abs(oloss)
    assert rel_err(grid.grad.cpu().numpy(), ograd) <= TOL


@pytest.mark.parametrize("beta", [0.0, 5e-3])
@pytest.mark.parametrize("source", ["rays", "uv"])
def test_fused_training_march_matches_oracle(case, beta, source):
    """K12 (ray generation + forward + MSE + backward in one kernel) against the oracle's loss and gradient."""
    d = case.cuda()
    gg = torch.zeros_like(d["grid"])
    bom = beta / (case.N * case.S)
    if source == "rays":
        rgba, loss = ops.render_train(d["grid"], gg, case.S, case.delta, case.gmin, case.pd, origins=d["origins"],
                                      dirs=d["dirs"], targets=d["targets"], rays_per_origin=case.R, beta_over_m=bom)
    else:
        rgba, loss = ops.render_train(d["grid"], gg, case.S, case.delta, case.gmin, case.pd, imgs=d["imgs"],
                                      poses=d["poses"], fov=case.fov, uv=d["uv"], beta_over_m=bom)
    orgba, _, _, _ = case.oracle_forward()
    oloss, gpix = po.mse_loss(orgba, case.targets)
    ograd = case.oracle_backward(gpix, beta=beta)
    assert rel_err(rgba.cpu().numpy(), orgba) <= TOL
    assert abs(float(loss) - oloss) <= TOL * abs(oloss)
    assert rel_err(gg.cpu().numpy(), ograd) <= TOL


def test_fused_training_march_opaque_cells():
    """Rays that saturate on alpha == 1 cells: the gradient of the opaque sample needs the colour BEHIND it."""
    case = Case(G=32, C=3, H=16, R=200, S=160, delta=6.0 / 160, kind="ball")
    grid = case.grid.clone()
    occ = grid[..., 3] > 0
    flip = occ & (torch.rand(grid.shape[:3], generator=torch.Generator().manual_seed(9)) < 0.15)
    grid[..., 3][flip] = 1.25                                       # clips to exactly 1
    d = case.cuda()
    gcu = grid.cuda()
    gg = torch.zeros_like(gcu)
    rgba, loss = ops.render_train(gcu, gg, case.S, case.delta, case.gmin, case.pd, origins=d["origins"], dirs=d["dirs"],
                                  targets=d["targets"], rays_per_origin=case.R)
    orgba, _, _, _ = case.oracle_forward(grid=grid.numpy())
    oloss, gpix = po.mse_loss(orgba, case.targets)
    ograd = case.oracle_backward(gpix, grid=grid.numpy())
    assert rel_err(rgba.cpu().numpy(), orgba) <= TOL
    assert rel_err(gg.cpu().numpy(), ograd) <= TOL
    # and the autograd pair K1/K2 agrees on the same scene
    g2 = gcu.clone().requires_grad_(True)
    pix = ops.render_rays(g2, d["origins"], d["dirs"], case.S, case.delta, case.gmin, case.pd, rays_per_origin=case.R)
    torch.nn.functional.mse_loss(pix, d["targets"]).backward()
    assert rel_err(g2.grad.cpu().numpy(), ograd) <= TOL


def test_backward_without_saved_carry_matches(case):
    """K2's in-kernel first pass (tcarry = NULL) must agree with the path that reuses K1's chunk transmittances."""
    import ctypes as C
    from plenoxels_b200 import _lib as L
    d = case.cuda()
    grid = d["grid"].clone().requires_grad_(True)
    rgba = ops.render_rays(grid, d["origins"], d["dirs"], case.S, case.delta, case.gmin, case.pd, rays_per_origin=case.R)
    g = torch.randn_like(rgba)
    rgba.backward(g)
    gg = torch.zeros_like(d["grid"])
    b = L.PlxRenderBwd()
    b.march = L.make_march(d["grid"], case.S, case.delta, case.gmin, case.pd, "nearest", True)
    b.rays = L.make_rays(d["origins"], d["dirs"], case.R)
    b.grid, b.grad_rgba, b.tcarry, b.grad_grid = d["grid"].data_ptr(), g.data_ptr(), None, gg.data_ptr()
    L.check(L.load().plx_render_bwd(C.byref(b), L.stream_ptr(gg.device)))
    torch.cuda.synchronize()
    assert rel_err(gg.cpu().numpy(), grid.grad.cpu().numpy()) <= 2e-6


@pytest.mark.parametrize("mode", ["nearest", "trilinear"])
def test_channel_planar_grid_matches_contiguous(case, mode):
    """Pooled grids reach the lookup channel-planar: the strided path must give the contiguous result
    and route the gradient back through the view."""
    d = case.cuda()
    planar = d["grid"].permute(3, 0, 1, 2).contiguous().requires_grad_(True)       # (4,X,Y,Z) storage
    view = planar.permute(1, 2, 3, 0)                                              # (X,Y,Z,4) with strides (YZ, Z, 1, XYZ)
    assert not view.is_contiguous()
    a = ops.render_rays(view, d["origins"], d["dirs"], case.S, case.delta, case.gmin, case.pd, mode=mode,
                        rays_per_origin=case.R)
    g = torch.randn_like(a)
    a.backward(g)
    ref = d["grid"].clone().requires_grad_(True)
    b = ops.render_rays(ref, d["origins"], d["dirs"], case.S, case.delta, case.gmin, case.pd, mode=mode,
                        rays_per_origin=case.R)
    b.backward(g)
    assert rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy()) <= 2e-6
    assert rel_err(planar.grad.permute(1, 2, 3, 0).cpu().numpy(), ref.grad.cpu().numpy()) <= 2e-6


def test_generate_rays_matches_oracle(case):
    d = case.cuda()
    dirs, targets = ops.generate_rays(d["imgs"], d["poses"], case.fov, uv=d["uv"])
    assert np.array_equal(targets.cpu().numpy(), case.targets), "target pixel lookup differs"
    assert np.abs(dirs.cpu().numpy() - case.dirs).max() <= 1.2e-7        # 1 ulp of a unit vector component
    assert np.array_equal(dirs.cpu().numpy(), case.dirs), "directions are expected to be bit-exact too"


def test_even_spread_lattice():
    poses = synth.lookat_poses(3).cuda()
    imgs = synth.random_images(3, 64, 64).cuda()
    dirs, targets = ops.generate_rays(imgs, poses, 0.6911112070083618, uv=None, rays_per_cam=4096)
    uv = po.even_spread_uv(3, 4096)
    odirs, otargets, _ = po.generate_rays(imgs.cpu().numpy(), poses.cpu().numpy(), 0.6911112070083618, uv)
    assert np.array_equal(targets.cpu().numpy(), otargets)
    assert np.array_equal(dirs.cpu().numpy(), odirs)


def test_eager_sample_normalize_gather(case):
    d = case.cuda()
    pos = ops.sample_points(d["origins"], d["dirs"], case.S, case.delta, rays_per_origin=case.R)
    opos = po.sample_positions(case.origins_per_ray, case.dirs, case.S, case.delta).reshape(-1, 3)
    assert np.array_equal(pos.cpu().numpy(), opos)
    ns = ops.normalize_points(pos, case.gmin, case.pd)
    ons = po.normalize_positions(opos, case.gmin, case.pd)
    assert np.array_equal(ns.cpu().numpy(), ons)
    clipped = d["grid"].clip(0, 1)
    vals, mask = ops.gather_nearest(ns, clipped)
    ovals, oinb = po.gather_nearest(ons, np.clip(case.grid.numpy(), 0, 1))
    assert np.array_equal(mask.cpu().numpy(), oinb)
    assert np.array_equal(vals.cpu().numpy(), ovals), "gathered values (wrapped, unmasked) must be bit-exact"
    tvals, tmask = ops.trilinear_lookup(ns, clipped)
    otv, otm = po.trilinear_lookup(ons, np.clip(case.grid.numpy(), 0, 1))
    assert np.array_equal(tmask.cpu().numpy(), otm)
    assert np.array_equal(tvals.cpu().numpy(), otv), "trilinear values must be bit-exact (same op order)"


def test_eager_composite_fwd_bwd(case):
    torch.manual_seed(3)
    s = torch.rand(case.C, case.R, case.S, 4, device="cuda")
    s[..., 3] *= 0.3
    s[torch.rand(case.C, case.R, case.S, device="cuda") < 0.02] = torch.tensor([0.5, 0.25, 0.75, 1.0], device="cuda")
    s.requires_grad_(True)
    out = ops.composite(s)
    g = torch.randn_like(out)
    out.backward(g)
    sn = s.detach().cpu().numpy()
    assert rel_err(out.detach().cpu().numpy(), po.composite(sn, dtype=np.float64)) <= TOL
    assert rel_err(s.grad.cpu().numpy(), po.composite_backward(sn, g.cpu().numpy())) <= TOL


def test_eager_chain_autograd_matches_fused(case):
    """The unfused path (reference call sequence on eager kernels + torch glue) and the fused K1/K2 agree."""
    d = case.cuda()
    g1 = d["grid"].clone().requires_grad_(True)
    pos = ops.sample_points(d["origins"], d["dirs"], case.S, case.delta, rays_per_origin=case.R)
    ns = ops.normalize_points(pos, case.gmin, case.pd)
    vals, mask = ops.gather_nearest(ns, g1.clip(0, 1))
    vals = vals * mask.unsqueeze(-1)
    pix = ops.composite(vals.reshape(case.C, case.R, case.S, 4)).reshape(-1, 4)
    torch.nn.functional.mse_loss(pix, d["targets"]).backward()
    g2 = d["grid"].clone().requires_grad_(True)
    pix2 = ops.render_rays(g2, d["origins"], d["dirs"], case.S, case.delta, case.gmin, case.pd, rays_per_origin=case.R)
    torch.nn.functional.mse_loss(pix2, d["targets"]).backward()
    assert rel_err(pix.detach().cpu().numpy(), pix2.detach().cpu().numpy()) <= 2e-6
    assert rel_err(g1.grad.cpu().numpy(), g2.grad.cpu().numpy()) <= 2e-6


@pytest.mark.parametrize("pd", [0.0125, 0.025, 0.05, 0.00625, 3.2 / 24, 3.2 / 93, 0.1 * 4, 1.0, 7.3e-4])
def test_selftest_exact_arithmetic(plx_lib, pd):
    """The hoisted-reciprocal quotient, inlined sqrt and per-element quotient are bit-identical to __fdiv_rn/__fsqrt_rn."""
    bad = torch.zeros(3, dtype=torch.int64, device="cuda")
    from plenoxels_b200 import _lib as L
    L.check(plx_lib.plx_selftest_arith(pd, 1 << 26, 12345, bad.data_ptr(), L.stream_ptr(bad.device)))
    assert bad.tolist() == [0, 0, 0], f"mismatches (div, sqrt, var-div) = {bad.tolist()}"


def test_adam_step_matches_oracle():
    torch.manual_seed(0)
    n = 64 * 64 * 16 * 4 + 4
    p = (torch.rand(n) * 1.4 - 0.2)
    m, v, ga = torch.zeros(n), torch.zeros(n), torch.zeros(n)
    pc, mc, vc, gac = (t.cuda() for t in (p, m, v, ga))
    pn, mn, vn, gan = (t.numpy().copy() for t in (p, m, v, ga))
    for step in range(1, 5):
        g = torch.randn(n) * 1e-3 * (torch.rand(n) < 0.6)
        gc = g.cuda()
        ops.adam_step(pc, gc, mc, vc, gac, step, lr=0.0075)
        pn, mn, vn, gan = po.adam_step(pn, g.numpy(), mn, vn, gan, 0.0075, step)
        assert float(gc.abs().max()) == 0.0, "gradient must be cleared for the next step"
    assert np.array_equal(mc.cpu().numpy(), mn)
    assert np.array_equal(vc.cpu().numpy(), vn)
    assert np.array_equal(gac.cpu().numpy(), gan)
    assert np.array_equal(pc.cpu().numpy(), pn), "Adam parameters are bit-exact against the fp32 oracle"


@pytest.mark.parametrize("skip_same", [1, 0])
def test_adam_sparse_steps_are_bit_exact(plx_lib, skip_same):
    """The large-grid form of K3 — per-line skipping of unchanged stores — forced on / off on a small array (plx_tune): sparse
    gradients, cells that were never touched, cells whose moments still move their parameter long after their last gradient,
    a NaN gradient; every step bit-exact against the fp32 oracle, untouched cells keep their exact bits (incl. -0.0)."""
    from plenoxels_b200 import _lib as L
    lib = L.load()
    torch.manual_seed(3)
    n = 48 * 1024 * 4
    p = (torch.rand(n) * 1.4 - 0.2)
    p[:64] = -0.0
    m, v, ga = torch.zeros(n), torch.zeros(n), torch.zeros(n)
    pc, mc, vc, gac = (t.cuda() for t in (p, m, v, ga))
    pn, mn, vn, gan = (t.numpy().copy() for t in (p, m, v, ga))
    try:
        L.check(lib.plx_tune(b"adam_skip_same", skip_same))
        for step in range(1, 9):
            g = torch.zeros(n)
            if step <= 5:                     # a few 128-byte lines per step; steps 6-8 have no gradient at all (moments only)
                lines = torch.randint(2, n // 32, (40,))
                for ln in lines.tolist():
                    g[ln * 32: ln * 32 + 32] = torch.randn(32) * 1e-3 * (torch.rand(32) < 0.7)
            if step == 3:
                g[n - 7] = float("nan")
            gc = g.cuda()
            ops.adam_step(pc, gc, mc, vc, gac, step, lr=0.0075)
            pn, mn, vn, gan = po.adam_step(pn, g.numpy(), mn, vn, gan, 0.0075, step)
            assert float(torch.nan_to_num(gc).abs().max()) == 0.0 and not bool(torch.isnan(gc).any())
            for name, a, b in (("m", mc, mn), ("v", vc, vn), ("|g|", gac, gan), ("p", pc, pn)):
                assert np.array_equal(a.cpu().numpy(), b, equal_nan=True), f"step {step}: {name}"
        if skip_same:       # untouched lines are never rewritten, so even the sign of a zero parameter survives; with every line
            # stored, -0.0 + (update of +0.0) comes back as +0.0 where ATen keeps -0.0: equal values, different sign bit of zero
            assert np.array_equal(np.signbit(pc[:64].cpu().numpy()), np.signbit(pn[:64]))
    finally:
        L.check(lib.plx_tune(b"adam_skip_same", -1))


@pytest.mark.parametrize("mode", ["nearest", "trilinear"])
def test_train_steps_match_oracle(mode):
    """Three whole steps (ray generation -> forward -> MSE -> backward -> Adam) through plx_train_step."""
    case = Case(G=32, C=4, H=16, R=96, S=80, delta=6.0 / 80, kind="ball")
    d = case.cuda()
    tr = VoxelTrainer(d["grid"], case.pd, d["poses"], case.fov, d["imgs"], case.R, case.S, case.delta, lr=0.0075, mode=mode)
    grid = case.grid.numpy().copy()
    m, v, ga = np.zeros_like(grid), np.zeros_like(grid), np.zeros_like(grid)
    for step in range(1, 4):
        uv = synth.random_uv(case.C, case.R, seed=10 + step)
        dirs, targets, _ = po.generate_rays(case.imgs.numpy(), case.poses.numpy(), case.fov, uv.numpy())
        loss_o, grad_o, grid, m, v, ga = po.train_step(grid, m, v, ga, case.origins_per_ray, dirs, targets, case.S, case.delta,
                                                       case.gmin, case.pd, 0.0075, step, mode=mode)
        if step % 2:
            loss = float(tr.step(uv.cuda()))
        else:
            loss_host = tr.step_host(uv.pin_memory())
            loss = tr.wait_result()                       # spins on the pinned {loss, step} record, no driver sync
            torch.cuda.synchronize()
            assert loss == float(loss_host) and int(tr._result_i32[1]) == step
        assert abs(loss - loss_o) <= TOL * abs(loss_o)
        assert rel_err(tr.grad_abs_sum.cpu().numpy(), ga) <= TOL
        # Adam's first steps are +-lr*sign(g): cells whose gradient is O(rounding noise) may flip; compare the bulk
        diff = np.abs(tr.grid.cpu().numpy() - grid)
        assert np.quantile(diff, 0.999) <= 1e-5, f"step {step}: {np.quantile(diff, 0.999)}"


@pytest.mark.parametrize("G,k,s", [(24, 3, 1), (24, 5, 1), (16, 2, 1), (20, 4, 1), (70, 3, 1), (12, 6, 1), (40, 9, 2), (40, 21, 5), (64, 33, 8), (31, 7, 3)])
def test_average_pool3d_grid_matches_library(plx_lib, G, k, s):
    """Box-filter pooling (plx_avgpool3d_fwd/bwd: three separable passes, or the one-pass sliding-window kernel for stride-1
    windows up to 5) vs F.avg_pool3d as average_pool3d_grid calls it (src/grid_functions.py:173-181), forward and backward;
    tolerance = summation order only."""
    torch.manual_seed(G + k)
    grid = (torch.rand(G, G + 1, G + 2, 4, device="cuda") * 1.4 - 0.2).requires_grad_(True)
    out = ops.avgpool3d_grid(grid, k, s)
    ref_in = grid.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.avg_pool3d(ref_in.permute(3, 0, 1, 2).unsqueeze(0), (k, k, k), stride=s).squeeze(0).permute(1, 2, 3, 0)
    assert out.shape == ref.shape
    # judge both against an fp64 evaluation: the library's own fp32 summation noise grows with the window (k^3 terms)
    in64 = grid.detach().double().requires_grad_(True)
    ref64 = torch.nn.functional.avg_pool3d(in64.permute(3, 0, 1, 2).unsqueeze(0), (k, k, k), stride=s).squeeze(0).permute(1, 2, 3, 0)
    err_ours = float((out.double() - ref64).abs().max())
    err_lib = float((ref.double() - ref64).abs().max())
    assert err_ours <= 2 * err_lib + 1e-6, (err_ours, err_lib)
    g = torch.randn_like(out)
    out.backward(g)
    ref.backward(g)
    ref64.backward(g.double())
    scale = float(in64.grad.abs().max())
    gerr_ours = float((grid.grad.double() - in64.grad).abs().max()) / scale
    gerr_lib = float((ref_in.grad.double() - in64.grad).abs().max()) / scale
    assert gerr_ours <= 2 * gerr_lib + 1e-6, (gerr_ours, gerr_lib)


@pytest.mark.parametrize("shape", [(12, 11, 9), (32, 32, 32), (5, 40, 7)])
def test_tv_loss_kernel_matches_oracle(plx_lib, shape):
    """plx_tv_loss (value + gradient accumulated into an existing buffer) vs the oracle (scripts/train.py:44-65)."""
    torch.manual_seed(sum(shape))
    grid = (torch.rand(*shape, 4, device="cuda") * 1.4 - 0.2)
    base = torch.randn(*shape, 4, device="cuda") * 1e-3
    grad = base.clone()
    oloss, ograd = po.tv_loss(grid.cpu().numpy())
    ops.tv_loss_(grid, 0.5, grad)                                           # accumulates into an existing gradient
    assert rel_err(grad.cpu().numpy(), base.cpu().numpy().astype(np.float64) + 0.5 * ograd) <= TOL
    tv = 1e-5
    out = ops.tv_loss_(grid, tv, None)                                      # value only
    assert abs(float(out) - tv * oloss) <= 1e-5 * tv * oloss
    g0 = torch.zeros_like(grid)
    ops.tv_loss_(grid, tv, g0)
    assert rel_err(g0.cpu().numpy(), tv * ograd) <= TOL
    const = torch.full((4, 4, 4, 4), 0.5, device="cuda")
    gz = torch.zeros_like(const)
    assert float(ops.tv_loss_(const, tv, gz)) == 0.0 and float(gz.abs().max()) == 0.0


def test_trainer_with_tv_and_beta_matches_oracle(plx_lib):
    """train.py's default regularisers (tv=1e-5, beta=5e-3) through VoxelTrainer: gradient = MSE + beta + TV terms."""
    case = Case(G=24, C=3, H=16, R=64, S=64, delta=6.0 / 64, kind="soft")
    d = case.cuda()
    tv, beta = 1e-5, 5e-3
    tr = VoxelTrainer(d["grid"], case.pd, d["poses"], case.fov, d["imgs"], case.R, case.S, case.delta, lr=0.0075, beta=beta, tv=tv)
    loss = float(tr.step(d["uv"]))
    orgba, _, _, _ = case.oracle_forward()
    oloss, gpix = po.mse_loss(orgba, case.targets)
    ograd = case.oracle_backward(gpix, beta=beta)
    tvl, tvg = po.tv_loss(case.grid.numpy())
    assert abs(loss - oloss) <= TOL * oloss
    assert abs(float(tr.tv_loss) - tv * tvl) <= 1e-5 * tv * tvl
    assert rel_err(tr.grad_abs_sum.cpu().numpy(), np.abs(ograd + tv * tvg)) <= TOL


def test_visulize_3d_in_2d_matches_reference_image(plx_lib):
    """The drop-in ray-marched inference (src/visualization.py:111-154, coherent ray-packet kernel) against the uint8 image
    the UNMODIFIED reference rendered for the same checkpoint and camera (tests/golden/inference_g24.npz)."""
    import os
    import src.grid_functions as gf
    import src.visualization as vz
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "inference_g24.npz"))
    G = z["grid"].shape[0]
    ck = {"grid": torch.from_numpy(z["grid"]), "param": {"points_distance": float(z["pd"]), "delta_step": float(z["delta"])}}
    coords, _, _, _ = gf.generate_grid(G, G, G, points_distance=float(z["pd"]), info_size=4, device="cuda:0")
    res = int(z["res"])
    img = vz.visulize_3d_in_2d(ck, torch.from_numpy(z["poses"]).cuda(), float(z["fov"]), torch.from_numpy(z["imgs"]).cuda(), coords,
                               True, float(z["threshold"]), res * res, int(z["S"]), device="cuda:0")
    assert img.shape == z["image"].shape and img.dtype == np.uint8
    assert np.abs(img.astype(int) - z["image"].astype(int)).max() <= 1
    assert (img == z["image"]).mean() > 0.99
