"""Secondary measurements for the other BASELINE.json configs (not the bench.py headline):
   c4  512^3 grid, inference-only render of an 800x800 view (even-spread rays, S=600, delta=0.01)  -> Mrays/s, GB/s
   c5  ray-batch sweep N in {2^10..2^20} x S in {64..512} on a 256^3 grid, forward + fused train march -> GB/s vs roofline
   python tools/bench_configs.py [c4] [c5] [trilinear]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoxels_b200 import ops, synth  # noqa: E402

PEAK = 6466.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda:0")


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def c4(mode="nearest"):
    G, S, delta, side = 512, 600, 0.01, 800
    pd = synth.GRID_EXTENT / G
    grid = synth.ball_grid(G).to(dev).clip(0, 1)
    grid[..., 3][grid[..., 3] < 0.2] = 0.0                       # scripts/compare_inference_to_image.py:91 threshold
    poses = synth.lookat_poses(4)[1:2].to(dev)
    gmin = ops.grid_origin(grid.shape, pd)
    dirs, _ = ops.generate_rays(None, poses, synth.CAMERA_ANGLE_X, uv=None, rays_per_cam=side * side, want_targets=False)
    o = poses[:, :3, 3]
    n = side * side
    _, cnt = ops.render_rays(grid, o, dirs, S, delta, gmin, pd, mode=mode, clamp=False, rays_per_origin=n, return_count=True)
    m_in = int(cnt.sum())
    per = 16 * (8 if mode == "trilinear" else 1)
    for coherent in (False, True):
        ms = timed(lambda: ops.render_rays(grid, o, dirs, S, delta, gmin, pd, mode=mode, clamp=False, rays_per_origin=n,
                                           coherent=coherent), n=10)
        gbs = (per * m_in + 40 * n) / (ms * 1e-3) / 1e9
        print(json.dumps({"config": "c4", "mode": mode, "kernel": "ray packets (thread/ray)" if coherent else "warp/ray",
                          "rays": n, "S": S, "m_in": m_in, "ms_per_frame": ms, "Mrays_per_s": n / ms / 1e3,
                          "algorithmic_GBs": gbs, "frac_of_measured_peak": gbs / PEAK}))


def c5():
    G = 256
    pd = synth.GRID_EXTENT / G
    grid = synth.ball_grid(G).to(dev)
    gg = torch.zeros_like(grid)
    gmin = ops.grid_origin(grid.shape, pd)
    poses = synth.lookat_poses(64).to(dev)
    quick = os.environ.get("PLX_C5_QUICK") == "1"
    for S in ((256, 512) if quick else (64, 128, 256, 512)):
        delta = 6.0 / S
        for logn in ((16, 20) if quick else (10, 12, 14, 16, 18, 20)):
            n = 1 << logn
            R = n // 64
            uv = torch.rand(64, R, 2, device=dev, generator=torch.Generator(device=dev).manual_seed(logn))
            imgs = torch.rand(64, 32, 32, 4, device=dev)
            dirs, targets = ops.generate_rays(imgs, poses, synth.CAMERA_ANGLE_X, uv=uv)
            o = poses[:, :3, 3]
            _, cnt = ops.render_rays(grid, o, dirs, S, delta, gmin, pd, rays_per_origin=R, return_count=True)
            m_in = int(cnt.sum())
            ms_f = timed(lambda: ops.render_rays(grid, o, dirs, S, delta, gmin, pd, rays_per_origin=R), n=10)
            ms_t = timed(lambda: ops.render_train(grid, gg, S, delta, gmin, pd, origins=o, dirs=dirs, targets=targets,
                                                  rays_per_origin=R), n=10)
            gf = (16 * m_in + 40 * n) / (ms_f * 1e-3) / 1e9
            gt = (64 * m_in + 96 * n) / (ms_t * 1e-3) / 1e9
            print(json.dumps({"config": "c5", "S": S, "rays": n, "m_in": m_in, "fwd_ms": round(ms_f, 4), "fwd_Mrays_s": round(n / ms_f / 1e3, 1),
                              "fwd_GBs": round(gf, 1), "fwd_frac": round(gf / PEAK, 3), "train_ms": round(ms_t, 4),
                              "train_Mrays_s": round(n / ms_t / 1e3, 1), "train_GBs": round(gt, 1), "train_frac": round(gt / PEAK, 3)}))


def pool():
    """average_pool3d_grid fwd+bwd (scripts/train.py:110-118) on a 256^3 grid: separable kernels vs the library call."""
    G = 256
    grid = torch.rand(G, G, G, 4, device=dev).requires_grad_(True)
    for k in (93, 45, 9, 3):
        s = max(1, k // 4)

        def ours():
            out = ops.avgpool3d_grid(grid, k, s)
            out.backward(torch.ones_like(out))
            grid.grad = None

        def lib():
            out = torch.nn.functional.avg_pool3d(grid.permute(3, 0, 1, 2).unsqueeze(0), (k, k, k), stride=s)
            out.backward(torch.ones_like(out))
            grid.grad = None

        gd = grid.detach()
        go = torch.ones_like(ops.avgpool3d_grid(gd, k, s))
        gin = torch.empty_like(gd)

        def ours_fit():          # the two calls plenoxels_b200.fit makes per pooled step (no autograd graph, gradient written in place)
            ops.avgpool3d_grid(gd, k, s)
            ops.avgpool3d_grid_backward_into(go, (G, G, G), k, s, gin)

        ms_o = timed(ours, n=5, warm=2)
        ms_f = timed(ours_fit, n=5, warm=2)
        ms_l = timed(lib, n=2, warm=1)
        print(json.dumps({"config": "pool 256^3", "kernel": k, "stride": s, "ours_fwd_bwd_ms": round(ms_o, 3),
                          "ours_fwd_bwd_as_fit_calls_it_ms": round(ms_f, 3),
                          "library_fwd_bwd_ms": round(ms_l, 3), "speedup": round(ms_l / ms_o, 1)}))


def fit_body():
    """The reference's loop body (scripts/train.py:130-184) written with the drop-in functions + torch.optim.Adam, as the
    unmodified fit() runs it: lazy-fusion bridge on / off, at the C2 shape."""
    import src.grid_functions as gf
    import src.ray_sampling as rs
    sc = synth.make_scene("c2", H=200)
    C_, R, S = sc.poses.shape[0], sc.rays_per_cam, sc.num_samples
    poses, imgs = sc.poses.to(dev), sc.imgs.to(dev)
    for lazy_on in ("1", "0"):
        os.environ["PLX_LAZY"] = lazy_on
        gi, cells, _, _ = gf.generate_grid(sc.G, sc.G, sc.G, points_distance=sc.points_distance, info_size=4, device=dev)
        with torch.no_grad():
            cells.copy_(sc.grid.to(dev))
        opt = torch.optim.Adam([cells], lr=sc.lr)
        gabs = torch.zeros_like(cells)

        def step():
            samples, targets, cam, dirs = rs.sample_camera_rays_batched(
                transform_matrices=poses, camera_angle_x=sc.fov, imgs=imgs, number_of_rays=R, num_samples=S,
                delta_step=sc.delta_step, even_spread=False, camera_ray=False, device=dev)
            ns = rs.normalize_samples_for_indecies(gi, samples, sc.points_distance)
            nearest, mask = gf.get_nearest_voxels(ns, cells.clip(0, 1))
            nearest = (nearest * mask.unsqueeze(-1)).reshape(C_, R, S, 4)
            pix = rs.compute_alpha_weighted_pixels(nearest).reshape(-1, 4)
            loss = torch.nn.functional.mse_loss(pix, targets)
            opt.zero_grad()
            loss.backward()
            opt.step()
            gabs.add_(cells.grad.abs())

        ms = timed(step, n=20, warm=3)
        print(json.dumps({"config": "fit() loop body via drop-in API, C2 shape", "lazy_bridge": lazy_on == "1",
                          "ms_per_step": round(ms, 3), "Mrays_per_s": round(C_ * R / ms / 1e3, 2)}))


def c2_regularised():
    """C2 with train.py's default regularisers (tv=1e-5, beta=5e-3; scripts/train.py:227-228) — the secondary C2 case of
    SURVEY.md §8d: the beta term touches every in-bounds sample (no early termination), TV adds two dense passes."""
    from plenoxels_b200.trainer import VoxelTrainer
    sc = synth.make_scene("c2", H=200)
    uvs = [synth.random_uv(sc.poses.shape[0], sc.rays_per_cam, seed=i).to(dev) for i in range(8)]
    for tv, beta in ((0.0, 0.0), (0.0, 5e-3), (1e-5, 5e-3)):
        tr = VoxelTrainer(sc.grid.to(dev), sc.points_distance, sc.poses.to(dev), sc.fov, sc.imgs.to(dev), sc.rays_per_cam,
                          sc.num_samples, sc.delta_step, lr=sc.lr, beta=beta, tv=tv)
        i = [0]

        def step():
            tr.step(uvs[i[0] % 8])
            i[0] += 1

        ms = timed(step, n=50, warm=5)
        print(json.dumps({"config": "c2 regularised", "tv": tv, "beta": beta, "ms_per_step": round(ms, 4),
                          "Mrays_per_s": round(sc.n_rays / ms / 1e3, 1)}))


def main_fit():
    """The whole training run of scripts/main.py:24-44 (256^3 grid, pd 0.0125, 100 views x 75 rays x 600 samples, lr 0.0025,
    1650 steps, progressive growing on, tv = beta = 0) through plenoxels_b200.fit.GridFitter on synthetic 800x800 views."""
    import time
    from plenoxels_b200.fit import GridFitter
    poses, imgs = synth.lookat_poses(100), synth.random_images(100, 800, 800)
    ft = GridFitter([256, 256, 256], 0.0125, poses, synth.CAMERA_ANGLE_X, imgs, 75, 600, 0.0125, 0.0025, tv=0, beta=0, device="cuda:0")
    for i in range(3):
        ft.step(i)
    torch.cuda.synchronize()
    ft2 = GridFitter([256, 256, 256], 0.0125, poses, synth.CAMERA_ANGLE_X, imgs, 75, 600, 0.0125, 0.0025, tv=0, beta=0, device="cuda:0")
    t0 = time.perf_counter()
    marks = {}
    for i in range(1650):
        ft2.step(i)
        if i == 229:
            torch.cuda.synchronize()
            marks["progressive_230_steps_s"] = time.perf_counter() - t0
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print(json.dumps({"config": "scripts/main.py training run (1650 steps, 256^3, 7500 rays x 600 samples/step)",
                      "total_s": round(total, 3), **{k: round(v, 3) for k, v in marks.items()},
                      "full_res_ms_per_step": round((total - marks["progressive_230_steps_s"]) / 1420 * 1e3, 4),
                      "final_mse": float(ft2.last["mse"]), "reference_README": "2-10 minutes depending on hardware"}))


def ref_cuda():
    """The reference's own op sequence (oracle/torch_port.py = scripts/train.py:130-184 restated op for op) on device="cuda":
    stock PyTorch eager on the same B200, the 'existing GPU path' of BASELINE.md §4."""
    from oracle.torch_port import ReferenceStep
    sc = synth.make_scene("c2", H=200)
    ref = ReferenceStep(sc.grid, sc.points_distance, sc.poses, sc.fov, sc.imgs, sc.rays_per_cam, sc.num_samples, sc.delta_step,
                        sc.lr, device="cuda:0")
    uv = synth.random_uv(sc.poses.shape[0], sc.rays_per_cam).to(dev)
    ms = timed(lambda: ref.step(uv), n=10, warm=2)
    print(json.dumps({"config": "reference op sequence, stock PyTorch eager on cuda, C2 shape", "ms_per_step": round(ms, 3),
                      "Mrays_per_s": round(sc.n_rays / ms / 1e3, 2)}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c4"]
    if "mainfit" in which:
        main_fit()
    if "c2reg" in which:
        c2_regularised()
    if "refcuda" in which:
        ref_cuda()
    if "pool" in which:
        pool()
    if "fit" in which:
        fit_body()
    if "c4" in which:
        c4("nearest")
        if "trilinear" in which:
            c4("trilinear")
    if "c5" in which:
        c5()
