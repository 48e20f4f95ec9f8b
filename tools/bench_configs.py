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
    for S in (64, 128, 256, 512):
        delta = 6.0 / S
        for logn in (10, 12, 14, 16, 18, 20):
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


if __name__ == "__main__":
    which = sys.argv[1:] or ["c4"]
    if "c4" in which:
        c4("nearest")
        if "trilinear" in which:
            c4("trilinear")
    if "c5" in which:
        c5()
