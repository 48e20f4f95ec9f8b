"""How much of the fused march's time is load imbalance?  Times K12 on the SAME C2 batch in three ray orders: as drawn, sorted by
in-bounds sample count (longest first), and randomly shuffled.  python tools/order_probe.py [c2|c3]"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoxels_b200 import ops, synth
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
sc = synth.make_scene(name, H=64)
grid = sc.grid.to(dev); gg = torch.zeros_like(grid)
gmin = ops.grid_origin(grid.shape, sc.points_distance)
poses, imgs = sc.poses.to(dev), sc.imgs.to(dev)
C_, R, S = poses.shape[0], sc.rays_per_cam, sc.num_samples
def timed(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
for seed in (0, 1):
    uv = synth.random_uv(C_, R, seed=seed).to(dev)
    dirs, targets = ops.generate_rays(imgs, poses, sc.fov, uv=uv)
    o = poses[:, :3, 3].repeat_interleave(R, 0).contiguous()
    _, cnt = ops.render_rays(grid, o, dirs, S, sc.delta_step, gmin, sc.points_distance, rays_per_origin=1, return_count=True)
    orders = {"as drawn": torch.arange(C_ * R, device=dev), "longest first": torch.argsort(cnt, descending=True),
              "shortest first": torch.argsort(cnt), "shuffled": torch.randperm(C_ * R, device=dev)}
    res = {}
    for tag, perm in orders.items():
        oo, dd, tt = o[perm].contiguous(), dirs[perm].contiguous(), targets[perm].contiguous()
        res[tag] = round(timed(lambda: ops.render_train(grid, gg, S, sc.delta_step, gmin, sc.points_distance, origins=oo, dirs=dd, targets=tt,
                                                        rays_per_origin=1)), 2)
    res["in-kernel ray generation (uv)"] = round(timed(lambda: ops.render_train(grid, gg, S, sc.delta_step, gmin, sc.points_distance, imgs=imgs,
                                                                                 poses=poses, fov=sc.fov, uv=uv)), 2)
    print(json.dumps({"workload": name, "seed": seed, "k12_us": res, "count_min_mean_max": [int(cnt.min()), float(cnt.float().mean()), int(cnt.max())]}), flush=True)
