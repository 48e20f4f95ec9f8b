"""Why does the fused march take 50 us inside the training loop and 42 us on its own?  Times K12 on the initial and on a TRAINED C2
grid, with a warm L2 (same launch repeated), after the optimiser's kind of L2 pollution (a 5 x 32 MiB streaming pass) and after a
plain 256 MiB flush.  python tools/k12_cold_probe.py"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoxels_b200 import ops, synth
from plenoxels_b200.trainer import VoxelTrainer
dev = torch.device("cuda:0")
sc = synth.make_scene("c2", H=64)
poses, imgs = sc.poses.to(dev), sc.imgs.to(dev)
gmin = ops.grid_origin(sc.grid.shape, sc.points_distance)
uvs = [synth.random_uv(poses.shape[0], sc.rays_per_cam, seed=i).to(dev) for i in range(16)]
tr = VoxelTrainer(sc.grid.to(dev), sc.points_distance, poses, sc.fov, imgs, sc.rays_per_cam, sc.num_samples, sc.delta_step, lr=sc.lr)
grids = {"initial ball grid": sc.grid.to(dev).clone()}
for i in range(120):
    tr.step(uvs[i % 16])
grids["after 120 training steps"] = tr.grid.clone()
junk = [torch.empty(32 << 20, dtype=torch.uint8, device=dev) for _ in range(5)]
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def k12(grid, gg, uv):
    ops.render_train(grid, gg, sc.num_samples, sc.delta_step, gmin, sc.points_distance, imgs=imgs, poses=poses, fov=sc.fov, uv=uv)
def timed(grid, pollute, n=24):
    gg = torch.zeros_like(grid)
    ts = []
    for i in range(n + 3):
        pollute(grid, gg, uvs[i % 16])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); k12(grid, gg, uvs[i % 16]); e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 2)
def flush(*_): big.add_(1)
def then(*what):
    def f(grid, gg, uv):
        big.add_(1)
        for w in what:
            {"grid": grid, "grad": gg, "uv": uv, "poses": poses}[w].sum()
        junk[0][:1024].add_(1)          # keeps the GPU busy while the host enqueues the march
    return f
for tag, grid in grids.items():
    print(json.dumps({"grid": tag, "k12_us": {
        "flush": timed(grid, then()),
        "flush, then re-read grid": timed(grid, then("grid")),
        "flush, then re-read gradient": timed(grid, then("grad")),
        "flush, then re-read grid + gradient": timed(grid, then("grid", "grad")),
        "flush, then re-read uv + poses": timed(grid, then("uv", "poses")),
        "flush, then re-read all four": timed(grid, then("grid", "grad", "uv", "poses")),
        "no flush (tiny kernel in between)": timed(grid, lambda *_: junk[0][:1024].add_(1))}}), flush=True)
