"""A/B of the Adam store skipping and launch shapes on one GPU: python tools/adam_probe.py [c2|c3]"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoxels_b200 import synth, _lib as L
from plenoxels_b200.trainer import VoxelTrainer
dev = torch.device("cuda:0")
lib = L.load()
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
sc = synth.make_scene(name, H=64)
uvs = [synth.random_uv(sc.poses.shape[0], sc.rays_per_cam, seed=i).to(dev) for i in range(16)]
def run(tag, steps=120, **tune):
    for k, v in tune.items():
        L.check(lib.plx_tune(k.encode(), v))
    tr = VoxelTrainer(sc.grid.to(dev), sc.points_distance, sc.poses.to(dev), sc.fov, sc.imgs.to(dev), sc.rays_per_cam, sc.num_samples, sc.delta_step, lr=sc.lr)
    out = {}
    for lo, hi in ((0, 10), (10, 60), (60, steps)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(lo, hi):
            tr.step(uvs[i % 16])
        e1.record(); torch.cuda.synchronize()
        out[f"steps {lo}-{hi}"] = round(e0.elapsed_time(e1) / (hi - lo) * 1e3, 1)
    print(json.dumps({"workload": name, "variant": tag, **tune, "us_per_step": out}))
for rep in range(2):
    run("skip_same=1", adam_skip_same=1, adam_blocks_per_sm=4)
    run("skip_same=0", adam_skip_same=0, adam_blocks_per_sm=4)
run("skip_same=1 bps=6", adam_skip_same=1, adam_blocks_per_sm=6)
run("skip_same=0 bps=6", adam_skip_same=0, adam_blocks_per_sm=6)
L.check(lib.plx_tune(b"adam_skip_same", 1)); L.check(lib.plx_tune(b"adam_blocks_per_sm", 4))
for pdl in (1, 0, 1, 0):
    run(f"pdl={pdl}", pdl=pdl)
