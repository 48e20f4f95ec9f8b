"""A/B of a library tuning switch (plx_tune) on the training step, one process, interleaved:
   python tools/tune_probe.py c3 adam_two_phase 0 1"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoxels_b200 import synth, _lib as L
from plenoxels_b200.trainer import VoxelTrainer
dev = torch.device("cuda:0"); lib = L.load()
name, knob, values = sys.argv[1], sys.argv[2], [int(v) for v in sys.argv[3:]]
sc = synth.make_scene(name, H=64)
uvs = [synth.random_uv(sc.poses.shape[0], sc.rays_per_cam, seed=i).to(dev) for i in range(16)]
def run(v, steps=140):
    L.check(lib.plx_tune(knob.encode(), v))
    tr = VoxelTrainer(sc.grid.to(dev), sc.points_distance, sc.poses.to(dev), sc.fov, sc.imgs.to(dev), sc.rays_per_cam, sc.num_samples, sc.delta_step, lr=sc.lr)
    out = []
    for lo, hi in ((0, 20), (20, 60), (60, 100), (100, steps)):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(lo, hi): loss = tr.step(uvs[i % 16])
        e1.record(); torch.cuda.synchronize(); out.append(round(e0.elapsed_time(e1) / (hi - lo) * 1e3, 2))
    print(json.dumps({"workload": name, knob: v, "us_per_step": out, "loss": float(loss), "grid_sum": float(tr.grid.double().sum()),
                      "grad_abs_sum": float(tr.grad_abs_sum.double().sum())}), flush=True)
run(values[0])
for rep in range(2):
    for v in values: run(v)
L.check(lib.plx_tune(knob.encode(), -1))
