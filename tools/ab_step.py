"""Time the training step of two BUILDS of the library on the same GPU in the same run (kernel-level A/B):
   python tools/ab_step.py [c2] [c3] [c2:trilinear] path/to/other_build.so [another.so ...]
Each measurement runs in its own process (PLX_AB_LIBRARY selects the build), interleaved A B A B; identical final losses show
that both builds computed the same thing."""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, torch
sys.path.insert(0, %r)
from plenoxels_b200 import synth, _lib as L
from plenoxels_b200.trainer import VoxelTrainer
dev = torch.device("cuda:0")
name, mode = sys.argv[1], sys.argv[2]
sc = synth.make_scene(name, H=64)
uvs = [synth.random_uv(sc.poses.shape[0], sc.rays_per_cam, seed=i).to(dev) for i in range(16)]
tr = VoxelTrainer(sc.grid.to(dev), sc.points_distance, sc.poses.to(dev), sc.fov, sc.imgs.to(dev), sc.rays_per_cam, sc.num_samples,
                  sc.delta_step, lr=sc.lr, mode=mode)
out = []
for lo, hi in ((0, 20), (20, 60), (60, 100), (100, 140)):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(lo, hi):
        loss = tr.step(uvs[i %% 16])
    e1.record(); torch.cuda.synchronize()
    out.append(round(e0.elapsed_time(e1) / (hi - lo) * 1e3, 2))
print(json.dumps({"workload": name, "mode": mode, "library": os.path.basename(L.lib_path()), "us_per_step": out[1:], "median": sorted(out[1:])[1],
                  "loss": float(loss), "grid_sum": float(tr.grid.double().sum())}))
''' % HERE
libs = [a for a in sys.argv[1:] if a.endswith(".so")]
specs = [a for a in sys.argv[1:] if not a.endswith(".so")] or ["c2"]
for spec in specs:
    name, _, mode = spec.partition(":")
    for rep in range(2):
        for lib in [None] + libs:
            env = dict(os.environ)
            env.pop("PLX_AB_LIBRARY", None)
            if lib:
                env["PLX_AB_LIBRARY"] = os.path.abspath(lib)
            r = subprocess.run([sys.executable, "-c", CHILD, name, mode or "nearest"], env=env, capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr[-400:], flush=True)
