import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoxels_b200 import ops, synth, _lib as L
dev = torch.device("cuda:0"); lib = L.load()
G, S, delta, side = 512, 600, 0.01, 800
pd = synth.GRID_EXTENT / G
grid = synth.ball_grid(G).to(dev).clip_(0, 1); grid[..., 3][grid[..., 3] < 0.2] = 0.0
pose = synth.lookat_poses(4)[1:2].to(dev); gmin = ops.grid_origin(grid.shape, pd); n = side * side
dirs, _ = ops.generate_rays(None, pose, synth.CAMERA_ANGLE_X, uv=None, rays_per_cam=n, want_targets=False)
o = pose[:, :3, 3]
def t(fn, reps=8):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for v in (0, 1, 2, 3, 4, 5, 0):
    L.check(lib.plx_tune(b"packet_variant", v))
    ms = t(lambda: ops.render_rays(grid, o, dirs, S, delta, gmin, pd, clamp=False, rays_per_origin=n, coherent=True))
    print(json.dumps({"packet_variant": v, "ms_per_frame": round(ms, 4)}))
