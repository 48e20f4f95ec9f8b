"""A/B of the ray-packet inference kernel on C4 (512^3 grid, one 800x800 view): lattice tiles vs lattice rows
   (the unroll / occupancy sweep behind the launch shape is profiles/r2_packet_sweep_c4.log).
   python tools/packet_probe.py"""
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
ref = {}
for mode in ("nearest", "trilinear"):
    for tile in (0, 1, 0, 1):
        L.check(lib.plx_tune(b"packet_tile", tile))
        fn = lambda: ops.render_rays(grid, o, dirs, S, delta, gmin, pd, mode=mode, clamp=False, rays_per_origin=n, coherent=True)
        ms = t(fn)
        out = fn()
        same = bool(torch.equal(out, ref.setdefault(mode, out)))
        print(json.dumps({"mode": mode, "packet_tile": tile, "ms_per_frame": round(ms, 4), "bit_equal_to_first": same}), flush=True)
L.check(lib.plx_tune(b"packet_tile", 1))
