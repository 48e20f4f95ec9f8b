"""Small single-GPU drivers for ncu captures of kernels bench.py's headline loop does not launch:
   python tools/profile_targets.py tri      C2 training steps with the trilinear lookup      (k_render_train<1,...>)
   python tools/profile_targets.py push1    C2 fused march, push exchange into 8 local slab buffers (k_render_train<...,PEER>)
   python tools/profile_targets.py c4       C4 inference frames, nearest + trilinear, uint8 image epilogue (k_render_fwd_packet)
   python tools/profile_targets.py splat    GPU point splat of a 256^3 grid                   (k_splat_points / k_splat_resolve)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoxels_b200 import ops, synth
from plenoxels_b200.trainer import VoxelTrainer
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "tri"
if what == "tri":
    sc = synth.make_scene("c2", H=64)
    tr = VoxelTrainer(sc.grid.to(dev), sc.points_distance, sc.poses.to(dev), sc.fov, sc.imgs.to(dev), sc.rays_per_cam, sc.num_samples,
                      sc.delta_step, lr=sc.lr, mode="trilinear")
    for i in range(8):
        tr.step(synth.random_uv(sc.poses.shape[0], sc.rays_per_cam, seed=i).to(dev))
elif what == "push1":
    sc = synth.make_scene("c2", H=64)
    grid = sc.grid.to(dev)
    bufs = [torch.zeros_like(grid) for _ in range(8)]
    gmin = ops.grid_origin(grid.shape, sc.points_distance)
    for i in range(8):
        ops.render_train(grid, bufs[0], sc.num_samples, sc.delta_step, gmin, sc.points_distance, imgs=sc.imgs.to(dev), poses=sc.poses.to(dev),
                         fov=sc.fov, uv=synth.random_uv(sc.poses.shape[0], sc.rays_per_cam, seed=i).to(dev), peer_grads=bufs)
elif what == "c4":
    G, S, delta, side = 512, 600, 0.01, 800
    pd = synth.GRID_EXTENT / G
    grid = synth.ball_grid(G).to(dev).clip_(0, 1)
    grid[..., 3][grid[..., 3] < 0.2] = 0.0
    pose = synth.lookat_poses(4)[1:2].to(dev)
    gmin = ops.grid_origin(grid.shape, pd)
    for mode in ("nearest", "trilinear"):
        for _ in range(3):
            ops.render_image_u8(grid, pose, synth.CAMERA_ANGLE_X, side, S, delta, gmin, pd, mode=mode)
elif what == "splat":
    import time
    from plenoxels_b200 import visualization as vz
    G = 256
    grid = synth.ball_grid(G).to(dev).clip_(0, 1)
    pose = synth.lookat_poses(4)[1].to(dev)
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        img = vz.visulize_3d_in_2d_fast(grid, synth.GRID_EXTENT / G, pose, synth.CAMERA_ANGLE_X, 500)
        print("splat 256^3 -> 500x500, host to host:", round((time.perf_counter() - t0) * 1e3, 2), "ms", img.shape)
torch.cuda.synchronize()
print("done", what)
