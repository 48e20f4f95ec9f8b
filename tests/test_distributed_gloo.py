"""world_size-2 gloo test of the multi-GPU host logic, on CPU.

What shards: cameras (with their rays) are split into contiguous blocks per rank (`shard_cameras`); every rank holds a
grid replica, computes the gradient of ITS rays scaled by 1/(4 N_global), and the dense gradients are SUM-all-reduced
(`all_reduce_sum_`).  The per-rank gradient kernel needs a GPU, so here the oracle stands in for it (test
infrastructure): the check is that sharding + global scaling + SUM reproduce the single-process gradient and loss.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import plenoxel_oracle as po
from plenoxels_b200 import synth
from plenoxels_b200.trainer import all_reduce_sum_, shard_cameras


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene():
    G, C, H, R, S = 16, 5, 8, 24, 40
    pd, delta = synth.GRID_EXTENT / G, 6.0 / S
    grid, poses, imgs, uv = synth.ball_grid(G), synth.lookat_poses(C), synth.random_images(C, H, H), synth.random_uv(C, R)
    return G, C, H, R, S, pd, delta, grid.numpy(), poses.numpy(), imgs.numpy(), uv.numpy()


def _rank_gradient(cams, n_global):
    G, C, H, R, S, pd, delta, grid, poses, imgs, uv = _scene()
    gmin = po.grid_origin(grid.shape[:3], pd)
    cams = list(cams)
    dirs, targets, _ = po.generate_rays(imgs[cams], poses[cams], synth.CAMERA_ANGLE_X, uv[cams])
    o = np.repeat(poses[cams][:, :3, 3], R, axis=0)
    rgba, _, _, _ = po.render_forward(grid, o, dirs, S, delta, gmin, pd)
    loss, gpix = po.mse_loss(rgba, targets, n_global=n_global)
    grad = po.render_backward(grid, o, dirs, S, delta, gmin, pd, gpix)
    return loss, grad


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G, C, H, R = _scene()[:4]
        cams = shard_cameras(C, rank, world)
        loss, grad = _rank_gradient(cams, n_global=C * R)
        g = torch.from_numpy(grad.astype(np.float32))
        all_reduce_sum_(g)
        l = torch.tensor([loss], dtype=torch.float64)
        all_reduce_sum_(l)
        if rank == 0:
            torch.save({"grad": g, "loss": l}, out)
    finally:
        dist.destroy_process_group()


def test_sharded_gradient_sum_matches_single_process(tmp_path):
    world, port, out = 2, _free_port(), str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    G, C, H, R = _scene()[:4]
    loss, grad = _rank_gradient(range(C), n_global=C * R)
    assert abs(float(got["loss"]) - loss) <= 1e-12 * abs(loss) + 1e-15
    assert np.abs(got["grad"].numpy() - grad).max() <= 1e-6 * np.abs(grad).max()


def test_all_reduce_is_a_no_op_without_a_process_group():
    t = torch.arange(4.0)
    assert torch.equal(all_reduce_sum_(t.clone()), t)
