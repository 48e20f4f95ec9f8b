"""Multi-GPU parity (run under torchrun on the GPU box; not collected by pytest):
   torchrun --nproc-per-node N tests/multigpu_check.py
Every rank renders its camera shard with the fused kernel, gradients are SUM-all-reduced over NCCL, and the result is
compared on rank 0 with the single-GPU gradient of the whole batch (sum-order tolerance) and with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import plenoxel_oracle as po          # noqa: E402
from plenoxels_b200 import ops, synth             # noqa: E402
from plenoxels_b200.trainer import all_reduce_sum_, shard_cameras   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    G, C, H, R, S = 48, 16, 32, 64, 128
    pd, delta = synth.GRID_EXTENT / G, 6.0 / S
    grid, poses, imgs, uv = synth.ball_grid(G), synth.lookat_poses(C), synth.random_images(C, H, H), synth.random_uv(C, R)
    gmin = ops.grid_origin(grid.shape, pd)
    cams = list(shard_cameras(C, rank, world))
    gg = torch.zeros(G, G, G, 4, device=dev)
    _, loss = ops.render_train(grid.to(dev), gg, S, delta, gmin, pd, imgs=imgs[cams].to(dev), poses=poses[cams].to(dev),
                               fov=synth.CAMERA_ANGLE_X, uv=uv[cams].to(dev), n_rays_global=C * R)
    all_reduce_sum_(gg)
    all_reduce_sum_(loss)
    if rank == 0:
        g1 = torch.zeros_like(gg)
        _, loss1 = ops.render_train(grid.to(dev), g1, S, delta, gmin, pd, imgs=imgs.to(dev), poses=poses.to(dev),
                                    fov=synth.CAMERA_ANGLE_X, uv=uv.to(dev))
        err = float((gg - g1).abs().max() / g1.abs().max())
        dirs, targets, _ = po.generate_rays(imgs.numpy(), poses.numpy(), synth.CAMERA_ANGLE_X, uv.numpy())
        o = np.repeat(poses[:, :3, 3].numpy(), R, axis=0)
        rgba, _, _, _ = po.render_forward(grid.numpy(), o, dirs, S, delta, np.float32(gmin), pd)
        oloss, gpix = po.mse_loss(rgba, targets)
        ograd = po.render_backward(grid.numpy(), o, dirs, S, delta, np.float32(gmin), pd, gpix)
        oerr = float(np.abs(gg.cpu().numpy() - ograd).max() / np.abs(ograd).max())
        print(f"world={world}: grad vs 1-GPU rel err {err:.2e}, vs oracle {oerr:.2e}; loss {float(loss):.7f} / {float(loss1):.7f} / {oloss:.7f}")
        assert err <= 1e-5 and oerr <= 1e-5 and abs(float(loss) - oloss) <= 1e-5 * oloss
        print("MULTIGPU_OK")
    # ---- whole steps: NCCL all-reduce trainer vs the peer-memory trainer (fused reduce + Adam + broadcast)
    from plenoxels_b200.trainer import PeerVoxelTrainer, VoxelTrainer
    mk = dict(lr=0.0075, n_rays_global=C * R)
    common = (pd, poses[cams].to(dev), synth.CAMERA_ANGLE_X, imgs[cams].to(dev), R, S, delta)
    ta = VoxelTrainer(grid.to(dev), *common, **mk)
    tb = PeerVoxelTrainer(grid.to(dev), *common, **mk)
    for step in range(4):
        u = synth.random_uv(C, R, seed=50 + step)[cams].to(dev)
        la, lb = ta.step(u).clone(), tb.step(u).clone()
    tb.flush()
    torch.cuda.synchronize()
    dgrid = float((ta.grid - tb.grid).abs().max())
    q = float(torch.quantile((ta.grid - tb.grid).abs().flatten()[::7], 0.999))
    dabs = float((ta.grad_abs_sum - tb.gathered_grad_abs_sum()).abs().max() / ta.grad_abs_sum.abs().max())
    allsame = [torch.zeros_like(tb.grid) for _ in range(world)]
    dist.all_gather(allsame, tb.grid.contiguous())
    replicas_equal = all(torch.equal(allsame[0], t) for t in allsame)
    if rank == 0:
        print(f"peer vs nccl trainer after 4 steps: max|dgrid| {dgrid:.2e} (q99.9 {q:.2e}), |grad| sum rel {dabs:.2e}, "
              f"loss {float(la):.7f}/{float(lb):.7f}, replicas equal: {replicas_equal}")
        assert q <= 1e-5 and dabs <= 1e-5 and replicas_equal
        print("PEER_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
