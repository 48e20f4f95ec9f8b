"""Multi-GPU parity (run under torchrun on the GPU box; launched by tests/test_gpu_multigpu.py, not collected by pytest):
   torchrun --nproc-per-node N tests/multigpu_check.py [case]
case = pull | pull_fused | pull_mc | push | push_mc | timeout      (default: pull)

Every rank renders its camera shard with the fused kernel.  Checked on rank 0 (and, for replica equality, on all ranks):
  * SUM of the shard gradients (NCCL) == single-GPU gradient of the whole batch == CPU oracle              (MULTIGPU_OK)
  * the peer-memory trainer of the chosen exchange == the NCCL all-reduce trainer after 4 steps, replicas bit-equal,
    the reported loss is the GLOBAL loss (== single-GPU loss of the whole batch), |grad| sum and Adam moments gathered
    from the slab owners == the replicated ones, step() == step_host(), and the same again with tv > 0        (PEER_OK)
  * timeout: one rank stops stepping; the other's bounded wait gives up, nothing is stored, flush() raises   (TIMEOUT_OK)
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import plenoxel_oracle as po          # noqa: E402
from plenoxels_b200 import _lib as L, ops, synth   # noqa: E402
from plenoxels_b200.trainer import PeerVoxelTrainer, VoxelTrainer, all_reduce_sum_, shard_cameras   # noqa: E402

CASES = {
    "pull": dict(exchange="pull"),
    "pull_fused": dict(exchange="pull", fused_sync=True),
    "pull_mc": dict(exchange="pull", multicast=True),          # k_adam_mc (NVLS ld_reduce / multimem.st) even at world 2
    "push": dict(exchange="push", multicast=False),
    "push_mc": dict(exchange="push", multicast=True),
}


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "pull"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    G, C, H, R, S = 48, 16, 32, 64, 128
    pd, delta = synth.GRID_EXTENT / G, 6.0 / S
    grid, poses, imgs, uv = synth.ball_grid(G), synth.lookat_poses(C), synth.random_images(C, H, H), synth.random_uv(C, R)
    gmin = ops.grid_origin(grid.shape, pd)
    cams = list(shard_cameras(C, rank, world))
    mk = dict(lr=0.0075, n_rays_global=C * R)
    common = (pd, poses[cams].to(dev), synth.CAMERA_ANGLE_X, imgs[cams].to(dev), R, S, delta)

    if case == "timeout":
        return timeout_case(rank, world, dev, grid, common, mk, C, R, cams)

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

    # ---- whole steps: NCCL all-reduce trainer vs the peer-memory trainer, without and with the TV term
    kw = CASES[case]
    solo = [dist.new_group([r]) for r in range(world)]        # one-rank groups: a trainer on one of them never communicates
    for tv in (0.0, 1e-4):
        if tv > 0 and kw.get("fused_sync"):
            continue
        ta = VoxelTrainer(grid.to(dev), *common, tv=tv, **mk)
        tb = PeerVoxelTrainer(grid.to(dev), *common, tv=tv, **kw, **mk)
        tc = PeerVoxelTrainer(grid.to(dev), *common, tv=tv, **kw, **mk)           # driven through step_host
        t1 = VoxelTrainer(grid.to(dev), pd, poses.to(dev), synth.CAMERA_ANGLE_X, imgs.to(dev), R, S, delta, lr=0.0075, tv=tv,
                          group=solo[0]) if rank == 0 else None                   # the whole batch on one GPU
        assert tb.exchange == kw["exchange"] and tb.multicast == bool(kw.get("multicast", False)), (tb.exchange, tb.multicast)
        for step in range(4):
            full_uv = synth.random_uv(C, R, seed=50 + step)
            u = full_uv[cams].to(dev)
            la, lb = float(ta.step(u)), float(tb.step(u))
            tc.step_host(full_uv[cams].contiguous().pin_memory())
            lc = tc.wait_result()
            if t1 is not None:
                l1 = float(t1.step(full_uv.to(dev)))
                assert abs(lb - l1) <= 2e-6 * abs(l1) and abs(la - l1) <= 2e-6 * abs(l1), ("global loss", la, lb, l1)
            assert abs(lb - lc) <= 1e-6 * abs(lb), ("step vs step_host loss", lb, lc)
        tb.flush(), tc.flush()
        torch.cuda.synchronize()
        dgrid = float((ta.grid - tb.grid).abs().max())
        q = float(torch.quantile((ta.grid - tb.grid).abs().flatten()[::7], 0.999))
        dhost = float((tb.grid - tc.grid).abs().max())
        dabs = float((ta.grad_abs_sum - tb.full_grad_abs_sum()).abs().max() / ta.grad_abs_sum.abs().max())
        ck = tb.checkpoint()
        dck = float((ta.grad_abs_sum.cpu() - ck["grid_grad"]).abs().max() / ta.grad_abs_sum.abs().max())
        rs = tb.resume_state()["optimizer"]
        dm = float((ta.exp_avg.cpu() - rs["exp_avg"]).abs().max() / ta.exp_avg.abs().max())
        dv = float((ta.exp_avg_sq.cpu() - rs["exp_avg_sq"]).abs().max() / ta.exp_avg_sq.abs().max())
        allsame = [torch.zeros_like(tb.grid) for _ in range(world)]
        dist.all_gather(allsame, tb.grid.contiguous())
        replicas_equal = all(torch.equal(allsame[0], t) for t in allsame)
        d1 = float((t1.grid - tb.grid).abs().max()) if t1 is not None else 0.0
        if rank == 0:
            print(f"[{case}, tv={tv:g}] peer vs nccl trainer after 4 steps: max|dgrid| {dgrid:.2e} (q99.9 {q:.2e}), vs 1 GPU {d1:.2e}, "
                  f"step vs step_host {dhost:.2e}, |grad| sum rel {dabs:.2e} (checkpoint {dck:.2e}), moments {dm:.2e}/{dv:.2e}, "
                  f"replicas equal: {replicas_equal}")
        # Adam divides by sqrt(v): a gradient that differs in summation order by 1e-7 moves a freshly touched parameter by up to
        # lr, so the bulk of the parameters is compared (99.9 % quantile) and the state arrays by their own scale
        assert q <= 1e-5 and dabs <= 1e-5 and dck <= 1e-5 and dm <= 1e-4 and dv <= 1e-4 and replicas_equal
        assert dhost <= 2e-2 and float(torch.quantile((tb.grid - tc.grid).abs().flatten()[::7], 0.999)) <= 1e-5
        del ta, tb, tc, t1
    if rank == 0:
        print("PEER_OK")
    dist.barrier()
    dist.destroy_process_group()


def timeout_case(rank, world, dev, grid, common, mk, C, R, cams):
    """Rank 1 stops stepping.  Rank 0's barrier gives up after peer_timeout_s, records the failure, the optimiser kernel
    stores nothing, flush() raises PlxError; a later wait_result() raises too instead of spinning."""
    for exchange in ("push", "pull"):
        tr = PeerVoxelTrainer(grid.to(dev), *common, exchange=exchange, peer_timeout_s=0.5, **mk)
        u = synth.random_uv(C, R, seed=7)[cams].to(dev)
        tr.step(u)
        tr.flush()
        torch.cuda.synchronize()
        dist.barrier()
        before = tr.grid.clone()
        if rank == 0:
            tr.step(u)                                   # nobody answers
            torch.cuda.synchronize()
            raised = False
            try:
                tr.flush()
            except L.PlxError as e:
                raised = "timed out" in str(e)
            assert raised, "flush() did not raise after a peer stalled"
            try:
                tr.checkpoint()
                raised = False
            except L.PlxError:
                raised = True
            assert raised, "checkpoint() did not raise after a peer stalled"
            assert torch.equal(before, tr.grid), "parameters changed although the exchange failed"
            print(f"[{exchange}] rank 0: wait gave up, PlxError raised, replica untouched")
        else:
            time.sleep(3.0)                              # "stalled" well beyond peer_timeout_s
        dist.barrier()
        del tr
    if rank == 0:
        print("TIMEOUT_OK")
    sys.stdout.flush()
    os._exit(0)                                          # the trainers are in a failed state: skip the collective teardown


if __name__ == "__main__":
    main()
