"""A few multi-GPU training steps for profiling one rank under ncu (the other ranks run plain):
   RANK=r WORLD_SIZE=n LOCAL_RANK=r MASTER_ADDR=127.0.0.1 MASTER_PORT=p python tools/mg_steps.py [c2|c3] [steps] [push|pull]"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoxels_b200 import synth
from plenoxels_b200.trainer import PeerVoxelTrainer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
exchange = sys.argv[3] if len(sys.argv) > 3 else "push"
sc = synth.make_scene(name, H=64)
tr = PeerVoxelTrainer(sc.grid.to(dev), sc.points_distance, sc.poses.to(dev), sc.fov, sc.imgs.to(dev), sc.rays_per_cam, sc.num_samples,
                      sc.delta_step, lr=sc.lr, n_rays_global=sc.n_rays * world, exchange=exchange, peer_timeout_s=120.0)
uvs = [synth.random_uv(sc.poses.shape[0], sc.rays_per_cam, seed=rank * 1000 + i).to(dev) for i in range(8)]
for i in range(steps):
    tr.step(uvs[i % 8])
tr.flush(); torch.cuda.synchronize()
print(f"rank {rank}: {steps} steps of {name} ({exchange}, multicast={tr.multicast}) done, loss {float(tr.loss):.6f}", flush=True)
dist.barrier(); dist.destroy_process_group()
