import os, sys, torch, torch.distributed as dist, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plenoxels_b200 import synth, _lib as L
from plenoxels_b200.trainer import PeerVoxelTrainer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sc = synth.make_scene("c2", H=64)
tr = PeerVoxelTrainer(sc.grid.to(dev), sc.points_distance, sc.poses.to(dev), sc.fov, sc.imgs.to(dev), sc.rays_per_cam, sc.num_samples, sc.delta_step, lr=sc.lr, n_rays_global=sc.n_rays*world)
uv = synth.random_uv(100, 128).to(dev)
for _ in range(5): tr.step(uv)
st = L.stream_ptr(dev)
def timeit(fn, n=50):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1000
tr.step_count += 1; tr._peers[0].step = tr.step_count; tr._peer = tr._peers[0]
res = {}
def _b():
    tr._epoch += 1; tr._barrier(tr._h_grads[0], 0, st)
res["barrier"] = timeit(_b)
for U in (1,2,4):
    os.environ["X"]=str(U)
res["adam_peer"] = timeit(lambda: L.check(tr.lib.plx_adam_step_peer(C.byref(tr._peer), st)))
res['memset'] = timeit(lambda: tr.grad.zero_())
res['render'] = timeit(lambda: tr.render_phase(uv))
res['tiny_allreduce'] = timeit(lambda: dist.all_reduce(tr.loss))
res['full_step'] = timeit(lambda: tr.step(uv))
if rank == 0: print("multicast", tr.multicast, {k: round(v,1) for k,v in res.items()})
dist.destroy_process_group()
