#!/usr/bin/env python
"""bench.py — training rays/s of the voxel-grid renderer hot path on N B200s (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c1] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one ray batch: ray generation -> fused forward + MSE -> fused backward ->
[gradient all-reduce when N > 1] -> Adam (+ |grad| accumulation), i.e. the loop body of the reference's fit()
(scripts/train.py:130-184, tv = beta = 0, full resolution).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from plenoxels_b200 import synth  # noqa: E402

METRIC = "training rays/sec (fwd+bwd+optimiser step)"
UNIT = "rays/s"
MAX_DISTINCT_BATCHES = 256


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_scene(name: str) -> synth.Scene:
    return synth.make_scene(name)


def workload_label(sc: synth.Scene) -> str:
    C, H = sc.imgs.shape[0], sc.imgs.shape[1]
    return (f"{sc.name}: {sc.G}^3 grid (pd={sc.points_distance:g}), {C} views {H}x{H}, {sc.rays_per_cam} rays/view = "
            f"{sc.n_rays} rays/GPU/step, {sc.num_samples} samples/ray (delta={sc.delta_step:g}), nearest lookup, Adam lr={sc.lr}")


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Polls SM clock and clock-event reasons through NVML while the timed regions run."""

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            log(f"[bench] NVML unavailable ({e}); clocks not sampled")
            self.nv = None

    def _names(self, mask: int):
        nv = self.nv
        table = [("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"),
                 ("hw_power_brake", "nvmlClocksThrottleReasonHwPowerBrakeSlowdown"),
                 ("sync_boost", "nvmlClocksThrottleReasonSyncBoost"),
                 ("app_clocks", "nvmlClocksThrottleReasonApplicationsClocksSetting")]
        return {n for n, attr in table if hasattr(nv, attr) and (mask & getattr(nv, attr))}

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.reasons |= self._names(get(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None and self._thread is None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------- CPU reference arm
def time_reference_port(sc: synth.Scene, uvs, steps: int, warmup: int, budget_s: float = 150.0):
    """The reference's step (scripts/train.py:130-184) through the torch-CPU port on all host cores."""
    from oracle.torch_port import ReferenceStep
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = ReferenceStep(sc.grid, sc.points_distance, sc.poses, sc.fov, sc.imgs, sc.rays_per_cam, sc.num_samples,
                        sc.delta_step, sc.lr)
    n_cams = sc.poses.shape[0]
    t0 = time.perf_counter()
    ref.step(uvs[0])
    first = time.perf_counter() - t0
    # bound the whole run to ~budget_s of CPU work: if needed, each step renders a camera subset of the batch
    cams = None
    if first * (steps + warmup) > budget_s:
        keep = max(1, int(n_cams * budget_s / (first * (steps + warmup))))
        cams = torch.arange(keep)
    n_step_rays = (len(cams) if cams is not None else n_cams) * sc.rays_per_cam

    def one(i):
        u = uvs[i % len(uvs)]
        return ref.step(u if cams is None else u[cams], cams)

    for i in range(1, warmup):
        one(i)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        one(warmup + i)
        times.append(time.perf_counter() - t0)
    total = float(sum(times))
    return {"rays_per_s": n_step_rays * steps / total, "ms_per_step": 1e3 * total / steps,
            "ms_min": 1e3 * min(times), "cores": cores, "rays_per_step": n_step_rays,
            "subsampled": cams is not None}


def time_c_port(sc: synth.Scene, uvs, steps: int, warmup: int):
    """The same step through the plain-C restatement (oracle/plenoxel_oracle.c, OpenMP on all host cores): ray generation,
    forward, MSE, backward, Adam.  Reported beside the torch-CPU figure as a second, faster CPU data point."""
    from oracle import c_oracle as co
    from oracle import plenoxel_oracle as po
    imgs, poses = sc.imgs.numpy(), sc.poses.numpy()
    grid = sc.grid.numpy().copy()
    m, v, ga = np.zeros_like(grid), np.zeros_like(grid), np.zeros_like(grid)
    o = np.repeat(poses[:, :3, 3], sc.rays_per_cam, axis=0)
    gmin = po.grid_origin(grid.shape[:3], sc.points_distance)
    S, delta, pd = sc.num_samples, sc.delta_step, sc.points_distance
    times = []
    for i in range(warmup + steps):
        uv = uvs[i % len(uvs)].numpy()
        t0 = time.perf_counter()
        dirs, targets, _ = co.generate_rays(imgs, poses, sc.fov, uv)
        rgba, _, _, _ = co.render_forward(grid, o, dirs, S, delta, gmin, pd, want_lin=False)
        _, gpix = co.mse_loss(rgba, targets)
        grad = co.render_backward(grid, o, dirs, S, delta, gmin, pd, gpix).astype(np.float32)
        co.adam_step_(grid, grad, m, v, ga, sc.lr, i + 1)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = float(sum(times))
    return {"value": sc.n_rays * steps / total, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{steps} steps of {sc.n_rays} rays after {warmup} warm-up ({1e3 * total / steps:.0f} ms/step), plain-C "
                      "restatement with OpenMP (oracle/plenoxel_oracle.c)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sc = workload_scene(args.workload)
    n_batches = min(args.steps + args.warmup, MAX_DISTINCT_BATCHES)
    uvs = [synth.random_uv(sc.poses.shape[0], sc.rays_per_cam, seed=1000 + i) for i in range(n_batches)]
    r = time_reference_port(sc, uvs, args.steps, args.warmup)
    sample = (f"{'camera subset: ' if r['subsampled'] else 'full batch: '}{r['rays_per_step']} of {sc.n_rays} rays per step, "
              f"{args.steps} steps after {args.warmup} warm-up")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["rays_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(sc), "device": "host CPU", "threads": r["cores"]},
        "cpu_baseline": {"value": r["rays_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["rays_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:            # second CPU data point: the C / OpenMP restatement (never allowed to break the line)
        line["cpu_baseline"]["c_port"] = time_c_port(sc, uvs, steps=min(args.steps, 3), warmup=1)
    except Exception as e:
        log(f"[bench] C port not timed ({type(e).__name__}: {e})")
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch.distributed as dist
    from plenoxels_b200 import _lib, ops
    from plenoxels_b200.trainer import PeerVoxelTrainer, VoxelTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner) write to fd 1, so park the real stdout
    # and point fd 1 at stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    sc = workload_scene(args.workload)
    C_, R, S = sc.poses.shape[0], sc.rays_per_cam, sc.num_samples
    n_rays = sc.n_rays
    K, W = args.steps, args.warmup
    n_batches = min(K + W, MAX_DISTINCT_BATCHES)
    # every rank draws its own rays (weak scaling: per-GPU batch fixed); uv generated on the host, seeded
    uv_host = [synth.random_uv(C_, R, seed=1000 + rank * 100003 + i).pin_memory() for i in range(n_batches)]
    uv_dev = [u.to(dev) for u in uv_host]

    multi = os.environ.get("PLX_MULTI", "peer") if world > 1 else "single"
    state = {"multi": multi}

    def new_trainer():
        args_ = (sc.grid.to(dev), sc.points_distance, sc.poses.to(dev), sc.fov, imgs_dev, R, S, sc.delta_step)
        kw = dict(lr=sc.lr, n_rays_global=n_rays * world)
        if state["multi"] == "peer":
            try:
                return PeerVoxelTrainer(*args_, **kw)
            except Exception as e:          # symmetric memory unavailable on this box: NCCL all-reduce path (all ranks agree:
                log(f"[bench] peer-memory trainer unavailable ({type(e).__name__}: {e}); using the NCCL all-reduce trainer")
                ok = torch.zeros(1, device=dev)           # rendezvous failures are collective, so every rank lands here)
                dist.all_reduce(ok)
                state["multi"] = "nccl"
        return VoxelTrainer(*args_, **kw)

    imgs_dev = sc.imgs.to(dev)
    tr = new_trainer()
    cells = sc.G ** 3

    # in-bounds sample count of the timed batches (oracle-mask definition, counted by K1's exact test; untimed)
    m_in = []
    for i in range(min(n_batches, 8)):
        dirs, _ = ops.generate_rays(imgs_dev, tr.poses, sc.fov, uv=uv_dev[i], want_targets=False)
        _, cnt = ops.render_rays(tr.grid, tr.poses[:, :3, 3], dirs, S, sc.delta_step, tr.gmin, sc.points_distance,
                                 rays_per_origin=R, return_count=True)
        m_in.append(int(cnt.sum().item()))
    m_in = float(np.mean(m_in))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)

    # ---- region 1: device-resident inputs ("value")
    for i in range(W):
        tr.step(uv_dev[i % n_batches])
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        tr.step(uv_dev[(W + i) % n_batches])
    tr.flush()                                                 # multi-GPU: every replica complete (no-op on one GPU)
    e1.record()
    barrier()
    sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = n_rays * world * K / (ms_total * 1e-3)
    final_loss = float(tr.loss.item())

    # ---- region 2: end to end from pinned host buffers, loss read on the host every step ("e2e")
    tr2 = new_trainer()
    for i in range(W):
        tr2.step_host(uv_host[i % n_batches])
        tr2.wait_result()
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    host_losses = []
    for i in range(K):
        tr2.step_host(uv_host[(W + i) % n_batches])
        host_losses.append(tr2.wait_result())                  # the host reads the loss of THIS step (scripts/train.py:159)
    tr2.flush()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.stop()
    e2e_value = n_rays * world * K / e2e_s
    del tr2

    # ---- region 3: per-kernel durations in situ (same kernels, launched one by one with events in between)
    import ctypes as C
    tr3 = new_trainer()
    lib = _lib.load()
    st = _lib.stream_ptr(dev)
    fused = os.environ.get("PLX_TRAIN_FUSED", "1") != "0"
    n_inst = min(K, 64)
    for w_ in range(3):
        tr3.step(uv_dev[w_ % n_batches])
    if world > 1:
        # multi-GPU: the two phases of the step (render | gradient exchange + optimiser)
        names = ["render_train", "exchange+adam"]
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_inst)]
        barrier()
        sampler.start()
        for i in range(n_inst):
            ev = evs[i]
            ev[0].record()
            tr3.render_phase(uv_dev[(W + i) % n_batches])
            ev[1].record()
            tr3.update_phase()
            ev[2].record()
    else:
        names = ["render_train", "adam"] if fused else ["generate_rays", "render_fwd", "render_bwd", "adam"]
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(n_inst)]
        a = tr3._args
        fwd, bwd, trn = _lib.PlxRenderFwd(), _lib.PlxRenderBwd(), _lib.PlxRenderTrain()
        rays = _lib.make_rays(tr3.poses[:, :3, 3], tr3.dirs, R)
        gs, ls = 2.0 / (4.0 * n_rays * world), 1.0 / (4.0 * n_rays * world)
        fwd.march, fwd.rays, fwd.grid, fwd.rgba, fwd.tcarry = a.march, rays, a.grid, a.rgba, a.tcarry
        fwd.targets, fwd.grad_rgba, fwd.loss = a.targets, a.grad_rgba, a.loss
        fwd.grad_scale, fwd.loss_scale = gs, ls
        bwd.march, bwd.rays, bwd.grid, bwd.grad_rgba, bwd.tcarry, bwd.grad_grid = a.march, rays, a.grid, a.grad_rgba, a.tcarry, a.grad
        trn.march, trn.grid, trn.grad_grid, trn.rgba, trn.loss = a.march, a.grid, a.grad, a.rgba, a.loss
        trn.rays.n_rays = n_rays
        trn.gen.imgs, trn.gen.n_cams, trn.gen.img_h, trn.gen.img_w = a.imgs, a.n_cams, a.img_h, a.img_w
        trn.gen.poses, trn.gen.fov, trn.gen.rays_per_cam = a.poses, a.fov, R
        trn.grad_scale, trn.loss_scale = gs, ls
        sampler.start()
        for i in range(n_inst):
            u = uv_dev[(W + i) % n_batches]
            ev = evs[i]
            tr3.loss.zero_()
            if tr3._dynamic:
                tr3._work_counter.zero_()
                trn.work_counter = tr3._work_counter.data_ptr()
            ev[0].record()
            if fused:
                trn.gen.uv = u.data_ptr()
                _lib.check(lib.plx_render_train(C.byref(trn), st))
            else:
                _lib.check(lib.plx_generate_rays(a.imgs, a.n_cams, a.img_h, a.img_w, a.poses, a.fov, u.data_ptr(), R, 0, a.dirs,
                                                 a.targets, st))
                ev[1].record()
                _lib.check(lib.plx_render_fwd(C.byref(fwd), st))
                ev[2].record()
                _lib.check(lib.plx_render_bwd(C.byref(bwd), st))
            ev[-2].record()
            tr3.step_count += 1
            _lib.check(lib.plx_adam_step(a.grid, a.grad, a.exp_avg, a.exp_avg_sq, a.grad_abs_sum, cells * 4, sc.lr, 0.9, 0.999,
                                         1e-8, tr3.step_count, 1, st))
            ev[-1].record()
    torch.cuda.synchronize(dev)
    sampler.stop()
    kms = {n: float(np.mean([evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(n_inst)])) for j, n in enumerate(names)}
    del tr3

    # ---- roofline of the dominant kernel + of the whole step (SURVEY.md §8d byte model)
    peak, peak_src = measured_peak()
    alg = {"generate_rays": 8.0 * n_rays + 28.0 * n_rays, "render_fwd": 16.0 * m_in + 40.0 * n_rays,
           "render_bwd": 48.0 * m_in + 56.0 * n_rays, "render_train": 64.0 * m_in + 96.0 * n_rays, "adam": 160.0 * cells}
    timed = {k: v for k, v in kms.items() if v and k in alg}
    dom = max(timed, key=timed.get)
    achieved = alg[dom] / (timed[dom] * 1e-3) / 1e9
    traffic = None
    prof_json = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(prof_json):
        try:
            traffic = json.load(open(prof_json)).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    step_bytes = 64.0 * m_in + 96.0 * n_rays + 160.0 * cells
    step_gbs = step_bytes / (ms_total / K * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom],
                "kernel_ms": kms, "kernel_gbs": {k: alg[k] / (v * 1e-3) / 1e9 for k, v in timed.items()},
                "note": "algorithmic bytes follow SURVEY.md 8d (160 B/cell for the optimiser); K3 does not store |grad| nor "
                        "re-clear the gradient of cells no ray touched this step (32 of those 160 B/cell), so its moved bytes "
                        "(traffic) are below the model and its fraction can exceed 1",
                "step": {"algorithmic_bytes": step_bytes, "achieved": step_gbs, "frac": step_gbs / peak,
                         "model": "64*M_in + 96*N + 160*cells", "m_in": m_in, "n_rays": n_rays, "cells": cells}}

    # ---- CPU baseline (rank 0, N = 1 only): the torch-CPU port of the reference's step on the host cores
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        uvs = [synth.random_uv(C_, R, seed=1000 + i) for i in range(4)]
        r = time_reference_port(sc, uvs, steps=3, warmup=1, budget_s=30.0)
        cpu_baseline = {"value": r["rays_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                        "sample": f"3 steps of {r['rays_per_step']} rays after 1 warm-up ({r['ms_per_step']:.0f} ms/step), torch-CPU port "
                                  "of scripts/train.py:130-184 (oracle/torch_port.py)"}
        try:        # second CPU data point: the C / OpenMP restatement (never allowed to break the bench line)
            cpu_baseline["c_port"] = time_c_port(sc, uvs, steps=3, warmup=1)
        except Exception as e:
            log(f"[bench] C port not timed ({type(e).__name__}: {e})")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_label(sc),
                       "l2": "no flush: each step streams 5 grid-sized state arrays (%.0f MB) plus gathers from a %.2f GB "
                             "image set, more than the 126 MB L2; a fresh uv batch every step" %
                             (5 * cells * 16 / 1e6, sc.imgs.numel() * 4 / 1e9),
                       "parallelism": ("single GPU" if world == 1 else
                                       f"ray-sharded replicas x{world}, " +
                                       ("gradient reduce-scatter + Adam + parameter all-gather fused in one kernel over NVLink peer memory"
                                        if state["multi"] == "peer" else "dense gradient all-reduce (NCCL) + replicated Adam")),
                       "distinct_batches": n_batches, "final_loss": final_loss},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_rays * 8, "d2h_bytes_per_step": 8,
                    "ms_per_step": 1e3 * e2e_s / K,
                    "note": "per step: the march kernel reads the uv draw out of pinned host memory (zero-copy over PCIe), the "
                            "optimiser kernel stores {loss, step} into pinned host memory and the host waits for and reads "
                            "that loss before issuing the next step; images/poses/grid stay resident as in the reference "
                            "(scripts/train.py:75)",
                    "last_loss": host_losses[-1]},
            "gpu_launches": (getattr(tr, "launches_per_step", 2) if world > 1 else (2 if fused else 4)) * K,
            "clocks": sampler.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch ourselves under torchrun
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
