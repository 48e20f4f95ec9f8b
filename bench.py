#!/usr/bin/env python
"""bench.py — training rays/s of the voxel-grid renderer hot path on N B200s (contract: see DESIGN.md §4 Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c1] [--impl reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one ray batch: fused march (ray generation + forward + MSE + backward) ->
[gradient exchange when N > 1] -> Adam (+ |grad| accumulation), i.e. the loop body of the reference's fit()
(scripts/train.py:130-184, tv = beta = 0, full resolution).  Prints ONE JSON line on rank 0.

Headline = BASELINE.json config #2 (C2).  The same line carries, under "extra", the other BASELINE configs measured in the
same run: C3 at the same N (config #3), and at N = 1 the C4 inference render (config #4, nearest + trilinear), three C5
sweep points (config #5) and trilinear training on C2 — each with its own roofline; at N > 1 also a self-check of the
multi-GPU exchange (replica bit-equality, gradient against the NCCL path) made in the untimed region.
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
MAX_DISTINCT_BATCHES = 64
REPEATS = 9                 # the K-step timed region is repeated; value = the median region
NVLINK_GBS = 770.0          # measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md)
L2_NOTE = ("no flush: each step streams 5 grid-sized state arrays plus gathers from the image set, more than the 126 MB L2 "
           "(c2: 168 MB + 1.02 GB, c3: 1.3 GB + 0.33 GB); a fresh uv batch every step")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_label(sc: synth.Scene) -> str:
    C, H = sc.imgs.shape[0], sc.imgs.shape[1]
    return (f"{sc.name}: {sc.G}^3 grid (pd={sc.points_distance:g}), {C} views {H}x{H}, {sc.rays_per_cam} rays/view = "
            f"{sc.n_rays} rays/GPU/step, {sc.num_samples} samples/ray (delta={sc.delta_step:g}), nearest lookup, Adam lr={sc.lr}")


def bench_config(sc: synth.Scene) -> dict:
    """`config` names the workload and nothing run-specific, so both arms (and every N) print the same object."""
    return {"workload": workload_label(sc), "l2": L2_NOTE}


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

    def nvlink_kib(self):
        """Aggregate NVLink payload / raw byte counters of this GPU in KiB (NVML field values, scope = all links), or None."""
        if self.nv is None:
            return None
        nv = self.nv
        try:
            ids = [(nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xffffffff), (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xffffffff),
                   (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX, 0xffffffff), (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_RX, 0xffffffff)]
            vals = nv.nvmlDeviceGetFieldValues(self.h, ids)
            out = {}
            for name, v in zip(("data_tx", "data_rx", "raw_tx", "raw_rx"), vals):
                if v.nvmlReturn != 0:
                    return None
                out[name] = int(v.value.ullVal)
            return out
        except Exception as e:      # noqa: BLE001
            log(f"[bench] NVLink counters unavailable ({type(e).__name__}: {e})")
            return None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------- CPU reference arm
NOMINAL_CPU_STEP_S = 0.35           # a C2 step of the torch-CPU port on this pool's 16 host cores


def reference_cameras(sc: synth.Scene, steps: int, warmup: int, budget_s: float):
    """The camera subset the CPU arm renders per step: ALL cameras unless K + W full steps would blow the time budget at the
    nominal step time — a function of (K, W) only, so every N of a scaling run times exactly the same batch."""
    n_cams = sc.poses.shape[0]
    need = NOMINAL_CPU_STEP_S * (sc.n_rays / 12800.0) * (steps + warmup)
    if need <= budget_s:
        return None
    return torch.arange(max(1, int(n_cams * budget_s / need)))


def time_reference_port(sc: synth.Scene, uvs, steps: int, warmup: int, budget_s: float = 170.0):
    """The reference's step (scripts/train.py:130-184) through the torch-CPU port on all host cores."""
    from oracle.torch_port import ReferenceStep
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = ReferenceStep(sc.grid, sc.points_distance, sc.poses, sc.fov, sc.imgs, sc.rays_per_cam, sc.num_samples,
                        sc.delta_step, sc.lr)
    cams = reference_cameras(sc, steps, warmup, budget_s)
    n_step_rays = (len(cams) if cams is not None else sc.poses.shape[0]) * sc.rays_per_cam

    def one(i):
        u = uvs[i % len(uvs)]
        return ref.step(u if cams is None else u[cams], cams)

    for i in range(warmup):
        one(i)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        one(warmup + i)
        times.append(time.perf_counter() - t0)
    total = float(sum(times))
    return {"rays_per_s": n_step_rays * steps / total, "ms_per_step": 1e3 * total / steps,
            "ms_min": 1e3 * min(times), "ms_max": 1e3 * max(times), "rays_per_s_best": n_step_rays / min(times), "cores": cores,
            "rays_per_step": n_step_rays, "subsampled": cams is not None}


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
    sc = synth.make_scene(args.workload)
    n_batches = min(args.steps + args.warmup, MAX_DISTINCT_BATCHES)
    uvs = [synth.random_uv(sc.poses.shape[0], sc.rays_per_cam, seed=1000 + i) for i in range(n_batches)]
    r = time_reference_port(sc, uvs, args.steps, args.warmup)
    sample = (f"{'camera subset: ' if r['subsampled'] else 'full batch: '}{r['rays_per_step']} of {sc.n_rays} rays per step, "
              f"{args.steps} steps after {args.warmup} warm-up; step time min/mean/max {r['ms_min']:.0f}/{r['ms_per_step']:.0f}/{r['ms_max']:.0f} ms")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["rays_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(sc),
        "run": {"device": "host CPU", "threads": r["cores"], "best_step_rays_per_s": r["rays_per_s_best"]},
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


# ----------------------------------------------------------------------------------------------------- GPU arm helpers
class Ctx:
    """Process-wide bits every measurement needs."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local_rank)
        self.peak, self.peak_src = measured_peak()
        self.multi = os.environ.get("PLX_MULTI", "peer") if self.world > 1 else "single"     # peer | nccl
        self.exchange = os.environ.get("PLX_EXCHANGE") or None                               # push | pull | None = auto
        mc = os.environ.get("PLX_MULTICAST")                                                 # 0 | 1 | unset = the trainer's default
        self.multicast = None if mc is None else mc == "1"
        self.sampler = ClockSampler(self.local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def make_trainer(cx: Ctx, sc: synth.Scene, dev_scene: dict, **kw):
    from plenoxels_b200.trainer import PeerVoxelTrainer, VoxelTrainer
    args_ = (dev_scene["grid"], sc.points_distance, dev_scene["poses"], sc.fov, dev_scene["imgs"], sc.rays_per_cam, sc.num_samples,
             sc.delta_step)
    k = dict(lr=sc.lr, n_rays_global=sc.n_rays * cx.world)
    k.update(kw)
    if cx.multi == "peer":
        try:
            return PeerVoxelTrainer(*args_, exchange=cx.exchange, multicast=cx.multicast, **k)
        except Exception as e:          # symmetric memory unavailable on this box: NCCL all-reduce path (rendezvous failures are
            log(f"[bench] peer-memory trainer unavailable ({type(e).__name__}: {e}); using the NCCL all-reduce trainer")
            ok = torch.zeros(1, device=cx.dev)           # collective, so every rank lands here together)
            cx.dist.all_reduce(ok)
            cx.multi = "nccl"
    return VoxelTrainer(*args_, **k)


def timed_regions(cx: Ctx, step_fn, flush_fn, K: int, W: int, repeats: int):
    """W warm-up steps, then `repeats` regions of exactly K steps, each bracketed by barrier + synchronize, CUDA events on the
    launching stream, max over ranks.  Returns the per-region totals in ms."""
    i = 0
    for _ in range(W):
        step_fn(i)
        i += 1
    out = []
    for _ in range(repeats):
        cx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            step_fn(i)
            i += 1
        flush_fn()
        e1.record()
        cx.barrier()
        out.append(cx.max_over_ranks(e0.elapsed_time(e1)))
    return out


def spread(ms_list, K):
    a = np.asarray(ms_list) / K
    return {"n": len(ms_list), "ms_per_step_median": float(np.median(a)), "ms_per_step_min": float(a.min()),
            "ms_per_step_max": float(a.max())}


def count_in_bounds(sc, tr, dev_scene, uv_dev, n=4):
    from plenoxels_b200 import ops
    m_in = []
    for i in range(min(len(uv_dev), n)):
        dirs, _ = ops.generate_rays(dev_scene["imgs"], tr.poses, sc.fov, uv=uv_dev[i], want_targets=False)
        _, cnt = ops.render_rays(tr.grid, tr.poses[:, :3, 3], dirs, sc.num_samples, sc.delta_step, tr.gmin, sc.points_distance,
                                 rays_per_origin=sc.rays_per_cam, return_count=True)
        m_in.append(int(cnt.sum().item()))
    return float(np.mean(m_in))


def phase_times(cx: Ctx, tr, uv_dev, n_inst: int):
    """Device time of the step's kernels in situ.  Events BETWEEN the kernels of a step distort a multi-GPU step (they book
    rank skew and host issue gaps to whichever phase follows), so each kernel is timed the robust way: the same launch, repeated
    back to back on every rank at once with one event pair around the loop — the march under everybody's reduction traffic,
    the optimiser / exchange kernel under everybody's parameter stores.  The trainer is a throw-away (its state is not a
    training state afterwards).  What the full step costs beyond the two kernels (barriers, launch gaps, rank skew) is
    reported by the caller as step - sum."""
    peer = hasattr(tr, "exchange")
    opt_name = ("slab_adam+allgather" if tr.exchange == "push" else "exchange+adam") if peer else ("adam" if cx.world == 1 else "exchange+adam")
    for w_ in range(3):
        tr.step(uv_dev[w_ % len(uv_dev)])
    tr.flush()

    def loop(fn):
        cx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_inst):
            fn(i)
        e1.record()
        cx.barrier()
        return cx.max_over_ranks(e0.elapsed_time(e1)) / n_inst

    if cx.world == 1:
        # one GPU: no rank skew to distort anything, so the kernels are timed inside real steps (the march right behind an
        # optimiser pass that has just streamed 168 MB through the L2, and vice versa), events between them; a device-side
        # head start lets the host queue the whole loop first so that no interval contains a host issue gap
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_inst)]
        torch.cuda.synchronize(cx.dev)
        torch.cuda._sleep(int(2.0e6 * n_inst))
        for i in range(n_inst):
            evs[i][0].record()
            tr.render_phase(uv_dev[i % len(uv_dev)])
            evs[i][1].record()
            tr.update_phase()
            evs[i][2].record()
        torch.cuda.synchronize(cx.dev)
        return {n: float(np.mean([e[j].elapsed_time(e[j + 1]) for e in evs])) for j, n in enumerate(["render_train", opt_name])}
    if peer:
        tr._no_barriers = True
    out = {"render_train": loop(lambda i: tr.render_phase(uv_dev[i % len(uv_dev)]))}
    if peer or cx.world == 1:
        out[opt_name] = loop(lambda i: tr.update_phase())
    else:                                        # NCCL trainer: all-reduce + replicated Adam
        out[opt_name] = loop(lambda i: tr.update_phase())
    if peer:
        tr._no_barriers = False
    return out


def roofline_of(cx: Ctx, kms: dict, m_in: float, n_rays: int, cells: int, tr, workload: str, ms_step: float):
    """Roofline of the dominant phase (HBM byte model of SURVEY.md 8d; NVLink bytes per direction for the exchange) and of the
    whole step."""
    W = cx.world
    alg = {"render_train": 64.0 * m_in + 96.0 * n_rays, "adam": 160.0 * cells}
    nv = {}
    if W > 1:
        f = (W - 1) / W
        mc = bool(getattr(tr, "multicast", False))
        exch = getattr(tr, "exchange", "nccl")
        if exch == "push":
            alg["slab_adam+allgather"] = 160.0 * cells / W + 16.0 * cells * f       # own slab + the peers' parameter stores landing here
            nv["slab_adam+allgather"] = 16.0 * cells * f                              # inbound parameters per GPU
            nv["render_train"] = 16.0 * m_in * f                                      # outbound reductions (upper bound: before run merging)
        else:
            alg["exchange+adam"] = 160.0 * cells / W + 16.0 * cells * f
            nv["exchange+adam"] = 16.0 * cells * (1.0 if mc else 2.0 * f)              # gradients in (reduced by the switch or per peer) + parameters in
    timed = {k: v for k, v in kms.items() if v and k in alg}
    kms = dict(kms, **{"step": ms_step, "barriers+gaps": ms_step - sum(timed.values())})
    dom = max(timed, key=timed.get)
    hbm_gbs = alg[dom] / (timed[dom] * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(REPO, "profiles", "traffic.json"))).get(workload, {}).get(dom)
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": hbm_gbs, "peak": cx.peak, "unit": "GB/s", "frac": hbm_gbs / cx.peak,
            "traffic": traffic, "traffic_source": "static: in-situ ncu capture (--cache-control none) of this kernel on this workload, N = 1 (profiles/traffic.json)"
            if traffic else None,
            "peak_source": cx.peak_src, "algorithmic_bytes_per_launch": alg[dom], "kernel_ms": kms,
            "kernel_gbs": {k: alg[k] / (v * 1e-3) / 1e9 for k, v in timed.items()},
            "note": "algorithmic bytes follow SURVEY.md 8d (160 B/cell for the optimiser); the optimiser does not store |grad| nor "
                    "re-clear the gradient of cells no ray touched this step (32 of those 160 B/cell), so its moved bytes (traffic) are "
                    "below the model and its fraction can exceed 1"}
    if dom in nv:
        nv_gbs = nv[dom] / (timed[dom] * 1e-3) / 1e9
        if nv[dom] / NVLINK_GBS > alg[dom] / cx.peak:        # the NVLink floor of this phase is above its HBM floor
            roof.update({"bound": "nvlink", "achieved": nv_gbs, "peak": NVLINK_GBS, "frac": nv_gbs / NVLINK_GBS,
                         "peak_source": "measured peer copy per direction per GPU (B200_PROFILING.md)",
                         "algorithmic_bytes_per_launch": nv[dom], "hbm_achieved": hbm_gbs})
    if nv:
        roof["nvlink"] = {k: {"bytes_per_direction": b, "achieved": b / (kms[k] * 1e-3) / 1e9, "peak": NVLINK_GBS,
                              "frac": b / (kms[k] * 1e-3) / 1e9 / NVLINK_GBS} for k, b in nv.items() if kms.get(k)}
    step_bytes = 64.0 * m_in + 96.0 * n_rays + 160.0 * cells
    step_gbs = step_bytes / (ms_step * 1e-3) / 1e9
    roof["step"] = {"algorithmic_bytes": step_bytes, "achieved": step_gbs, "frac": step_gbs / cx.peak,
                    "model": "64*M_in + 96*N + 160*cells (per GPU)", "m_in": m_in, "n_rays": n_rays, "cells": cells}
    return roof


def to_device(sc: synth.Scene, dev):
    return {"grid": sc.grid.to(dev), "poses": sc.poses.to(dev), "imgs": sc.imgs.to(dev)}


def measure_training(cx: Ctx, sc: synth.Scene, K: int, W: int, repeats: int, e2e: bool, dev_scene=None, **trainer_kw):
    """value / e2e / per-phase numbers of one training workload on the current process group."""
    dev_scene = dev_scene or to_device(sc, cx.dev)
    C_, R = sc.poses.shape[0], sc.rays_per_cam
    n_batches = min(K + W, MAX_DISTINCT_BATCHES)
    # every rank draws its own rays (weak scaling: per-GPU batch fixed); uv generated on the host, seeded
    uv_host = [synth.random_uv(C_, R, seed=1000 + cx.rank * 100003 + i).pin_memory() for i in range(n_batches)]
    uv_dev = [u.to(cx.dev) for u in uv_host]
    tr = make_trainer(cx, sc, dev_scene, **trainer_kw)
    cells = sc.G ** 3
    m_in = count_in_bounds(sc, tr, dev_scene, uv_dev)
    out = {}
    # ---- region 1: device-resident inputs ("value")
    cx.sampler.start()
    ms = timed_regions(cx, lambda i: tr.step(uv_dev[i % n_batches]), tr.flush, K, W, repeats)
    cx.sampler.stop()
    ms_med = float(np.median(ms))
    out.update(value=sc.n_rays * cx.world * K / (ms_med * 1e-3), ms_per_step=ms_med / K, repeats=spread(ms, K),
               final_loss=float(tr.loss.item()), launches_per_step=tr.launches_per_step)
    # ---- NVLink traffic actually moved per step (NVML byte counters of this GPU around a long untimed run of the same steps)
    if cx.world > 1 and cx.sampler.nvlink_kib() is not None:      # N/A on this pool's B200 boxes (NVML_ERROR_NOT_SUPPORTED): then skipped
        n_nv = 1500
        cx.barrier()
        time.sleep(0.3)
        nv0 = cx.sampler.nvlink_kib()
        t0 = time.perf_counter()
        for i in range(n_nv):
            tr.step(uv_dev[i % n_batches])
        tr.flush()
        cx.barrier()
        wall = time.perf_counter() - t0
        time.sleep(0.3)
        nv1 = cx.sampler.nvlink_kib()
        if nv0 and nv1:
            per = {k: (nv1[k] - nv0[k]) * 1024.0 / n_nv for k in nv0}
            out["nvlink_measured"] = {"bytes_per_step": per, "steps": n_nv, "ms_per_step_wall": 1e3 * wall / n_nv,
                                      "source": "NVML NVLINK_THROUGHPUT_DATA / RAW counters of rank 0's GPU, all links",
                                      "raw_gbs": {d: per["raw_" + d] / (out["ms_per_step"] * 1e-3) / 1e9 for d in ("tx", "rx")},
                                      "data_gbs": {d: per["data_" + d] / (out["ms_per_step"] * 1e-3) / 1e9 for d in ("tx", "rx")}}
    # ---- region 2: end to end from pinned host buffers, loss read on the host every step ("e2e")
    if e2e:
        tr2 = make_trainer(cx, sc, dev_scene, **trainer_kw)
        losses = []

        def host_step(i):
            tr2.step_host(uv_host[i % n_batches])
            losses.append(tr2.wait_result())          # the host reads the loss of THIS step (scripts/train.py:159)

        for i in range(W):
            host_step(i)
        walls = []
        cx.sampler.start()
        for rep in range(repeats):
            cx.barrier()
            t0 = time.perf_counter()
            for i in range(K):
                host_step(W + rep * K + i)
            tr2.flush()
            cx.barrier()
            walls.append(cx.max_over_ranks(time.perf_counter() - t0) * 1e3)
        cx.sampler.stop()
        w_med = float(np.median(walls))
        out["e2e"] = {"value": sc.n_rays * cx.world * K / (w_med * 1e-3), "unit": UNIT, "h2d_bytes_per_step": sc.n_rays * 8,
                      "d2h_bytes_per_step": 8, "ms_per_step": w_med / K, "repeats": spread(walls, K), "last_loss": losses[-1],
                      "note": "per step: the march kernel reads the uv draw out of pinned host memory (zero-copy over PCIe), the "
                              "optimiser kernel stores {loss, step} into pinned host memory and the host waits for and reads "
                              "that loss before issuing the next step; images/poses/grid stay resident as in the reference "
                              "(scripts/train.py:75); wall clock, max over ranks, median of the repeated K-step regions"}
        del tr2
    # ---- region 3: per-phase durations in situ
    tr3 = make_trainer(cx, sc, dev_scene, **trainer_kw)
    cx.sampler.start()
    kms = phase_times(cx, tr3, uv_dev, min(K, 64))
    cx.sampler.stop()
    out["roofline"] = roofline_of(cx, kms, m_in, sc.n_rays, cells, tr3, sc.name, out["ms_per_step"])
    if "nvlink_measured" in out:
        out["roofline"]["nvlink_measured"] = out.pop("nvlink_measured")
    out["parallelism"] = ("single GPU" if cx.world == 1 else f"ray-sharded replicas x{cx.world}, " + (
        {"push": "push exchange: the march reduces every touched cell into the slab owner's buffer over NVLink, slab Adam, parameters "
                 "stored to every replica" + (" through NVLS multicast" if tr3.multicast else " per peer"),
         "pull": "pull exchange: gradient reduce-scatter + Adam + parameter all-gather fused in one kernel over NVLink peer memory" +
                 (" (NVLS in-switch reduce / multicast)" if getattr(tr3, "multicast", False) else "")}[tr3.exchange]
        if hasattr(tr3, "exchange") else "dense gradient all-reduce (NCCL) + replicated Adam"))
    del tr, tr3
    return out


def multi_gpu_selfcheck(cx: Ctx, sc: synth.Scene, dev_scene):
    """Untimed: (1) after a few steps of the product multi-GPU trainer all replicas hold the same bits; (2) the gradient it
    applied equals the NCCL path's: both trainers take ONE step from the same state and their first Adam moment
    (= (1 - beta1) * gradient after step 1) is compared over the whole grid."""
    from plenoxels_b200.trainer import VoxelTrainer
    dist = cx.dist
    C_, R = sc.poses.shape[0], sc.rays_per_cam
    uvs = [synth.random_uv(C_, R, seed=777 + cx.rank * 131 + i).to(cx.dev) for i in range(3)]
    tp = make_trainer(cx, sc, dev_scene)
    tn = VoxelTrainer(dev_scene["grid"], sc.points_distance, dev_scene["poses"], sc.fov, dev_scene["imgs"], R, sc.num_samples,
                      sc.delta_step, lr=sc.lr, n_rays_global=sc.n_rays * cx.world)
    lp, ln = float(tp.step(uvs[0])), float(tn.step(uvs[0]))
    tp.flush()
    torch.cuda.synchronize(cx.dev)
    m_peer, _ = tp._full_moments()
    num = float((m_peer - tn.exp_avg).abs().max())
    den = float(tn.exp_avg.abs().max())
    for u in uvs[1:]:
        tp.step(u)
    tp.flush()
    torch.cuda.synchronize(cx.dev)
    ref = tp.grid.detach().clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([1.0 if torch.equal(ref, tp.grid) else 0.0], device=cx.dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out = {"replicas_bit_equal": bool(same.item() == 1.0), "grad_relerr_vs_nccl_path": num / max(den, 1e-30),
           "global_loss_peer": lp, "global_loss_nccl": ln, "exchange": getattr(tp, "exchange", "nccl"),
           "multicast": bool(getattr(tp, "multicast", False)), "steps_checked": 3, "workload": sc.name}
    del tp, tn
    return out


# ----------------------------------------------------------------------------------------------------- other BASELINE configs
def timed_call(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def extra_c4(cx: Ctx):
    """BASELINE config #4: 512^3 grid, inference-only render of one 800x800 view (even-spread rays, S = 600, delta = 0.01,
    alpha threshold 0.2 as scripts/compare_inference_to_image.py:91): the ray-packet kernel behind visulize_3d_in_2d."""
    from plenoxels_b200 import ops
    dev = cx.dev
    G, S, delta, side = 512, 600, 0.01, 800
    pd = synth.GRID_EXTENT / G
    grid = synth.ball_grid(G).to(dev).clip_(0, 1)
    grid[..., 3][grid[..., 3] < 0.2] = 0.0
    poses = synth.lookat_poses(4)[1:2].to(dev)
    gmin = ops.grid_origin(grid.shape, pd)
    n = side * side
    dirs, _ = ops.generate_rays(None, poses, synth.CAMERA_ANGLE_X, uv=None, rays_per_cam=n, want_targets=False)
    o = poses[:, :3, 3]
    out = {"workload": f"c4: {G}^3 grid, one {side}x{side} view = {n} rays, {S} samples/ray (delta={delta}), inference only"}
    for mode in ("nearest", "trilinear"):
        _, cnt = ops.render_rays(grid, o, dirs, S, delta, gmin, pd, mode=mode, clamp=False, rays_per_origin=n, return_count=True)
        m_in = int(cnt.sum())
        ms = timed_call(lambda: ops.render_rays(grid, o, dirs, S, delta, gmin, pd, mode=mode, clamp=False, rays_per_origin=n,
                                                coherent=True), n=8)
        ms_img = timed_call(lambda: ops.render_image_u8(grid, poses, synth.CAMERA_ANGLE_X, side, S, delta, gmin, pd, mode=mode), n=8) \
            if hasattr(ops, "render_image_u8") else None
        alg = (16.0 * (8 if mode == "trilinear" else 1)) * m_in + 40.0 * n
        gbs = alg / (ms * 1e-3) / 1e9
        out[mode] = {"ms_per_frame": ms, "Mrays_per_s": n / ms / 1e3, "m_in": m_in,
                     "ms_per_frame_raygen+march+uint8_image": ms_img,
                     "roofline": {"bound": "hbm", "kernel": "k_render_fwd_packet", "achieved": gbs, "peak": cx.peak, "unit": "GB/s",
                                  "frac": gbs / cx.peak, "algorithmic_bytes_per_launch": alg,
                                  "model": ("128" if mode == "trilinear" else "16") + "*M_in + 40*N"}}
    del grid
    torch.cuda.empty_cache()
    return out


def extra_c5(cx: Ctx):
    """BASELINE config #5 (three points of the sweep): random ray batches on a 256^3 grid, forward march and fused training
    march, HBM GB/s against the roofline."""
    from plenoxels_b200 import ops
    dev = cx.dev
    G = 256
    pd = synth.GRID_EXTENT / G
    grid = synth.ball_grid(G).to(dev)
    gg = torch.zeros_like(grid)
    gmin = ops.grid_origin(grid.shape, pd)
    poses = synth.lookat_poses(64).to(dev)
    imgs = torch.rand(64, 32, 32, 4, device=dev)
    pts = []
    for logn, S in ((16, 256), (20, 256), (20, 512)):
        n, delta = 1 << logn, 6.0 / S
        R = n // 64
        uv = torch.rand(64, R, 2, device=dev, generator=torch.Generator(device=dev).manual_seed(logn))
        dirs, targets = ops.generate_rays(imgs, poses, synth.CAMERA_ANGLE_X, uv=uv)
        o = poses[:, :3, 3]
        _, cnt = ops.render_rays(grid, o, dirs, S, delta, gmin, pd, rays_per_origin=R, return_count=True)
        m_in = int(cnt.sum())
        ms_f = timed_call(lambda: ops.render_rays(grid, o, dirs, S, delta, gmin, pd, rays_per_origin=R), n=6)
        ms_t = timed_call(lambda: ops.render_train(grid, gg, S, delta, gmin, pd, origins=o, dirs=dirs, targets=targets,
                                                   rays_per_origin=R), n=6)
        gf, gt = (16.0 * m_in + 40.0 * n) / (ms_f * 1e-3) / 1e9, (64.0 * m_in + 96.0 * n) / (ms_t * 1e-3) / 1e9
        pts.append({"rays": n, "S": S, "m_in": m_in, "fwd_ms": ms_f, "fwd_Mrays_s": n / ms_f / 1e3, "train_ms": ms_t,
                    "train_Mrays_s": n / ms_t / 1e3,
                    "roofline_fwd": {"bound": "hbm", "achieved": gf, "peak": cx.peak, "frac": gf / cx.peak, "model": "16*M_in + 40*N"},
                    "roofline_train": {"bound": "hbm", "achieved": gt, "peak": cx.peak, "frac": gt / cx.peak, "model": "64*M_in + 96*N"}})
    del grid, gg
    torch.cuda.empty_cache()
    return {"workload": "c5: 256^3 grid (10 % ball), random rays from 64 cameras", "points": pts}


# ----------------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    from plenoxels_b200 import _lib

    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner) write to fd 1, so park the real stdout
    # and point fd 1 at stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    cx = Ctx()
    torch.cuda.set_device(cx.local_rank)
    if cx.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        cx.dist.init_process_group("nccl", device_id=cx.dev)
    lib = _lib.load()
    for kv in args.tune:            # A/B runs only: library tuning switches (plx_tune); the defaults are the measured optimum
        k, v = kv.split("=")
        _lib.check(lib.plx_tune(k.encode(), int(v)), f"plx_tune({kv})")
    K, W = args.steps, args.warmup

    sc = synth.make_scene(args.workload)
    dev_scene = to_device(sc, cx.dev)
    head = measure_training(cx, sc, K, W, REPEATS, e2e=True, dev_scene=dev_scene)
    clocks = cx.sampler.summary()
    extra = {}

    def guarded(name, fn):
        """An extra record must never break the headline line; a failure is reported in its place (all ranks agree on it)."""
        try:
            extra[name] = fn()
        except Exception as e:      # noqa: BLE001
            log(f"[bench] extra '{name}' failed: {type(e).__name__}: {e}")
            extra[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()

    if cx.world > 1 and cx.multi == "peer":
        guarded("selfcheck", lambda: multi_gpu_selfcheck(cx, sc, dev_scene))
    del dev_scene
    torch.cuda.empty_cache()

    if not args.no_extras:
        if args.workload != "c3":
            def c3():
                s3 = synth.make_scene("c3")
                d3 = to_device(s3, cx.dev)
                r = measure_training(cx, s3, K, W, 5, e2e=False, dev_scene=d3)
                rec = {"config": bench_config(s3), "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "n_gpus": cx.world,
                       "scaling": "weak", "repeats": r["repeats"], "kernel_ms": r["roofline"]["kernel_ms"], "roofline": r["roofline"],
                       "parallelism": r["parallelism"]}
                if cx.world > 1 and cx.multi == "peer":
                    rec["selfcheck"] = multi_gpu_selfcheck(cx, s3, d3)
                return rec
            guarded("c3", c3)
        if cx.world == 1:
            def c2_tri():
                d2 = to_device(sc, cx.dev)
                r = measure_training(cx, sc, K, W, 5, e2e=False, dev_scene=d2, mode="trilinear")
                roof = r["roofline"]
                m_in, n = roof["step"]["m_in"], sc.n_rays
                ms_r = roof["kernel_ms"]["render_train"]
                alg = (128.0 + 32.0 * 8) * m_in + 96.0 * n      # 8 corner gathers + 8 gradient read-modify-writes per in-bounds sample
                return {"workload": workload_label(sc).replace("nearest lookup", "trilinear lookup"), "value": r["value"], "unit": UNIT,
                        "ms_per_step": r["ms_per_step"], "repeats": r["repeats"], "kernel_ms": roof["kernel_ms"],
                        "roofline": {"bound": "hbm", "kernel": "render_train (trilinear)", "achieved": alg / (ms_r * 1e-3) / 1e9,
                                     "peak": cx.peak, "unit": "GB/s", "frac": alg / (ms_r * 1e-3) / 1e9 / cx.peak,
                                     "algorithmic_bytes_per_launch": alg, "model": "(128 + 256)*M_in + 96*N"}}
            guarded("c2_trilinear", c2_tri)
            guarded("c4", lambda: extra_c4(cx))
            guarded("c5", lambda: extra_c5(cx))

    # ---- CPU baseline (rank 0, N = 1 only): the torch-CPU port of the reference's step on the host cores
    cpu_baseline = None
    if cx.world == 1 and not args.no_cpu_baseline:
        C_, R = sc.poses.shape[0], sc.rays_per_cam
        uvs = [synth.random_uv(C_, R, seed=1000 + i) for i in range(4)]
        r = time_reference_port(sc, uvs, steps=3, warmup=1, budget_s=30.0)
        cpu_baseline = {"value": r["rays_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                        "sample": f"3 steps of {r['rays_per_step']} rays after 1 warm-up ({r['ms_per_step']:.0f} ms/step), torch-CPU port "
                                  "of scripts/train.py:130-184 (oracle/torch_port.py)"}
        try:        # second CPU data point: the C / OpenMP restatement (never allowed to break the bench line)
            cpu_baseline["c_port"] = time_c_port(sc, uvs, steps=3, warmup=1)
        except Exception as e:
            log(f"[bench] C port not timed ({type(e).__name__}: {e})")

    if cx.rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": cx.world, "steps": K, "warmup": W,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(sc),
            "run": {"parallelism": head["parallelism"], "distinct_batches": min(K + W, MAX_DISTINCT_BATCHES),
                    "final_loss_global": head["final_loss"], "repeats": head["repeats"],
                    "timing": f"{REPEATS} regions of exactly {K} steps each (barrier + synchronize both sides, CUDA events, max over ranks); "
                              "value and ms_per_step are the median region"},
            "e2e": head["e2e"],
            "gpu_launches": head["launches_per_step"] * K,
            "clocks": clocks,
            "roofline": head["roofline"],
            "cpu_baseline": cpu_baseline,
            "extra": extra,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if cx.world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (skip the C3 / C4 / C5 / trilinear records)")
    ap.add_argument("--tune", action="append", default=[], metavar="NAME=INT", help="A/B runs: set a library tuning switch (plx_tune)")
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
