"""Fused training step driver: the loop body of `fit()` (scripts/train.py:104-191) as TWO kernel launches issued by ONE
C-ABI call per step (plx_train_step): fused march (ray generation + forward + MSE + backward) -> Adam.

Multi-GPU: rays are independent, so every rank holds a full grid replica and renders its own ray batch.
The loss / gradient scale uses the GLOBAL ray count, so a plain SUM of the ranks' gradients reproduces the single-GPU
mean-MSE gradient exactly (up to summation order).  Three exchanges exist:

  VoxelTrainer + process group          NCCL all-reduce of the dense gradient, replicated Adam          (baseline)
  PeerVoxelTrainer(exchange="pull")     one kernel: owner pulls its slab of every peer's gradient over NVLink (or lets the
                                        NVSwitch reduce it), Adam on the slab, parameters stored to every replica
  PeerVoxelTrainer(exchange="push")     the march itself reduces every touched cell straight into the slab owner's buffer
                                        (only the touched cells cross NVLink, overlapped with the march), then Adam on the
                                        slab with all-local loads and the parameters stored to every replica
Every rank reports the GLOBAL loss (the number the reference prints, scripts/train.py:156-159) on all three paths.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import time
import weakref

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops


_NO_CONTEXT = contextlib.nullcontext()


def shard_cameras(n_cams: int, rank: int, world: int) -> range:
    """Contiguous block of cameras owned by `rank` (ray r of camera c lives with the camera: the reference's rays are
    camera-major, src/ray_sampling.py:243-245).  Blocks differ by at most one camera."""
    base, extra = divmod(n_cams, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """SUM-all-reduce in place when a process group is up and has more than one rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def image_format(imgs: torch.Tensor) -> int:
    """PLX_IMG_F32 for fp32 (C,H,W,4) images in [0,1] (what the reference's loader returns), PLX_IMG_U8 for the PNGs' own
    uint8 RGBA (a quarter of the bytes; the kernels evaluate fp32(u8) / 255 when they fetch a target pixel)."""
    if imgs.dim() != 4 or imgs.shape[3] != 4:
        raise L.PlxError(f"imgs must be (C,H,W,4), got {tuple(imgs.shape)}")
    if imgs.dtype == torch.uint8:
        return L.PLX_IMG_U8
    if imgs.dtype == torch.float32:
        return L.PLX_IMG_F32
    raise L.PlxError(f"imgs must be float32 or uint8, got {imgs.dtype}")


class VoxelTrainer:
    """State + scratch of one grid replica and the single-call step.

    grid (X,Y,Z,4) fp32 raw values (clipped to [0,1] on lookup, scripts/train.py:146); Adam state and the running
    sum of |grad| (`grid_cells_full_grad`, :184) live next to it.  `imgs` (C,H,W,4; fp32, or uint8 as decoded) and `poses`
    (C,4,4) stay resident on the device as in the reference (:75); per step only the (C,R,2) uv draw changes.

    Loss: `step()` returns a (1,) VIEW of a two-slot device buffer — valid until the step after the next one starts (the
    optimiser kernel of step s+1 clears the slot of step s+2 = the slot of step s).  Read it (`float(...)`), `.clone()` it,
    or use `step_host` + `wait_result` if it has to outlive that.
    """

    def __init__(self, grid, points_distance, poses, fov, imgs, rays_per_cam, num_samples, delta_step, lr,
                 mode="nearest", beta=0.0, betas=(0.9, 0.999), eps=1e-8, n_rays_global=None, group=None, tv=0.0):
        dev = L.require_cuda(grid, poses, imgs)
        self.lib = L.load()
        self.device = dev
        self.group = group
        self.grid = grid.detach().to(torch.float32).contiguous().clone()
        self.grad = torch.zeros_like(self.grid)
        self.exp_avg = torch.zeros_like(self.grid)
        self.exp_avg_sq = torch.zeros_like(self.grid)
        self.grad_abs_sum = torch.zeros_like(self.grid)
        self.poses = poses.detach().to(torch.float32).contiguous()
        self.img_format = image_format(imgs)
        self.imgs = imgs.detach().contiguous()
        self.fov = float(fov)
        self.points_distance = float(points_distance)
        self.rays_per_cam, self.num_samples, self.delta_step = int(rays_per_cam), int(num_samples), float(delta_step)
        self.lr, self.betas, self.eps, self.mode, self.beta = float(lr), betas, float(eps), mode, float(beta)
        self.step_count = 0
        self.tv = float(tv)                     # weight of tv_loss (scripts/train.py:163-168); its value lands in self.tv_loss
        self.tv_loss = torch.zeros((1,), dtype=torch.float32, device=dev)
        self._tv_scratch = torch.zeros((1,), dtype=torch.float64, device=dev)
        self._dims = (C.c_int32 * 3)(*[int(s) for s in self.grid.shape[:3]])
        C_ = self.poses.shape[0]
        n = C_ * self.rays_per_cam
        self.n_rays = n
        self.n_rays_global = int(n_rays_global) if n_rays_global is not None else n
        self.gmin = ops.grid_origin(self.grid.shape, points_distance)
        # scratch
        self.uv = torch.empty((C_, self.rays_per_cam, 2), dtype=torch.float32, device=dev)
        self.dirs = torch.empty((n, 3), dtype=torch.float32, device=dev)
        self.targets = torch.empty((n, 4), dtype=torch.float32, device=dev)
        self.rgba = torch.empty((n, 4), dtype=torch.float32, device=dev)
        self.grad_rgba = torch.empty((n, 4), dtype=torch.float32, device=dev)
        self.tcarry = torch.empty((n, self.lib.plx_num_chunks(self.num_samples)), dtype=torch.float32, device=dev)
        # two loss slots: step s accumulates into slot s & 1, the optimiser kernel clears the other one (plenoxel_abi.h)
        self._loss2 = self._alloc_loss_slots()
        self.loss = self._loss2[1:2]                                   # view of the slot of the latest step
        # pinned { float loss; int32 step } the optimiser kernel publishes to on the host path
        self.result_host = torch.zeros((2,), dtype=torch.float32).pin_memory()
        self.loss_host = self.result_host[:1]
        self._result_f32 = self.result_host.numpy()
        self._result_i32 = self._result_f32.view("int32")
        self._result_ptr = self.result_host.data_ptr()
        self._dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self._uv_host_ok = {}
        self.unfused = False              # True: plx_generate_rays + K1 + K2 instead of the fused march (PLX_STEP_UNFUSED)
        self.wait_timeout_s = 60.0        # wait_result() gives up (raises) after this long without a published result
        self._args = self._make_args()
        self.launches_per_step = 2        # fused march, optimiser (memset / memcpy are not kernels)

    def _alloc_loss_slots(self) -> torch.Tensor:
        return torch.zeros((2,), dtype=torch.float32, device=self.device)

    def _make_args(self) -> L.PlxTrainStep:
        a = L.PlxTrainStep()
        a.march = L.make_march(self.grid, self.num_samples, self.delta_step, self.gmin, self.points_distance, self.mode, True)
        a.imgs, a.n_cams, a.img_h, a.img_w = self.imgs.data_ptr(), self.imgs.shape[0], self.imgs.shape[1], self.imgs.shape[2]
        a.img_format = self.img_format
        a.poses, a.fov = self.poses.data_ptr(), self.fov
        a.uv, a.rays_per_cam = self.uv.data_ptr(), self.rays_per_cam
        a.n_rays_global = self.n_rays_global
        a.grid, a.grad = self.grid.data_ptr(), self.grad.data_ptr()
        a.exp_avg, a.exp_avg_sq, a.grad_abs_sum = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.grad_abs_sum.data_ptr()
        a.lr, a.beta1, a.beta2, a.eps = self.lr, self.betas[0], self.betas[1], self.eps
        a.step = 0
        m_global = self.n_rays_global * self.num_samples
        a.beta_over_m = self.beta / m_global if (self.beta and m_global) else 0.0
        a.dirs, a.targets, a.rgba = self.dirs.data_ptr(), self.targets.data_ptr(), self.rgba.data_ptr()
        a.grad_rgba, a.tcarry, a.loss = self.grad_rgba.data_ptr(), self.tcarry.data_ptr(), self._loss2.data_ptr()
        return a

    # ---------------------------------------------------------------------------------------------------------
    def _distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _phase(self, phase: int) -> int:
        return phase | (L.PLX_STEP_UNFUSED if self.unfused else 0)

    def render_phase(self, uv: torch.Tensor | None = None) -> None:
        """Ray generation + forward + loss + backward of this rank's batch into the local gradient buffer."""
        self._args.uv = uv.data_ptr() if uv is not None else self.uv.data_ptr()
        self._begin_step()
        with torch.cuda.device(self.device):
            L.check(self.lib.plx_train_step(C.byref(self._args), self._phase(L.PLX_STEP_RENDER), L.stream_ptr(self.device)),
                    "plx_train_step(render)")
        self._args.uv = self.uv.data_ptr()

    def _add_tv(self, begin_cell: int | None = None, end_cell: int | None = None, atomic: bool = False) -> None:
        """tv * tv_loss gradient into the gradient buffer — of the whole grid, or of the slab of cells this rank's optimiser
        step consumes (`atomic`: peers may be reducing into the buffer meanwhile).  Every replica computes the same loss."""
        n_cells = self.grid.numel() // 4
        b, e = (0, n_cells) if begin_cell is None else (begin_cell, end_cell)
        L.check(self.lib.plx_tv_loss_range(self.grid.data_ptr(), self._dims, self.tv, self.grad.data_ptr(), b, e, int(atomic),
                                           self._tv_scratch.data_ptr(), self.tv_loss.data_ptr(), L.stream_ptr(self.device)),
                "plx_tv_loss")

    def _exchange(self) -> None:
        """NCCL baseline: SUM of the dense gradient and of this step's loss accumulator over the ranks."""
        if self._distributed():
            all_reduce_sum_(self.grad, self.group)
            all_reduce_sum_(self.loss, self.group)

    def update_phase(self) -> None:
        """[gradient exchange] + [TV gradient] + Adam (+ |grad| accumulation, gradient clear)."""
        with torch.cuda.device(self.device):
            self._exchange()
            if self.tv > 0:
                self._add_tv()
            L.check(self.lib.plx_train_step(C.byref(self._args), L.PLX_STEP_OPTIM, L.stream_ptr(self.device)),
                    "plx_train_step(optim)")

    def step(self, uv: torch.Tensor | None = None) -> torch.Tensor:
        """One step with the uv draw already on the device (`uv` (C,R,2) cuda, or the trainer's own `self.uv`).
        Returns the device loss (1,) without synchronising — the GLOBAL loss when a process group is up; see the class
        docstring for how long the returned view stays valid."""
        if self._distributed() or self.tv > 0:
            self.render_phase(uv)
            self.update_phase()
            return self.loss
        self._args.uv = uv.data_ptr() if uv is not None else self.uv.data_ptr()
        self._begin_step()
        with torch.cuda.device(self.device):
            L.check(self.lib.plx_train_step(C.byref(self._args), self._phase(L.PLX_STEP_ALL), L.stream_ptr(self.device)), "plx_train_step")
        self._args.uv = self.uv.data_ptr()
        return self.loss

    def _begin_step(self) -> None:
        self.step_count += 1
        self._args.step = self.step_count
        s = self.step_count & 1
        self.loss = self._loss2[s:s + 1]

    def check_errors(self) -> None:
        """Raise if a device-side wait of this trainer gave up (multi-GPU trainers override; nothing can fail here)."""

    def wait_result(self, step: int | None = None) -> float:
        """Host path: poll the pinned {loss, step} record until the optimiser kernel of `step` (default: the latest step)
        has published it, then return the loss.  No driver synchronisation call on the fast path; every few thousand polls
        it also checks the failure record of the cross-GPU waits, whether the stream ran dry without publishing (a kernel
        fault) and the deadline `wait_timeout_s` — each of those raises `PlxError` instead of spinning forever."""
        step = self.step_count if step is None else step
        seq = self._result_i32
        t0 = None
        polls = 0
        while seq[1] != step:
            polls += 1
            if polls & 0x3fff:
                continue
            now = time.perf_counter()
            if t0 is None:
                t0 = now
            self.check_errors()
            if torch.cuda.current_stream(self.device).query() and seq[1] != step:
                torch.cuda.synchronize(self.device)      # surfaces a sticky kernel fault as the CUDA error it is
                if seq[1] != step:
                    raise L.PlxError(f"wait_result: the stream is idle but step {step} was never published (last published: {int(seq[1])})")
            if now - t0 > self.wait_timeout_s:
                raise L.PlxError(f"wait_result: step {step} not published after {self.wait_timeout_s:.0f} s (last published: {int(seq[1])})")
        return float(self._result_f32[0])

    def _check_uv_host(self, uv_host: torch.Tensor) -> None:
        """Pinned, contiguous fp32 (C,R,2) on the host.  `is_pinned()` is a driver query, so a tensor object that passed once
        is remembered (by identity, through a weak reference) and not asked again on every step."""
        ref = self._uv_host_ok.get(id(uv_host))
        if ref is not None and ref() is uv_host:
            return
        if uv_host.is_cuda or uv_host.dtype != torch.float32 or not uv_host.is_contiguous() or \
                uv_host.numel() != self.uv.numel() or not uv_host.is_pinned():
            raise L.PlxError("uv_host must be a pinned, contiguous float32 host tensor of shape (C,R,2)")
        if len(self._uv_host_ok) > 64:
            self._uv_host_ok.clear()
        self._uv_host_ok[id(uv_host)] = weakref.ref(uv_host)

    def step_host(self, uv_host: torch.Tensor) -> torch.Tensor:
        """End-to-end step from HOST memory: the march reads this step's uv straight out of pinned `uv_host` (C,R,2)
        (zero-copy) and the optimiser kernel publishes {loss, step} into pinned `self.result_host`; two kernel launches,
        asynchronous.  Read the loss with `wait_result()` (or synchronise the stream and read `self.loss_host`).  Same
        objective as `step()`: the exchange and the TV term run between the two halves when they apply."""
        self._check_uv_host(uv_host)
        self._begin_step()
        st = L.stream_ptr(self.device)
        res = self._result_ptr
        # the host is on the critical path here (it may only issue step t+1 once it has read the loss of step t): no device
        # context switch when this trainer's device is already current
        with (_NO_CONTEXT if torch.cuda.current_device() == self._dev_index else torch.cuda.device(self.device)):
            if self._distributed() or self.tv > 0:
                L.check(self.lib.plx_train_step_host(C.byref(self._args), uv_host.data_ptr(), res, self._phase(L.PLX_STEP_RENDER), st),
                        "plx_train_step_host(render)")
                self._exchange()
                if self.tv > 0:
                    self._add_tv()
                L.check(self.lib.plx_train_step_host(C.byref(self._args), None, res, L.PLX_STEP_OPTIM, st),
                        "plx_train_step_host(optim)")
            else:
                L.check(self.lib.plx_train_step_host(C.byref(self._args), uv_host.data_ptr(), res, self._phase(L.PLX_STEP_ALL), st),
                        "plx_train_step_host")
        return self.loss_host

    # ------------------------------------------------------------------------------------------------ CUDA-graph replay
    GRAPH_TABLE_STEPS = 1 << 16

    def _refill_replay_table(self) -> None:
        """Adam's two step-dependent scalars for the next GRAPH_TABLE_STEPS steps, formed on the host in double like
        plx_adam_step forms them (plx_adam_table), uploaded outside the graph."""
        n = self.GRAPH_TABLE_STEPS
        host = torch.empty((n, 2), dtype=torch.float32)
        L.check(self.lib.plx_adam_table(self.lr, self.betas[0], self.betas[1], self.step_count + 1, n,
                                        C.cast(host.data_ptr(), C.POINTER(C.c_float))), "plx_adam_table")
        self._replay_table.copy_(host)
        self._replay.table_base = self.step_count

    def capture_graph(self, host_uv: bool = False) -> None:
        """Capture the whole step (fused march + optimiser, both launched programmatically dependent) into ONE CUDA graph that
        `step_graph()` replays: the step number, the loss slot and Adam's bias-corrected scalars come from device memory
        (PlxReplayState), so nothing in the graph changes from step to step.  The march reads the uv draw from `self.uv`
        (device; `host_uv=False`) or zero-copy from the pinned host buffer `self.uv_host` (`host_uv=True`, the optimiser kernel
        then also publishes {loss, step} to `self.result_host` as `step_host` does).  Single GPU, tv = 0."""
        if self._distributed() or self.tv > 0:
            raise L.PlxError("capture_graph: single-GPU trainers with tv = 0 only (the exchange / TV kernels are not part of the graph)")
        dev = self.device
        self._step_dev = torch.tensor([self.step_count], dtype=torch.int32, device=dev)
        self._replay_counter = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._replay_table = torch.empty((self.GRAPH_TABLE_STEPS, 2), dtype=torch.float32, device=dev)
        rp = L.PlxReplayState()
        rp.step_dev, rp.table, rp.block_counter = self._step_dev.data_ptr(), self._replay_table.data_ptr(), self._replay_counter.data_ptr()
        rp.table_len = self.GRAPH_TABLE_STEPS
        self._replay = rp
        self._refill_replay_table()
        self._args.replay = C.pointer(rp)
        if host_uv:
            self.uv_host = torch.empty(tuple(self.uv.shape), dtype=torch.float32).pin_memory()
        self._graph_host = bool(host_uv)
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(graph):
            st = L.stream_ptr(dev)
            if host_uv:
                L.check(self.lib.plx_train_step_host(C.byref(self._args), self.uv_host.data_ptr(), self._result_ptr, self._phase(L.PLX_STEP_ALL), st),
                        "plx_train_step_host(capture)")
            else:
                self._args.uv = self.uv.data_ptr()
                L.check(self.lib.plx_train_step(C.byref(self._args), self._phase(L.PLX_STEP_ALL), st), "plx_train_step(capture)")
        self._graph = graph

    def step_graph(self):
        """Replay the captured step on the current contents of `self.uv` / `self.uv_host`.  Returns the device loss view
        (`host_uv=False`) or the pinned `loss_host` (`host_uv=True`: read it with `wait_result()`)."""
        if self.step_count - self._replay.table_base >= self.GRAPH_TABLE_STEPS:
            torch.cuda.current_stream(self.device).synchronize()
            self._refill_replay_table()
        self.step_count += 1
        s = self.step_count & 1
        self.loss = self._loss2[s:s + 1]
        self._graph.replay()
        return self.loss_host if self._graph_host else self.loss

    def flush(self) -> None:
        """Everything a step promised is complete once the stream is (single GPU / NCCL exchange): nothing to do here.
        PeerVoxelTrainer overrides it."""

    def full_grad_abs_sum(self) -> torch.Tensor:
        """`grid_cells_full_grad` of scripts/train.py:184 (PeerVoxelTrainer gathers it from the slab owners)."""
        return self.grad_abs_sum

    def _full_moments(self):
        return self.exp_avg, self.exp_avg_sq

    def checkpoint(self, extra_param: dict | None = None) -> dict:
        """The reference's `.pth` payload (scripts/train.py:194-210): {grid, grid_grad, param{...}} on the CPU."""
        self.flush()
        self.check_errors()
        param = {"device": str(self.device), "number_of_rays": self.rays_per_cam, "num_samples": self.num_samples,
                 "delta_step": self.delta_step, "even_spread": False, "camera_ray": False,
                 "points_distance": self.points_distance, "gridsize": list(self.grid.shape[:3])}
        if extra_param:
            param.update(extra_param)
        return {"grid": self.grid.detach().cpu(), "grid_grad": self.full_grad_abs_sum().detach().cpu(), "param": param}

    def resume_state(self) -> dict:
        """What the reference's checkpoint lacks for resuming (it saves no optimiser state): Adam moments and
        the step counter, as FULL arrays whatever the trainer's sharding (so a run can resume on another world size).
        `checkpoint()` merged with this dict is a superset of the reference format; scripts that read only `grid` /
        `grid_grad` / `param` (scripts/compare_inference_to_image.py:41-49) are unaffected."""
        self.flush()
        self.check_errors()
        m, v = self._full_moments()
        return {"optimizer": {"exp_avg": m.detach().cpu(), "exp_avg_sq": v.detach().cpu(),
                              "step": self.step_count, "lr": self.lr, "betas": tuple(self.betas), "eps": self.eps}}

    def load_state(self, ckpt: dict) -> None:
        """Restore grid, |grad| sum and (if present) the optimiser state saved by `checkpoint()` / `resume_state()`."""
        with torch.no_grad():
            self.grid.copy_(ckpt["grid"].to(self.device))
            if "grid_grad" in ckpt:
                self.grad_abs_sum.copy_(ckpt["grid_grad"].to(self.device))
            opt = ckpt.get("optimizer")
            if opt:
                self.exp_avg.copy_(opt["exp_avg"].to(self.device))
                self.exp_avg_sq.copy_(opt["exp_avg_sq"].to(self.device))
                self.step_count = int(opt["step"])
            self.grad.zero_()
            self._loss2.zero_()


def slab_range(n_cells: int, rank: int, world: int):
    """[begin, end) in floats of the contiguous block of cells whose optimiser state `rank` owns (pull exchange)."""
    base, extra = divmod(n_cells, world)
    start = rank * base + min(rank, extra)
    return 4 * start, 4 * (start + base + (1 if rank < extra else 0))


def slab_partition(n_cells: int, rank: int, world: int):
    """(owner_mul, begin_cell, end_cell) of the push exchange — python mirror of plx_slab_partition (include/plenoxel_abi.h):
    the march sends the gradient of cell `lin` to rank (lin * owner_mul) >> 32, so rank r owns the preimage of r."""
    if not (1 <= world <= L.PLX_MAX_PEERS and 0 <= rank < world and world < n_cells < 2 ** 31):
        raise L.PlxError(f"slab partition: need 0 <= rank < world <= {L.PLX_MAX_PEERS} and world < n_cells < 2^31")
    mul = (world << 32) // n_cells
    first = lambda r: min(n_cells, -((-(r << 32)) // mul))
    return mul, first(rank), n_cells if rank + 1 == world else first(rank + 1)


def pick_exchange(n_rays: int, num_samples: int, n_cells: int) -> str:
    """"push" unless a rank's batch is so dense that its reductions would outweigh the dense exchange many times over.
    Measured on B200 (bench.py, 20-step regions): sparse batches (C3: 0.56 M in-bounds samples into 16.8 M cells) 278 vs 501 us per
    step at N = 2 and 426 vs 624 us at N = 8; dense batches (C2: 3.3 M samples into 2.1 M cells) still 97 vs 127 us at N = 2 and
    131 vs 147 us at N = 8, because the reductions travel while the march computes.  About half of a ray's samples lie inside
    the grid; beyond ~8 in-bounds samples per cell the dense pull exchange moves fewer bytes."""
    return "push" if n_rays * num_samples // 2 < 8 * n_cells else "pull"


class PeerVoxelTrainer(VoxelTrainer):
    """Multi-GPU step whose gradient exchange runs over NVLink peer memory inside our own kernels — no NCCL on the data path.

    Every rank keeps a full grid replica and full-size gradient buffer(s) in symmetric memory
    (torch.distributed._symmetric_memory), so each process holds a mapped pointer to every peer's copy, and owns the
    optimiser state of one contiguous slab of cells.  Per step

      exchange="push"   K12   render this rank's rays; every gradient reduction goes to the OWNER of its cell
                              (`red.global.add.v4.f32` on a peer-mapped address)            NVLink, overlapped with the march
                        --    barrier: every rank's reductions have landed
                        K3s   Adam over the owned slab (local loads), parameters stored to every replica (`multimem.st` /
                              per-peer stores), gradient slab cleared in place                NVLink stores only
                        --    barrier: every replica holds the new parameters
      exchange="pull"   K12   render into the LOCAL gradient buffer                           no communication
                        --    barrier
                        K3p   owned slab: sum the partial gradients out of the peers' buffers (or `multimem.ld_reduce`),
                              Adam, store the parameters into every replica                   NVLink loads + stores
                        --    barrier; the consumed buffer is cleared on a side stream (two buffers alternate)

    Failures are loud: every device-side wait is bounded (`peer_timeout_s`); one that gives up is recorded in a device word
    and a pinned host word, later kernels of the trainer then store nothing, and `flush()`, `wait_result()`, `checkpoint()`
    and every `check_every`-th `step()` raise `PlxError`.
    """

    def __init__(self, grid, *args, group=None, exchange: str | None = None, multicast: bool | None = None,
                 peer_timeout_s: float = 30.0, check_every: int = 64, fused_sync: bool = False, **kwargs):
        import torch.distributed._symmetric_memory as symm_mem
        if not (dist.is_available() and dist.is_initialized()):
            raise L.PlxError("PeerVoxelTrainer needs an initialised NCCL process group")
        group = group or dist.group.WORLD
        self._symm, self._group = symm_mem, group
        super().__init__(grid, *args, group=group, **kwargs)
        dev = self.device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > L.PLX_MAX_PEERS:
            raise L.PlxError(f"at most {L.PLX_MAX_PEERS} peers")
        n_cells = self.grid.numel() // 4
        self.exchange = exchange or pick_exchange(self.n_rays, self.num_samples, n_cells)
        if self.exchange not in ("push", "pull"):
            raise L.PlxError(f"unknown exchange {self.exchange!r}")
        shape = tuple(self.grid.shape)
        sym_grid = symm_mem.empty(shape, dtype=torch.float32, device=dev)
        sym_grid.copy_(self.grid)
        self._h_grid = symm_mem.rendezvous(sym_grid, group)
        self.grid = sym_grid
        dist.broadcast(self.grid, src=dist.get_global_rank(group, 0), group=group)      # identical replicas to start from
        # gradient buffers: one (push: the owner clears its slab in place) or two used on alternate steps (pull: the one
        # consumed by step i is cleared on a side stream while step i+1 renders into the other)
        self._grads, self._h_grads = [], []
        for _ in range(1 if self.exchange == "push" else 2):
            gbuf = symm_mem.empty(shape, dtype=torch.float32, device=dev)
            gbuf.zero_()
            self._grads.append(gbuf)
            self._h_grads.append(symm_mem.rendezvous(gbuf, group))
        self.grad = self._grads[0]
        self._clear_stream = torch.cuda.Stream(device=dev)
        self._cleared = [None, None]          # event: buffer b is zero again
        self._dirty = None                    # buffer that holds a consumed gradient and still has to be cleared
        self._done_ev = [torch.cuda.Event(), torch.cuda.Event()]
        self._clear_ev = [torch.cuda.Event(), torch.cuda.Event()]
        # barrier flags: int32[channels * max_peers] per rank in symmetric memory, epochs only ever grow
        flags = symm_mem.empty((4 * L.PLX_MAX_PEERS,), dtype=torch.int32, device=dev)
        flags.zero_()
        self._flags = flags
        self._h_flags = symm_mem.rendezvous(flags, group)
        self._flag_ptrs = (L.c_void * L.PLX_MAX_PEERS)(*[int(self._h_flags.buffer_ptrs[r]) if r < self.world else None
                                                           for r in range(L.PLX_MAX_PEERS)])
        self._loss_peer_ptrs = [int(self._h_loss.buffer_ptrs[r]) for r in range(self.world)]
        self._loss_global = torch.zeros((2,), dtype=torch.float32, device=dev)
        # failure record of the device-side waits (PlxPeerError)
        self._err_dev = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._err_host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self._err_np = self._err_host.numpy()
        self._err = L.PlxPeerError()
        self._err.device_word, self._err.host_word = self._err_dev.data_ptr(), self._err_host.data_ptr()
        self._err.timeout_ns = int(peer_timeout_s * 1e9)
        self.check_every = int(check_every)
        self._epoch = 0
        # fused ordering (PlxPeerSync, pull exchange only, opt-in): the march signals "partial gradient complete"
        # (channel 0) and waits for "all slabs of the previous step stored" (channel 1); the exchange kernel waits on 0 and
        # signals 1.  No stand-alone barrier launches on the step; `flush()` (channel 2) is the full barrier before anyone
        # reads a replica from outside.  Measured neutral (N=2: 123.7 vs 126.1 us/step, N=4: 150.3 vs 146.0): the two
        # barrier launches were already hidden behind the kernels they order, what remains is waiting for the slowest
        # rank — so the default stays the simpler contract (a finished stream = a complete replica).
        self._fused = bool(fused_sync)
        if self._fused and self.exchange != "pull":
            raise L.PlxError("fused_sync is implemented for the pull exchange only")
        if self._fused and self.tv > 0:
            raise L.PlxError("fused_sync leaves no place for the TV term between march and optimiser; use fused_sync=False")
        self._sync_counters = torch.zeros((2,), dtype=torch.int32, device=dev)
        self._sync_render, self._sync_adam = L.PlxPeerSync(), L.PlxPeerSync()
        for k, sy in enumerate((self._sync_render, self._sync_adam)):
            for r in range(self.world):
                sy.flags[r] = int(self._h_flags.buffer_ptrs[r])
            sy.rank, sy.world = self.rank, self.world
            sy.block_counter = self._sync_counters.data_ptr() + 4 * k
            sy.err = self._err
        self._flush_epoch = 0

        # NVLS multicast mappings, when the fabric offers them (in-switch reduce / replicate).  Default: pull exchange from 4
        # ranks up (in-switch reduction of the gathered gradients; at N=2 the switch round trip costs more than it saves: 93 vs
        # 60 us); push exchange never — a multicast store also loops back into the sender's own replica, so every GPU receives
        # the WHOLE grid (16 B x cells) instead of the other ranks' slabs; measured per-peer vs multicast stores, us per step:
        # C2 110 vs 116 (N=4), 115.5 vs 116.5 (N=8); C3 379 vs 433 (N=4), 418 vs 427 (N=8)
        def mc_ptr(h):
            try:
                return int(h.multicast_ptr or 0)
            except Exception:       # no multicast support on this fabric
                return 0
        mc_grid, mc_grads = mc_ptr(self._h_grid), [mc_ptr(h) for h in self._h_grads]
        have_mc = bool(mc_grid) and (self.exchange == "push" or all(mc_grads))
        self.multicast = have_mc and ((self.exchange == "pull" and self.world >= 4) if multicast is None else bool(multicast))
        if multicast and not have_mc:
            raise L.PlxError("multicast requested but the symmetric-memory handles expose no multicast mapping")
        if self.exchange == "push":
            mul, b_cell, e_cell = slab_partition(n_cells, self.rank, self.world)
            self._slab_cells = (b_cell, e_cell)
            pg = L.PlxPeerGrad()
            for r in range(self.world):
                pg.grads[r] = int(self._h_grads[0].buffer_ptrs[r])
            pg.owner_mul, pg.world = mul, self.world
            self._peer_grad = pg
            s = L.PlxAdamSlab()
            s.world, s.rank = self.world, self.rank
            for r in range(self.world):
                s.grids[r] = int(self._h_grid.buffer_ptrs[r])
            s.grid_mc = mc_grid if self.multicast else None
            s.grad = self.grad.data_ptr()
            s.exp_avg, s.exp_avg_sq, s.grad_abs_sum = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.grad_abs_sum.data_ptr()
            s.begin, s.end = 4 * b_cell, 4 * e_cell
            s.lr, s.beta1, s.beta2, s.eps = self.lr, self.betas[0], self.betas[1], self.eps
            s.err = self._err
            self._slab = s
        else:
            b, e = slab_range(n_cells, self.rank, self.world)
            self._slab_cells = (b // 4, e // 4)
            self._peers = []
            for k in range(2):
                p = L.PlxAdamPeer()
                p.world, p.rank = self.world, self.rank
                for r in range(self.world):
                    p.grids[r] = int(self._h_grid.buffer_ptrs[r])
                    p.grads[r] = int(self._h_grads[k].buffer_ptrs[r])
                p.exp_avg, p.exp_avg_sq, p.grad_abs_sum = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.grad_abs_sum.data_ptr()
                p.begin, p.end = b, e
                p.lr, p.beta1, p.beta2, p.eps = self.lr, self.betas[0], self.betas[1], self.eps
                if self.multicast:
                    p.grid_mc, p.grad_mc = mc_grid, mc_grads[k]
                p.sync.err = self._err
                self._peers.append(p)
        self._cur = 0
        self._no_barriers = False             # profiling only (bench.py times the kernels of a throw-away trainer in isolation)
        self._args = self._make_args()
        self.launches_per_step = 2 if self._fused else 4          # march, [barrier], exchange / optimiser, [barrier]
        torch.cuda.synchronize(dev)
        dist.barrier(group)

    # the loss accumulators live in symmetric memory: the optimiser kernel of every rank sums all of them (global loss)
    def _alloc_loss_slots(self) -> torch.Tensor:
        t = self._symm.empty((2,), dtype=torch.float32, device=self.device)
        t.zero_()
        self._h_loss = self._symm.rendezvous(t, self._group)
        return t

    def _make_args(self) -> L.PlxTrainStep:
        a = super()._make_args()
        if getattr(self, "exchange", None) == "push" and hasattr(self, "_peer_grad"):
            a.peer_grad = C.pointer(self._peer_grad)
        return a

    def check_errors(self) -> None:
        code = int(self._err_np[0])
        if code:
            raise L.PlxError(f"cross-GPU wait timed out on rank {self.rank} (channel {(code & 0xff) - 1}, epoch {code >> 8}): a peer "
                             "died or stalled beyond peer_timeout_s; parameters were not updated after the failure "
                             f"(code {L.PLX_E_PEER_TIMEOUT})")

    def _begin_step(self) -> None:
        super()._begin_step()
        s = self.step_count & 1
        self.loss = self._loss_global[s:s + 1]          # the global loss the optimiser kernel writes (same bits on every rank)
        if self.check_every and self.step_count % self.check_every == 0:
            self.check_errors()

    def _select_buffer(self) -> None:
        if self.exchange == "push":
            return
        b = self.step_count % 2                 # buffer of the step about to run
        self._cur = b
        self.grad = self._grads[b]
        self._args.grad = self.grad.data_ptr()
        if self._cleared[b] is not None:        # its clear (issued two steps ago on the side stream) must have finished
            torch.cuda.current_stream(self.device).wait_event(self._cleared[b])
        if self._fused:
            e = self._epoch + 1                 # this step's epoch (the exchange below advances self._epoch to it)
            sy = self._sync_render
            sy.wait_channel, sy.wait_epoch = 1, e - 1
            sy.signal_channel, sy.signal_epoch = 0, e
            self._args.render_sync = C.pointer(sy)

    def _tv_own_slab(self) -> None:
        """TV term of the cells this rank owns, added right behind the march: the replica is stable until every rank has
        entered the barrier that follows (afterwards the peers' optimiser kernels store new parameters into it, and the
        stencil must not read a mix of old and new values); the owner's partial sum carries the term once."""
        if self.tv > 0:
            self._add_tv(*self._slab_cells, atomic=True)

    def render_phase(self, uv=None) -> None:
        self._select_buffer()
        super().render_phase(uv)
        with torch.cuda.device(self.device):
            self._tv_own_slab()

    def _loss_tail(self, a, result_host) -> None:
        s = self.step_count & 1
        for r in range(self.world):
            a.loss_peers[r] = self._loss_peer_ptrs[r] + 4 * s
        a.loss_out = self._loss_global.data_ptr() + 4 * s
        a.loss_clear = self._loss2.data_ptr() + 4 * (1 - s)
        a.result_host = result_host
        a.step = self.step_count

    def _exchange_and_update(self, st, result_host=None):
        self._epoch += 1
        if self.exchange == "push":
            self._barrier(0, st)                              # every rank's reductions into my slab have landed
            self._loss_tail(self._slab, result_host)
            L.check(self.lib.plx_adam_step_slab(C.byref(self._slab), st), "plx_adam_step_slab")
            self._barrier(1, st)                              # every replica holds the new parameters
            return
        b = self._cur
        peer = self._peers[b]
        self._loss_tail(peer, result_host)
        if self._fused:
            sy = peer.sync
            sy.flags, sy.rank, sy.world, sy.block_counter = self._sync_adam.flags, self.rank, self.world, self._sync_adam.block_counter
            sy.wait_channel, sy.wait_epoch = 0, self._epoch
            sy.signal_channel, sy.signal_epoch = 1, self._epoch
        else:
            self._barrier(0, st)                              # every rank's partial gradient is complete
        # the buffer consumed by the PREVIOUS step (its readers passed that step's closing barrier / signalled channel 1,
        # which this step's march waited for) is cleared now, on the side stream, behind this step's exchange kernel: that
        # kernel is NVLink-bound and leaves HBM idle, whereas clearing during the march (measured) slowed the march by as
        # much as the clear itself takes
        if self._dirty is not None:
            o = self._dirty
            done, ev = self._done_ev[o], self._clear_ev[o]
            done.record(torch.cuda.current_stream(self.device))
            self._clear_stream.wait_event(done)
            with torch.cuda.stream(self._clear_stream):
                self._grads[o].zero_()
                ev.record(self._clear_stream)
            self._cleared[o] = ev
        L.check(self.lib.plx_adam_step_peer(C.byref(peer), st), "plx_adam_step_peer")
        if not self._fused:
            self._barrier(1, st)                              # every replica holds the new parameters; peers done reading
        self._dirty = b

    def flush(self) -> None:
        """Full cross-rank barrier on the stream when the step itself does not end in one (fused ordering): after it, this
        rank's replica holds every peer's slab of the latest step.  Always checks the failure record."""
        if self._fused:
            self._flush_epoch += 1
            with torch.cuda.device(self.device):
                L.check(self.lib.plx_peer_barrier(self._flag_ptrs, self.rank, self.world, 2, self._flush_epoch, C.byref(self._err),
                                                  L.stream_ptr(self.device)), "plx_peer_barrier(flush)")
        self.check_errors()

    def _barrier(self, channel, st):
        if self._no_barriers:
            return
        L.check(self.lib.plx_peer_barrier(self._flag_ptrs, self.rank, self.world, channel, self._epoch, C.byref(self._err), st),
                "plx_peer_barrier")

    def update_phase(self) -> None:
        with torch.cuda.device(self.device):
            self._exchange_and_update(L.stream_ptr(self.device))

    def step(self, uv=None):
        self.render_phase(uv)
        self.update_phase()
        return self.loss

    def step_host(self, uv_host):
        self._check_uv_host(uv_host)
        self._select_buffer()
        self._begin_step()
        st = L.stream_ptr(self.device)
        with torch.cuda.device(self.device):
            L.check(self.lib.plx_train_step_host(C.byref(self._args), uv_host.data_ptr(), None, L.PLX_STEP_RENDER, st),
                    "plx_train_step_host(render)")
            self._tv_own_slab()
            self._exchange_and_update(st, self.result_host.data_ptr())
        return self.loss_host

    def _gather_slabs(self, t: torch.Tensor) -> torch.Tensor:
        """Full array out of the per-rank slabs of a slab-sharded state array (every rank gets the whole thing)."""
        b, e = self._slab_cells
        full = torch.zeros_like(t)
        full.view(-1, 4)[b:e] = t.view(-1, 4)[b:e]
        all_reduce_sum_(full, self.group)
        return full

    def full_grad_abs_sum(self):
        """Full `grid_grad` (scripts/train.py:184): each rank accumulated |grad| for the cells it owns only."""
        return self._gather_slabs(self.grad_abs_sum)

    gathered_grad_abs_sum = full_grad_abs_sum

    def _full_moments(self):
        return self._gather_slabs(self.exp_avg), self._gather_slabs(self.exp_avg_sq)

    def load_state(self, ckpt: dict) -> None:
        super().load_state(ckpt)
        with torch.no_grad():
            for g in self._grads:
                g.zero_()
            self._loss_global.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
