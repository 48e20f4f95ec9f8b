"""Fused training step driver: the loop body of `fit()` (scripts/train.py:104-191, tv = 0) as four kernel launches
issued by ONE C-ABI call per step (plx_train_step): ray generation -> K1 forward + MSE -> K2 backward -> K3 Adam.

Multi-GPU: rays are independent, so every rank holds a full grid replica, renders its own ray batch,
and the dense gradient (X,Y,Z,4) is SUM-all-reduced over NCCL between K2 and K3.  The loss/gradient scale uses the
GLOBAL ray count so a plain SUM reproduces the single-GPU mean-MSE gradient exactly (up to summation order).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
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


class VoxelTrainer:
    """State + scratch of one grid replica and the single-call step.

    grid (X,Y,Z,4) fp32 raw values (clipped to [0,1] on lookup, scripts/train.py:146); Adam state and the running
    sum of |grad| (`grid_cells_full_grad`, :184) live next to it.  `imgs` (C,H,W,4) and `poses` (C,4,4) stay resident
    on the device as in the reference (:75); per step only the (C,R,2) uv draw changes.
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
        self.imgs = imgs.detach().to(torch.float32).contiguous()
        self.fov = float(fov)
        self.points_distance = float(points_distance)
        self.rays_per_cam, self.num_samples, self.delta_step = int(rays_per_cam), int(num_samples), float(delta_step)
        self.lr, self.betas, self.eps, self.mode, self.beta = float(lr), betas, float(eps), mode, float(beta)
        self.step_count = 0
        self.tv = float(tv)                     # weight of tv_loss (scripts/train.py:163-168); its value lands in self.tv_loss
        self.tv_loss = torch.zeros((1,), dtype=torch.float32, device=dev)
        self._tv_scratch = torch.zeros((1,), dtype=torch.float64, device=dev)
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
        self._loss2 = torch.zeros((2,), dtype=torch.float32, device=dev)
        # work counter of the fused march (rays are claimed dynamically by the warps of a one-wave grid); the optimiser
        # kernel resets it every step.  Opt-in (PLX_TRAIN_DYNAMIC=1): the hardware block scheduler balanced C2 better.
        self._work_counter = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._dynamic = os.environ.get("PLX_TRAIN_DYNAMIC", "0") == "1"      # measured slower on C2 (61 vs 48 us, even with the ticket drawn one ray ahead): off
        self.loss = self._loss2[1:2]                                   # view of the slot of the latest step
        # pinned { float loss; int32 step } the optimiser kernel publishes to on the host path
        self.result_host = torch.zeros((2,), dtype=torch.float32).pin_memory()
        self.loss_host = self.result_host[:1]
        self._result_f32 = self.result_host.numpy()
        self._result_i32 = self._result_f32.view("int32")
        self._result_ptr = self.result_host.data_ptr()
        self._dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self._uv_host_ok = {}
        self._args = self._make_args()
        self.launches_per_step = 4        # generate_rays, render_fwd, render_bwd, adam (memset/memcpy are not kernels)

    def _make_args(self) -> L.PlxTrainStep:
        a = L.PlxTrainStep()
        a.march = L.make_march(self.grid, self.num_samples, self.delta_step, self.gmin, self.points_distance, self.mode, True)
        a.imgs, a.n_cams, a.img_h, a.img_w = self.imgs.data_ptr(), self.imgs.shape[0], self.imgs.shape[1], self.imgs.shape[2]
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
        a.work_counter = self._work_counter.data_ptr() if self._dynamic else None
        return a

    # ---------------------------------------------------------------------------------------------------------
    def _distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def render_phase(self, uv: torch.Tensor | None = None) -> None:
        """Ray generation + forward + loss + backward of this rank's batch into the local gradient buffer."""
        self._args.uv = uv.data_ptr() if uv is not None else self.uv.data_ptr()
        self._begin_step()
        with torch.cuda.device(self.device):
            L.check(self.lib.plx_train_step(C.byref(self._args), L.PLX_STEP_RENDER, L.stream_ptr(self.device)),
                    "plx_train_step(render)")
        self._args.uv = self.uv.data_ptr()

    def _add_tv(self) -> None:
        """tv * tv_loss gradient into the (already exchanged) gradient buffer; every replica computes the same term."""
        dims = (C.c_int32 * 3)(*[int(s) for s in self.grid.shape[:3]])
        L.check(self.lib.plx_tv_loss(self.grid.data_ptr(), dims, self.tv, self.grad.data_ptr(), self._tv_scratch.data_ptr(),
                                     self.tv_loss.data_ptr(), L.stream_ptr(self.device)), "plx_tv_loss")

    def update_phase(self) -> None:
        """[gradient exchange] + [TV gradient] + Adam (+ |grad| accumulation, gradient clear)."""
        with torch.cuda.device(self.device):
            if self._distributed():
                all_reduce_sum_(self.grad, self.group)
            if self.tv > 0:
                self._add_tv()
            L.check(self.lib.plx_train_step(C.byref(self._args), L.PLX_STEP_OPTIM, L.stream_ptr(self.device)),
                    "plx_train_step(optim)")

    def step(self, uv: torch.Tensor | None = None) -> torch.Tensor:
        """One step with the uv draw already on the device (`uv` (C,R,2) cuda, or the trainer's own `self.uv`).
        Returns the device loss tensor (1,) without synchronising.  With a process group: this rank's partial loss."""
        if self._distributed() or self.tv > 0:
            self.render_phase(uv)
            self.update_phase()
            return self.loss
        self._args.uv = uv.data_ptr() if uv is not None else self.uv.data_ptr()
        self._begin_step()
        with torch.cuda.device(self.device):
            L.check(self.lib.plx_train_step(C.byref(self._args), L.PLX_STEP_ALL, L.stream_ptr(self.device)), "plx_train_step")
        self._args.uv = self.uv.data_ptr()
        return self.loss

    def _begin_step(self) -> None:
        self.step_count += 1
        self._args.step = self.step_count
        s = self.step_count & 1
        self.loss = self._loss2[s:s + 1]

    def wait_result(self, step: int | None = None) -> float:
        """Host path: spin on the pinned {loss, step} record until the optimiser kernel of `step` (default: the latest
        step) has published it, then return the loss.  No driver synchronisation call is involved."""
        step = self.step_count if step is None else step
        seq = self._result_i32
        while seq[1] != step:
            pass
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
        asynchronous.  Read the loss with `wait_result()` (or synchronise the stream and read `self.loss_host`)."""
        self._check_uv_host(uv_host)
        self._begin_step()
        st = L.stream_ptr(self.device)
        res = self._result_ptr
        # the host is on the critical path here (it may only issue step t+1 once it has read the loss of step t): no device
        # context switch when this trainer's device is already current
        with (_NO_CONTEXT if torch.cuda.current_device() == self._dev_index else torch.cuda.device(self.device)):
            if self._distributed():
                L.check(self.lib.plx_train_step_host(C.byref(self._args), uv_host.data_ptr(), res, L.PLX_STEP_RENDER, st),
                        "plx_train_step_host(render)")
                all_reduce_sum_(self.grad, self.group)
                L.check(self.lib.plx_train_step_host(C.byref(self._args), None, res, L.PLX_STEP_OPTIM, st),
                        "plx_train_step_host(optim)")
            else:
                L.check(self.lib.plx_train_step_host(C.byref(self._args), uv_host.data_ptr(), res, L.PLX_STEP_ALL, st),
                        "plx_train_step_host")
        return self.loss_host

    def flush(self) -> None:
        """Everything a step promised is complete once the stream is (single GPU / NCCL exchange): nothing to do here.
        PeerVoxelTrainer overrides it."""

    def checkpoint(self, extra_param: dict | None = None) -> dict:
        """The reference's `.pth` payload (scripts/train.py:194-210): {grid, grid_grad, param{...}} on the CPU."""
        param = {"device": str(self.device), "number_of_rays": self.rays_per_cam, "num_samples": self.num_samples,
                 "delta_step": self.delta_step, "even_spread": False, "camera_ray": False,
                 "points_distance": self.points_distance, "gridsize": list(self.grid.shape[:3])}
        if extra_param:
            param.update(extra_param)
        return {"grid": self.grid.detach().cpu(), "grid_grad": self.grad_abs_sum.detach().cpu(), "param": param}

    def resume_state(self) -> dict:
        """What the reference's checkpoint lacks for resuming (it saves no optimiser state): Adam moments and
        the step counter.  `checkpoint()` merged with this dict is a superset of the reference format; scripts that read
        only `grid` / `grid_grad` / `param` (scripts/compare_inference_to_image.py:41-49) are unaffected."""
        return {"optimizer": {"exp_avg": self.exp_avg.detach().cpu(), "exp_avg_sq": self.exp_avg_sq.detach().cpu(),
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
            self._work_counter.zero_()


def slab_range(n_cells: int, rank: int, world: int):
    """[begin, end) in floats of the contiguous block of cells whose optimiser state `rank` owns."""
    base, extra = divmod(n_cells, world)
    start = rank * base + min(rank, extra)
    return 4 * start, 4 * (start + base + (1 if rank < extra else 0))


class PeerVoxelTrainer(VoxelTrainer):
    """Multi-GPU step whose gradient exchange is fused into the optimiser kernel over NVLink peer memory.

    Every rank keeps a full grid replica and a full local gradient buffer, both in symmetric memory
    (torch.distributed._symmetric_memory), so each process holds a mapped pointer to every peer's copy.  Per step:
      K12   render this rank's rays, scatter-add into the LOCAL gradient buffer            (no communication)
      --    barrier: all partial gradients are complete
      K3p   for the cells this rank owns: sum the partial gradients straight out of the peers' buffers, Adam,
            store the new parameters into every replica (plx_adam_step_peer)              (NVLink loads + stores)
      --    barrier: all replicas updated; then clear the local gradient buffer
    No NCCL collective on the data path, no staging copy; optimiser state and its traffic are sharded world-ways.
    """

    def __init__(self, grid, *args, group=None, **kwargs):
        import torch.distributed._symmetric_memory as symm_mem
        if not (dist.is_available() and dist.is_initialized()):
            raise L.PlxError("PeerVoxelTrainer needs an initialised NCCL process group")
        group = group or dist.group.WORLD
        super().__init__(grid, *args, group=group, **kwargs)
        dev = self.device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > L.PLX_MAX_PEERS:
            raise L.PlxError(f"at most {L.PLX_MAX_PEERS} peers")
        shape = tuple(self.grid.shape)
        sym_grid = symm_mem.empty(shape, dtype=torch.float32, device=dev)
        sym_grid.copy_(self.grid)
        self._h_grid = symm_mem.rendezvous(sym_grid, group)
        self.grid = sym_grid
        dist.broadcast(self.grid, src=dist.get_global_rank(group, 0), group=group)      # identical replicas to start from
        # two gradient buffers, used on alternate steps: the one consumed by step i is cleared on a side stream while
        # step i+1 renders into the other, so the 16 B/cell clear never sits on the critical path
        self._grads, self._h_grads = [], []
        for _ in range(2):
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
        self._epoch = 0
        self._own_barrier = os.environ.get("PLX_PEER_BARRIER", "own") == "own"
        self._args = self._make_args()
        # fused ordering (PlxPeerSync, opt-in with PLX_PEER_FUSED=1): the march signals "partial gradient complete"
        # (channel 0) and waits for "all slabs of the previous step stored" (channel 1); the exchange kernel waits on 0 and
        # signals 1.  No stand-alone barrier launches on the step; `flush()` (channel 2) is the full barrier before anyone
        # reads a replica from outside.  Measured neutral (N=2: 123.7 vs 126.1 us/step, N=4: 150.3 vs 146.0): the two
        # barrier launches were already hidden behind the kernels they order, what remains is waiting for the slowest
        # rank — so the default stays the simpler contract (a finished stream = a complete replica).
        self._fused = self._own_barrier and os.environ.get("PLX_PEER_FUSED", "0") == "1"
        self._sync_counters = torch.zeros((2,), dtype=torch.int32, device=dev)
        self._sync_render, self._sync_adam = L.PlxPeerSync(), L.PlxPeerSync()
        for k, sy in enumerate((self._sync_render, self._sync_adam)):
            for r in range(self.world):
                sy.flags[r] = int(self._h_flags.buffer_ptrs[r])
            sy.rank, sy.world = self.rank, self.world
            sy.block_counter = self._sync_counters.data_ptr() + 4 * k
        self._flush_epoch = 0
        slab = slab_range(self.grid.numel() // 4, self.rank, self.world)
        self.multicast = False
        self._peers = []
        for b in range(2):
            p = L.PlxAdamPeer()
            p.world, p.rank = self.world, self.rank
            for r in range(self.world):
                p.grids[r] = int(self._h_grid.buffer_ptrs[r])
                p.grads[r] = int(self._h_grads[b].buffer_ptrs[r])
            p.exp_avg, p.exp_avg_sq, p.grad_abs_sum = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.grad_abs_sum.data_ptr()
            p.begin, p.end = slab
            p.lr, p.beta1, p.beta2, p.eps = self.lr, self.betas[0], self.betas[1], self.eps
            # NVLS multicast mappings of the same buffers, when the fabric offers them (in-switch reduce / replicate)
            try:
                mc_grid, mc_grad = int(self._h_grid.multicast_ptr or 0), int(self._h_grads[b].multicast_ptr or 0)
                if mc_grid and mc_grad:
                    p.grid_mc, p.grad_mc = mc_grid, mc_grad
                    self.multicast = True
            except Exception:       # no multicast support: per-peer pointers are used
                pass
            self._peers.append(p)
        self._peer = self._peers[0]
        self._cur = 0
        self.launches_per_step = 2 if self._fused else (4 if self._own_barrier else 2)   # march, [barrier], exchange+Adam, [barrier]
        torch.cuda.synchronize(dev)
        dist.barrier(group)

    def _select_buffer(self) -> None:
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

    def render_phase(self, uv=None) -> None:
        self._select_buffer()
        super().render_phase(uv)

    def _exchange_and_update(self, st, result_host=None):
        b = self._cur
        peer = self._peers[b]
        peer.step = self.step_count
        s = self.step_count & 1
        peer.loss_src = self._loss2.data_ptr() + 4 * s
        peer.loss_clear = self._loss2.data_ptr() + 4 * (1 - s)
        peer.result_host = result_host
        peer.counter_clear = self._work_counter.data_ptr() if self._dynamic else None
        h = self._h_grads[b]
        self._epoch += 1
        if self._fused:
            sy = peer.sync
            sy.flags, sy.rank, sy.world, sy.block_counter = self._sync_adam.flags, self.rank, self.world, self._sync_adam.block_counter
            sy.wait_channel, sy.wait_epoch = 0, self._epoch
            sy.signal_channel, sy.signal_epoch = 1, self._epoch
        else:
            self._barrier(h, 0, st)                           # every rank's partial gradient is complete
        # the buffer consumed by the PREVIOUS step (its readers passed that step's closing barrier / signalled channel 1,
        # which this step's march waited for) is cleared now, on
        # the side stream, behind this step's exchange kernel: that kernel is NVLink-bound and leaves HBM idle, whereas
        # clearing during the march (measured) slowed the march by as much as the clear itself takes
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
            self._barrier(h, 1, st)                           # every replica holds the new parameters; peers done reading
        self._dirty = b

    def flush(self) -> None:
        """Full cross-rank barrier on the stream: after it, this rank's replica holds every peer's slab of the latest step.
        With the fused ordering a step only guarantees that to the NEXT step's march; call this before reading `grid`
        from anywhere else (checkpoint, evaluation render, end of a timed region)."""
        if not self._fused:
            return
        self._flush_epoch += 1
        with torch.cuda.device(self.device):
            L.check(self.lib.plx_peer_barrier(self._flag_ptrs, self.rank, self.world, 2, self._flush_epoch,
                                              L.stream_ptr(self.device)), "plx_peer_barrier(flush)")

    def checkpoint(self, extra_param=None) -> dict:
        self.flush()
        return super().checkpoint(extra_param)

    def _barrier(self, handle, channel, st):
        if self._own_barrier:
            L.check(self.lib.plx_peer_barrier(self._flag_ptrs, self.rank, self.world, channel, self._epoch, st), "plx_peer_barrier")
        else:
            handle.barrier(channel=channel)

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
            self._exchange_and_update(st, self.result_host.data_ptr())
        return self.loss_host

    def gathered_grad_abs_sum(self):
        """Full `grid_grad` (scripts/train.py:184): each rank accumulated |grad| for the cells it owns only."""
        self.flush()
        full = self.grad_abs_sum.clone()
        b, e = self._peers[0].begin, self._peers[0].end
        flat = full.view(-1)
        mask = torch.zeros_like(flat)
        mask[b:e] = 1
        flat *= mask
        all_reduce_sum_(full, self.group)
        return full
