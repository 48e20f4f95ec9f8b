"""ctypes binding of libplenoxel_b200.so (include/plenoxel_abi.h).

There is no CPU fallback: if the shared library is missing or fails to load, importing a compute entry point raises.
The library is built in-tree by `plenoxels_b200.build` / `__graft_entry__.build()`.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import build as _build

c_f32p = C.POINTER(C.c_float)
c_void = C.c_void_p

PLX_NEAREST, PLX_TRILINEAR = 0, 1
PLX_CLAMP01, PLX_NO_CLIP, PLX_NO_EARLY_STOP, PLX_COHERENT_RAYS = 1, 2, 4, 8
PLX_STEP_RENDER, PLX_STEP_OPTIM, PLX_STEP_ALL, PLX_STEP_UNFUSED = 1, 2, 3, 4
PLX_IMG_F32, PLX_IMG_U8 = 0, 1
PLX_E_PEER_TIMEOUT = -5
ABI_VERSION = 2
MODES = {"nearest": PLX_NEAREST, "trilinear": PLX_TRILINEAR}


class PlxError(RuntimeError):
    """A libplenoxel_b200 call returned a non-zero code."""


class PlxMarch(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("num_samples", C.c_int32),
                ("sx", C.c_int64), ("sy", C.c_int64), ("sz", C.c_int64), ("sc", C.c_int64),
                ("gmin", C.c_float * 3), ("points_distance", C.c_float), ("delta_step", C.c_float),
                ("mode", C.c_int32), ("flags", C.c_uint32)]


class PlxRays(C.Structure):
    _fields_ = [("origins", c_void), ("dirs", c_void), ("n_rays", C.c_int64), ("rays_per_origin", C.c_int64),
                ("origin_stride", C.c_int64), ("origin_comp_stride", C.c_int64)]


class PlxRenderFwd(C.Structure):
    _fields_ = [("march", PlxMarch), ("rays", PlxRays), ("grid", c_void), ("rgba", c_void), ("depth", c_void),
                ("count", c_void), ("sample_index", c_void), ("tcarry", c_void), ("targets", c_void),
                ("grad_rgba", c_void), ("loss", c_void), ("grad_scale", C.c_float), ("loss_scale", C.c_float),
                ("image_u8", c_void), ("image_side", C.c_int32)]


class PlxRenderBwd(C.Structure):
    _fields_ = [("march", PlxMarch), ("rays", PlxRays), ("grid", c_void), ("grad_rgba", c_void), ("tcarry", c_void),
                ("grad_grid", c_void), ("beta_over_m", C.c_float)]


class PlxRayGen(C.Structure):
    _fields_ = [("imgs", c_void), ("n_cams", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
                ("poses", c_void), ("fov", C.c_float), ("uv", c_void), ("rays_per_cam", C.c_int32),
                ("img_format", C.c_int32)]


PLX_MAX_PEERS = 8


class PlxPeerError(C.Structure):
    _fields_ = [("device_word", c_void), ("host_word", c_void), ("timeout_ns", C.c_uint64)]


class PlxPeerGrad(C.Structure):
    _fields_ = [("grads", c_void * PLX_MAX_PEERS), ("owner_mul", C.c_uint32), ("world", C.c_int32)]


class PlxPeerSync(C.Structure):
    _fields_ = [("flags", c_void * PLX_MAX_PEERS), ("rank", C.c_int32), ("world", C.c_int32),
                ("wait_channel", C.c_int32), ("wait_epoch", C.c_int32), ("signal_channel", C.c_int32),
                ("signal_epoch", C.c_int32), ("block_counter", c_void), ("err", PlxPeerError)]


class PlxReplayState(C.Structure):
    _fields_ = [("step_dev", c_void), ("table", c_void), ("table_base", C.c_int64), ("table_len", C.c_int32),
                ("block_counter", c_void)]


class PlxRenderTrain(C.Structure):
    _fields_ = [("march", PlxMarch), ("rays", PlxRays), ("targets", c_void), ("gen", PlxRayGen), ("grid", c_void),
                ("grad_grid", c_void), ("rgba", c_void), ("loss", c_void), ("grad_scale", C.c_float),
                ("loss_scale", C.c_float), ("beta_over_m", C.c_float), ("sync", PlxPeerSync), ("peer_grad", PlxPeerGrad),
                ("step_dev", c_void)]


class PlxAdamPeer(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("grids", c_void * PLX_MAX_PEERS),
                ("grads", c_void * PLX_MAX_PEERS), ("exp_avg", c_void), ("exp_avg_sq", c_void), ("grad_abs_sum", c_void),
                ("begin", C.c_int64), ("end", C.c_int64), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("step", C.c_int64), ("grid_mc", c_void), ("grad_mc", c_void),
                ("loss_src", c_void), ("loss_clear", c_void), ("result_host", c_void),
                ("loss_peers", c_void * PLX_MAX_PEERS), ("loss_out", c_void), ("sync", PlxPeerSync)]


class PlxAdamSlab(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("grids", c_void * PLX_MAX_PEERS), ("grid_mc", c_void),
                ("grad", c_void), ("exp_avg", c_void), ("exp_avg_sq", c_void), ("grad_abs_sum", c_void),
                ("begin", C.c_int64), ("end", C.c_int64), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("step", C.c_int64), ("loss_peers", c_void * PLX_MAX_PEERS), ("loss_out", c_void),
                ("loss_clear", c_void), ("result_host", c_void), ("err", PlxPeerError)]


class PlxTrainStep(C.Structure):
    _fields_ = [("march", PlxMarch),
                ("imgs", c_void), ("n_cams", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
                ("poses", c_void), ("fov", C.c_float),
                ("uv", c_void), ("rays_per_cam", C.c_int32),
                ("n_rays_global", C.c_int64),
                ("grid", c_void), ("grad", c_void), ("exp_avg", c_void), ("exp_avg_sq", c_void),
                ("grad_abs_sum", c_void),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("step", C.c_int64),
                ("beta_over_m", C.c_float),
                ("dirs", c_void), ("targets", c_void), ("rgba", c_void), ("grad_rgba", c_void), ("tcarry", c_void),
                ("loss", c_void), ("render_sync", C.POINTER(PlxPeerSync)), ("peer_grad", C.POINTER(PlxPeerGrad)),
                ("img_format", C.c_int32), ("replay", C.POINTER(PlxReplayState))]


# name -> (restype, argtypes); every symbol include/plenoxel_abi.h declares
PROTOTYPES = {
    "plx_version": (C.c_int, []),
    "plx_last_error": (C.c_char_p, []),
    "plx_num_chunks": (C.c_int32, [C.c_int32]),
    "plx_render_fwd": (C.c_int, [C.POINTER(PlxRenderFwd), c_void]),
    "plx_render_bwd": (C.c_int, [C.POINTER(PlxRenderBwd), c_void]),
    "plx_render_train": (C.c_int, [C.POINTER(PlxRenderTrain), c_void]),
    "plx_adam_step": (C.c_int, [c_void, c_void, c_void, c_void, c_void, C.c_int64, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_int64, C.c_int32, c_void]),
    "plx_adam_table": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int32, C.POINTER(C.c_float)]),
    "plx_adam_step_peer": (C.c_int, [C.POINTER(PlxAdamPeer), c_void]),
    "plx_adam_step_slab": (C.c_int, [C.POINTER(PlxAdamSlab), c_void]),
    "plx_slab_partition": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_uint32), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int64)]),
    "plx_peer_barrier": (C.c_int, [C.POINTER(c_void), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(PlxPeerError),
                                   c_void]),
    "plx_generate_rays_gen": (C.c_int, [C.POINTER(PlxRayGen), C.c_int32, c_void, c_void, c_void]),
    "plx_generate_rays": (C.c_int, [c_void, C.c_int32, C.c_int32, C.c_int32, c_void, C.c_float, c_void, C.c_int32,
                                    C.c_int32, c_void, c_void, c_void]),
    "plx_sample_points": (C.c_int, [C.POINTER(PlxRays), C.c_int32, C.c_float, c_void, c_void]),
    "plx_normalize_points": (C.c_int, [c_void, C.c_int64, C.POINTER(C.c_float), C.c_float, c_void, c_void]),
    "plx_gather_nearest": (C.c_int, [c_void, C.c_int64, c_void, C.POINTER(C.c_int32), C.POINTER(C.c_int64), c_void,
                                     c_void, c_void, c_void]),
    "plx_gather_nearest_bwd": (C.c_int, [c_void, C.c_int64, c_void, C.POINTER(C.c_int32), c_void, c_void]),
    "plx_trilinear_fwd": (C.c_int, [c_void, C.c_int64, c_void, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_int32,
                                    c_void, c_void, c_void]),
    "plx_trilinear_bwd": (C.c_int, [c_void, C.c_int64, c_void, C.POINTER(C.c_int32), C.c_int32, c_void, c_void]),
    "plx_composite_fwd": (C.c_int, [c_void, C.c_int64, C.c_int32, c_void, c_void]),
    "plx_composite_bwd": (C.c_int, [c_void, C.c_int64, C.c_int32, c_void, c_void, c_void]),
    "plx_avgpool3d_fwd": (C.c_int, [c_void, C.POINTER(C.c_int32), C.c_int32, C.c_int32, c_void, c_void, c_void, c_void]),
    "plx_avgpool3d_bwd": (C.c_int, [c_void, C.POINTER(C.c_int32), C.c_int32, C.c_int32, c_void, c_void, c_void, c_void]),
    "plx_tv_loss": (C.c_int, [c_void, C.POINTER(C.c_int32), C.c_float, c_void, c_void, c_void, c_void]),
    "plx_tv_loss_range": (C.c_int, [c_void, C.POINTER(C.c_int32), C.c_float, c_void, C.c_int64, C.c_int64, C.c_int32, c_void,
                                    c_void, c_void]),
    "plx_splat_view": (C.c_int, [c_void, C.POINTER(C.c_int32), C.c_float, C.POINTER(C.c_float), C.c_float, C.c_int32, C.c_int32,
                                 c_void, c_void, c_void]),
    "plx_tune": (C.c_int, [C.c_char_p, C.c_int32]),
    "plx_selftest_arith": (C.c_int, [C.c_float, C.c_uint64, C.c_uint64, c_void, c_void]),
    "plx_train_step": (C.c_int, [C.POINTER(PlxTrainStep), C.c_int32, c_void]),
    "plx_train_step_host": (C.c_int, [C.POINTER(PlxTrainStep), c_void, c_void, C.c_int32, c_void]),
}

_lib = None


def lib_path() -> str:
    """The in-tree library.  `PLX_AB_LIBRARY` (tools/ab_step.py only) points at another BUILD of the same sources, so that two
    versions of a kernel can be timed in the same run on the same GPU; it is still this library or nothing (no fallback)."""
    return os.environ.get("PLX_AB_LIBRARY") or _build.LIB_PATH


def load():
    """Load (once) and return the shared library; raises if it is not built — there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise PlxError(f"{path} is missing: run `python -m plenoxels_b200.build` (or __graft_entry__.build()); "
                       "plenoxels_b200 has no CPU / PyTorch fallback")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.plx_version() != ABI_VERSION:
        raise PlxError(f"ABI version mismatch: library reports {lib.plx_version()}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().plx_last_error().decode(errors="replace")
        raise PlxError(f"{what or 'plx call'} failed with code {rc}: {msg}")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """All tensors must live on one CUDA device (the library has no host implementation)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise PlxError("plenoxels_b200 runs on CUDA (sm_100a) only: got a tensor on "
                           f"{t.device}; there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise PlxError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def make_march(grid: torch.Tensor, num_samples: int, delta_step: float, gmin, points_distance: float, mode: str,
               clamp: bool, no_clip: bool = False, no_early_stop: bool = False, coherent: bool = False) -> PlxMarch:
    if grid.dim() != 4 or grid.shape[3] != 4:
        raise PlxError(f"grid must be (X,Y,Z,4), got {tuple(grid.shape)}")
    if grid.dtype != torch.float32:
        raise PlxError(f"grid must be float32, got {grid.dtype}")
    m = PlxMarch()
    m.nx, m.ny, m.nz = (int(s) for s in grid.shape[:3])
    m.num_samples = int(num_samples)
    m.sx, m.sy, m.sz, m.sc = (int(s) for s in grid.stride())
    g = [float(x) for x in gmin]
    m.gmin[0], m.gmin[1], m.gmin[2] = g
    m.points_distance = float(points_distance)
    m.delta_step = float(delta_step)
    m.mode = MODES[mode]
    m.flags = ((PLX_CLAMP01 if clamp else 0) | (PLX_NO_CLIP if no_clip else 0) | (PLX_NO_EARLY_STOP if no_early_stop else 0)
               | (PLX_COHERENT_RAYS if coherent else 0))
    return m


def make_rays(origins: torch.Tensor, dirs: torch.Tensor, rays_per_origin: int) -> PlxRays:
    """origins: (n_origins,3) with any strides (e.g. the view poses[:, :3, 3]); dirs: (N,3) contiguous fp32."""
    if dirs.dim() != 2 or dirs.shape[1] != 3 or origins.dim() != 2 or origins.shape[1] != 3:
        raise PlxError(f"origins/dirs must be (n,3), got {tuple(origins.shape)} / {tuple(dirs.shape)}")
    if dirs.dtype != torch.float32 or origins.dtype != torch.float32:
        raise PlxError("origins/dirs must be float32")
    if not dirs.is_contiguous():
        raise PlxError("dirs must be contiguous")
    n = dirs.shape[0]
    if origins.shape[0] * rays_per_origin < n:
        raise PlxError(f"{origins.shape[0]} origins x {rays_per_origin} rays/origin < {n} rays")
    r = PlxRays()
    r.origins, r.dirs = origins.data_ptr(), dirs.data_ptr()
    r.n_rays, r.rays_per_origin = n, int(rays_per_origin)
    r.origin_stride, r.origin_comp_stride = (int(s) for s in origins.stride())
    if origins.shape[0] == 1:
        r.origin_stride = 0
    return r
