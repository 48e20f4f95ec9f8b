"""ctypes front end of oracle/plenoxel_oracle.c — TEST INFRASTRUCTURE ONLY (same rule as plenoxel_oracle.py: nothing under
plenoxels_b200/ may import this).

    build()                      gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC -> oracle/_build/libplenoxel_oracle.so
    render_forward / render_backward / mse_loss / adam_step / generate_rays / train_step
                                 same arguments and results as the numpy oracle's functions of the same name

The C restatement is ~100x faster than the numpy one (and multi-threaded in the forward pass and Adam), so whole
BASELINE-sized batches (12 800 rays x 600 samples on a 128^3 grid) can be checked ray by ray instead of by subsample.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "plenoxel_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libplenoxel_oracle.so")
CFLAGS = ["-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC"]
MODES = {"nearest": 0, "trilinear": 1}
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement if the library is missing or older than its source."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run(["gcc", *CFLAGS, SRC, "-o", LIB, "-lm"], check=True)
    return LIB


def _p(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        f32p, f64p, i64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        lib.plxo_render_forward.argtypes = [f32p, i64p, f32p, f32p, C.c_int64, C.c_int32, C.c_float, f32p, C.c_float, C.c_int32,
                                            C.c_int32, f32p, f32p, i32p, i64p]
        lib.plxo_render_backward.argtypes = [f32p, i64p, f32p, f32p, C.c_int64, C.c_int32, C.c_float, f32p, C.c_float, f64p,
                                             C.c_int32, C.c_int32, C.c_double, f64p]
        lib.plxo_mse_loss.argtypes = [f32p, f32p, C.c_int64, C.c_int64, f64p]
        lib.plxo_mse_loss.restype = C.c_double
        lib.plxo_adam_step.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_int64, C.c_double, C.c_int64, C.c_double, C.c_double,
                                       C.c_double]
        lib.plxo_adam_step.restype = None
        lib.plxo_generate_rays.argtypes = [f32p, C.c_int32, C.c_int32, C.c_int32, f32p, C.c_float, f32p, C.c_int32, f32p, f32p,
                                           i64p]
        lib.plxo_even_spread_uv.argtypes = [C.c_int32, C.c_int32, f32p]
        lib.plxo_tv_loss.argtypes = [f32p, i64p, f64p]
        lib.plxo_tv_loss.restype = C.c_double
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _geometry(grid, origins, dirs, gmin):
    grid, origins, dirs = _f32(grid), _f32(origins), _f32(dirs)
    if origins.shape != dirs.shape:
        raise ValueError("one origin per ray: origins and dirs must both be (N,3)")
    dims = np.asarray(grid.shape[:3], dtype=np.int64)
    return grid, origins, dirs, dims, _f32(gmin)


def render_forward(grid, origins, dirs, num_samples, delta_step, gmin, points_distance, mode="nearest", clamp=True,
                   want_lin=True):
    """rgba (N,4) f32, depth (N,) f32, count (N,) int32, lin (N,S) int64 (-1 = out of bounds; 0 in bounds for trilinear;
    None with want_lin=False)."""
    grid, origins, dirs, dims, gmin = _geometry(grid, origins, dirs, gmin)
    n = dirs.shape[0]
    rgba, depth = np.zeros((n, 4), np.float32), np.zeros(n, np.float32)
    count, lin = np.zeros(n, np.int32), (np.zeros((n, num_samples), np.int64) if want_lin else None)
    load().plxo_render_forward(_p(grid, C.c_float), _p(dims, C.c_int64), _p(origins, C.c_float), _p(dirs, C.c_float), n,
                               int(num_samples), float(np.float32(delta_step)), _p(gmin, C.c_float),
                               float(np.float32(points_distance)), MODES[mode], int(clamp), _p(rgba, C.c_float),
                               _p(depth, C.c_float), _p(count, C.c_int32), _p(lin, C.c_int64))
    return rgba, depth, count, lin


def render_backward(grid, origins, dirs, num_samples, delta_step, gmin, points_distance, grad_rgba, mode="nearest",
                    clamp=True, beta=0.0):
    """d sum(rgba * grad_rgba) [+ beta term] / d raw grid, (X,Y,Z,4) float64."""
    grid, origins, dirs, dims, gmin = _geometry(grid, origins, dirs, gmin)
    n = dirs.shape[0]
    g = np.ascontiguousarray(grad_rgba, dtype=np.float64)
    out = np.zeros(grid.shape, np.float64)
    m = n * num_samples
    rc = load().plxo_render_backward(_p(grid, C.c_float), _p(dims, C.c_int64), _p(origins, C.c_float), _p(dirs, C.c_float), n,
                                     int(num_samples), float(np.float32(delta_step)), _p(gmin, C.c_float),
                                     float(np.float32(points_distance)), _p(g, C.c_double), MODES[mode], int(clamp),
                                     float(beta) / m if (beta and m) else 0.0, _p(out, C.c_double))
    if rc:
        raise MemoryError("plxo_render_backward")
    return out


def mse_loss(pixels, targets, n_global=None):
    pixels, targets = _f32(pixels), _f32(targets)
    grad = np.zeros(pixels.shape, np.float64)
    loss = load().plxo_mse_loss(_p(pixels, C.c_float), _p(targets, C.c_float), pixels.shape[0], int(n_global or 0),
                                _p(grad, C.c_double))
    return float(loss), grad


def adam_step_(p, g, m, v, gabs, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """In place on contiguous float32 arrays (the timing loop of bench.py uses this to avoid four grid-sized copies)."""
    for a in (p, g, m, v, gabs):
        if a.dtype != np.float32 or not a.flags.c_contiguous or a.size != p.size:
            raise ValueError("adam_step_ needs contiguous float32 arrays of equal size")
    load().plxo_adam_step(_p(p, C.c_float), _p(g, C.c_float), _p(m, C.c_float), _p(v, C.c_float), _p(gabs, C.c_float), p.size,
                          float(lr), int(step), float(beta1), float(beta2), float(eps))


def adam_step(p, g, m, v, gabs, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """Returns new (p, m, v, gabs) like the numpy oracle (inputs are not modified)."""
    p, m, v, gabs = (_f32(a).copy() for a in (p, m, v, gabs))
    g = _f32(g)
    load().plxo_adam_step(_p(p, C.c_float), _p(g, C.c_float), _p(m, C.c_float), _p(v, C.c_float), _p(gabs, C.c_float), p.size,
                          float(lr), int(step), float(beta1), float(beta2), float(eps))
    return p, m, v, gabs


def generate_rays(imgs, poses, fov, uv):
    """dirs (C*R,3) f32, targets (C*R,4) f32, pix (C*R,2) int64 as (u_pix, v_pix)."""
    imgs, poses, uv = _f32(imgs), _f32(poses), _f32(uv)
    n_cams, rays = uv.shape[0], uv.shape[1]
    n = n_cams * rays
    dirs, targets, pix = np.zeros((n, 3), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 2), np.int64)
    load().plxo_generate_rays(_p(imgs, C.c_float), n_cams, imgs.shape[1], imgs.shape[2], _p(poses, C.c_float),
                              float(np.float32(fov)), _p(uv, C.c_float), rays, _p(dirs, C.c_float), _p(targets, C.c_float),
                              _p(pix, C.c_int64))
    return dirs, targets, pix


def train_step(grid, m, v, gabs, origins, dirs, targets, num_samples, delta_step, gmin, points_distance, lr, step,
               mode="nearest", n_global=None):
    """One full step of scripts/train.py:130-184 (tv = beta = 0): loss, grad, and the new (grid, m, v, gabs)."""
    rgba, _, _, _ = render_forward(grid, origins, dirs, num_samples, delta_step, gmin, points_distance, mode)
    loss, gpix = mse_loss(rgba, targets, n_global)
    grad = render_backward(grid, origins, dirs, num_samples, delta_step, gmin, points_distance, gpix, mode)
    p2, m2, v2, ga2 = adam_step(grid, grad.astype(np.float32), m, v, gabs, lr, step)
    return loss, grad, p2, m2, v2, ga2


def even_spread_uv(n_cams, number_of_rays):
    """(C, round(sqrt(R))^2, 2) u-major lattice on [0,1]^2 — src/ray_sampling.py:220-223."""
    n = int(np.round(np.sqrt(number_of_rays)))
    uv = np.zeros((n_cams, n * n, 2), np.float32)
    if load().plxo_even_spread_uv(int(n_cams), n, _p(uv, C.c_float)):
        raise MemoryError("plxo_even_spread_uv")
    return uv


def tv_loss(grid):
    """(loss, gradient (X,Y,Z,4) float64) of scripts/train.py:44-65."""
    grid = _f32(grid)
    dims = np.asarray(grid.shape[:3], dtype=np.int64)
    grad = np.zeros(grid.shape, np.float64)
    loss = load().plxo_tv_loss(_p(grid, C.c_float), _p(dims, C.c_int64), _p(grad, C.c_double))
    return float(loss), grad
