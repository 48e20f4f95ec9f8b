"""CPU oracle for the voxel-grid volume-rendering hot path — TEST INFRASTRUCTURE ONLY.

This is a from-scratch numpy restatement of the algorithm the reference
(DanJbk/Plenoxels) runs through torch ATen ops.  It exists to *check* the CUDA
path; nothing under `plenoxels_b200/` may import it.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs use it.

Parity status: **pinned against the reference itself, run in the build
container** — the reference ships no tests or golden vectors, so
`oracle/validate_against_reference.py` imports `/root/reference` and checks every
function below against it (indices / masks / gathered values bit-exact, pixels
and gradients <= 1e-6), and `tests/golden/make_golden.py` freezes reference
outputs as fixtures that `tests/test_oracle_golden.py` re-checks everywhere.

Arithmetic contract: every fp32 operation is rounded
separately (numpy float32 does exactly that — no FMA contraction, true IEEE
division), `round` is half-to-even, integer modulo is non-negative.

Citations are `file:line` in the reference tree.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- grid geometry
def grid_origin(dims, points_distance, start_index: int = 0) -> np.ndarray:
    """World coordinate of cell (start_index,)*3 — what `grid_indices.min(0)[0]` evaluates to.

    src/grid_functions.py:205-209: coords = (arange(s) - ceil(s/2) + 1) * pd, all fp32.
    src/ray_sampling.py:13 reduces that (G^3,3) tensor with `.min(0)` every step; for pd > 0 the
    minimum is the first cell.  Under progressive growing (`scripts/train.py:113-116`) the grid of
    coordinates is sliced `[start::stride]`, so the minimum is the cell at `start_index`.
    """
    out = np.empty(3, dtype=F32)
    pd32 = F32(points_distance)
    for a in range(3):
        centred = F32(start_index - math.ceil(dims[a] / 2) + 1)
        out[a] = centred * pd32
    return out


# --------------------------------------------------------------------------- ray generation
def torch_like_linspace01(n: int) -> np.ndarray:
    """`torch.linspace(0, 1, n)` in fp32 (src/ray_sampling.py:222): symmetric two-sided fill."""
    if n == 1:
        return np.zeros(1, dtype=F32)
    step = F32(1.0) / F32(n - 1)
    i = np.arange(n)
    lo = (F32(0.0) + step * i.astype(F32)).astype(F32)
    # upper half is end - step*(n-1-i) contracted to one fma by ATen's CPU kernel (probed: bit-exact for
    # n = 17..800 on this image; an ulp-level detail that other builds/devices need not share)
    hi = (1.0 - np.float64(step) * (n - 1 - i).astype(np.float64)).astype(F32)
    return np.where(i < n // 2, lo, hi).astype(F32)


def even_spread_uv(n_cams: int, number_of_rays: int) -> np.ndarray:
    """(C, round(sqrt(R))^2, 2) u-major lattice on [0,1]^2 — src/ray_sampling.py:220-223."""
    n = int(np.round(np.sqrt(number_of_rays)))
    line = torch_like_linspace01(n)
    uu, vv = np.meshgrid(line, line, indexing="ij")
    uv = np.stack([uu.reshape(-1), vv.reshape(-1)], axis=-1)
    return np.broadcast_to(uv, (n_cams,) + uv.shape).copy()


def _norm3(v: np.ndarray, fused: bool = True) -> np.ndarray:
    """Euclidean norm over the last axis as torch-CPU computes it for fp32 (probed on this image, 100 %
    bit-exact on 1e5 random vectors): for a contiguous last axis x0^2, then two fused multiply-adds, then
    sqrt (`fused=True`, the ray directions); for the strided pose-column slices `T[:, :3, j]` plain
    mul/add in order (`fused=False`)."""
    v = v.astype(F32)
    if not fused:
        acc = (v[..., 0] * v[..., 0]).astype(F32)
        acc = (acc + (v[..., 1] * v[..., 1]).astype(F32)).astype(F32)
        acc = (acc + (v[..., 2] * v[..., 2]).astype(F32)).astype(F32)
        return np.sqrt(acc).astype(F32)
    d = v.astype(np.float64)
    acc = (v[..., 0] * v[..., 0]).astype(F32)
    acc = (d[..., 1] * d[..., 1] + acc.astype(np.float64)).astype(F32)
    acc = (d[..., 2] * d[..., 2] + acc.astype(np.float64)).astype(F32)
    return np.sqrt(acc).astype(F32)


def generate_rays(imgs: np.ndarray, poses: np.ndarray, fov: float, uv: np.ndarray):
    """Ray directions, target pixels and pixel indices for given (C,R,2) uv — src/ray_sampling.py:212-264.

    Returns dirs (C*R,3) f32 camera-major, targets (C*R,4) f32, pix (C*R,2) int64 as (u_pix, v_pix).
    Quirks kept on purpose: u is scaled by imgs.shape[1], v by imgs.shape[2], the
    lookup is imgs[cam, v_pix, u_pix]; angles are linear in u,v; v is negated; no tan().
    """
    poses = poses.astype(F32)
    uv = uv.astype(F32)
    C, R, _ = uv.shape
    X = poses[:, :3, 0]
    Y = poses[:, :3, 1]
    Zn = (-poses[:, :3, 2]).astype(F32)
    aspect = (_norm3(X, fused=False) / _norm3(Y, fused=False)).astype(F32)                                     # :218
    fov32 = F32(fov)
    half = F32(0.5)
    u_ang = (fov32 * (uv[..., 0] - half).astype(F32)).astype(F32)                    # :234
    inv_aspect = (F32(1.0) / aspect).astype(F32)
    v_scale = (fov32 * inv_aspect).astype(F32)                                       # :235
    v_ang = (-(v_scale[:, None] * (uv[..., 1] - half).astype(F32)).astype(F32)).astype(F32)
    H, W = imgs.shape[1], imgs.shape[2]
    u_pix = np.minimum(np.rint((F32(H) * uv[..., 0]).astype(F32)), F32(H - 1)).astype(np.int64)   # :238
    v_pix = np.minimum(np.rint((F32(W) * uv[..., 1]).astype(F32)), F32(W - 1)).astype(np.int64)   # :239
    cam = np.repeat(np.arange(C), R)
    u_pix = u_pix.reshape(-1)
    v_pix = v_pix.reshape(-1)
    targets = imgs[cam, v_pix, u_pix].astype(F32)                                    # :248
    ux = (u_ang[..., None] * X[:, None, :]).astype(F32)                              # :261
    vy = (v_ang[..., None] * Y[:, None, :]).astype(F32)
    dirs = ((ux + vy).astype(F32) + Zn[:, None, :]).astype(F32)
    dirs = (dirs / _norm3(dirs)[..., None]).astype(F32)                              # :262
    return dirs.reshape(C * R, 3), targets, np.stack([u_pix, v_pix], axis=-1)


# --------------------------------------------------------------------------- sample placement / indexing
def sample_steps(num_samples: int, delta_step: float) -> np.ndarray:
    """t_k = fl32(fl32(delta) * fl32(k)), k = 1..S — src/ray_sampling.py:161."""
    k = np.arange(1, num_samples + 1, dtype=np.int64).astype(F32)
    return (F32(delta_step) * k).astype(F32)


def sample_positions(origins: np.ndarray, dirs: np.ndarray, num_samples: int, delta_step: float) -> np.ndarray:
    """(N,S,3) world positions o + d*t with the product and the sum rounded separately — :164-167."""
    t = sample_steps(num_samples, delta_step)
    dt = (dirs.astype(F32)[:, None, :] * t[None, :, None]).astype(F32)
    return (origins.astype(F32)[:, None, :] + dt).astype(F32)


def normalize_positions(pos: np.ndarray, gmin: np.ndarray, points_distance: float) -> np.ndarray:
    """(pos - gmin) / fl32(pd): subtraction then true division — src/ray_sampling.py:13."""
    return ((pos.astype(F32) - gmin.astype(F32)).astype(F32) / F32(points_distance)).astype(F32)


def nearest_indices(ns: np.ndarray, dims):
    """round-half-even -> int64, in-bounds mask — src/grid_functions.py:111, :58-61.

    Returns idx (...,3) int64 (unwrapped) and inb (...) bool (True = inside, despite the reference's name).
    """
    idx = np.rint(ns.astype(F32)).astype(np.int64)
    inb = np.ones(idx.shape[:-1], dtype=bool)
    for a in range(3):
        inb &= (idx[..., a] >= 0) & (idx[..., a] < dims[a])
    return idx, inb


def gather_nearest(ns: np.ndarray, grid: np.ndarray):
    """`get_nearest_voxels` — src/grid_functions.py:103-114: values at periodically wrapped indices
    (unmasked, :75-77 python-style modulo) plus the in-bounds mask."""
    dims = grid.shape[:3]
    idx, inb = nearest_indices(ns, dims)
    w = [np.mod(idx[..., a], dims[a]) for a in range(3)]
    return grid[w[0], w[1], w[2]], inb


def trilinear_lookup(ns: np.ndarray, grid: np.ndarray):
    """Trilinear composition of the reference's (uncalled) pieces, SURVEY.md §8a row T.

    mask: float test 0 <= ns < dim (src/grid_functions.py:58-61 applied to the float coordinates);
    corners ceil/floor (:230-244) wrapped periodically (:75-77); weights frac(ns) (:29);
    lerp x, then y, then z, each as mul, mul, add (:34-42).  Returns masked values (M,4) and the mask.
    """
    ns = ns.astype(F32)
    dims = grid.shape[:3]
    mask = np.ones(ns.shape[:-1], dtype=bool)
    for a in range(3):
        mask &= (ns[..., a] >= 0) & (ns[..., a] < dims[a])
    cl = [np.mod(np.ceil(ns[..., a]).astype(np.int64), dims[a]) for a in range(3)]
    fl = [np.mod(np.floor(ns[..., a]).astype(np.int64), dims[a]) for a in range(3)]
    frac = (ns - np.trunc(ns)).astype(F32)                      # torch.frac keeps the sign
    fx, fy, fz = (frac[..., a][..., None] for a in range(3))
    one = F32(1.0)

    def cell(ix, iy, iz):
        return grid[ix, iy, iz].astype(F32)

    def lerp(hi, lo, f):
        return ((hi * f).astype(F32) + (lo * (one - f).astype(F32)).astype(F32)).astype(F32)

    # x-lerp: ceil_x corners weighted by f_x, floor_x by 1-f_x; corner order [cc, cf, fc, ff] over (y,z)
    x_cc = lerp(cell(cl[0], cl[1], cl[2]), cell(fl[0], cl[1], cl[2]), fx)
    x_cf = lerp(cell(cl[0], cl[1], fl[2]), cell(fl[0], cl[1], fl[2]), fx)
    x_fc = lerp(cell(cl[0], fl[1], cl[2]), cell(fl[0], fl[1], cl[2]), fx)
    x_ff = lerp(cell(cl[0], fl[1], fl[2]), cell(fl[0], fl[1], fl[2]), fx)
    y_c = lerp(x_cc, x_fc, fy)
    y_f = lerp(x_cf, x_ff, fy)
    out = lerp(y_c, y_f, fz)
    return (out * mask[..., None]).astype(F32), mask


# --------------------------------------------------------------------------- compositing
def composite(samples: np.ndarray, steps: np.ndarray | None = None, dtype=F32):
    """Front-to-back alpha compositing — src/ray_sampling.py:181-191.

    samples (..., S, 4) -> rgba (..., 4); T_k = prod_{j<k} (1 - alpha_j) sequentially, w = alpha*T.
    With `steps` (S,) also returns depth = sum_k w_k t_k (not in the reference).
    """
    s = samples.astype(dtype)
    alpha = s[..., 3]
    S = alpha.shape[-1]
    T = np.ones(alpha.shape[:-1], dtype=dtype)
    acc = np.zeros(alpha.shape[:-1] + (4,), dtype=dtype)
    depth = np.zeros(alpha.shape[:-1], dtype=dtype)
    one = dtype(1.0)
    for k in range(S):
        w = (alpha[..., k] * T).astype(dtype)
        acc[..., :3] += (s[..., k, :3] * w[..., None]).astype(dtype)
        acc[..., 3] += w
        if steps is not None:
            depth += (w * dtype(steps[k])).astype(dtype)
        T = (T * (one - alpha[..., k])).astype(dtype)
    if steps is not None:
        return acc, depth
    return acc


def composite_backward(samples: np.ndarray, grad_rgba: np.ndarray, dtype=np.float64):
    """d loss / d samples for `composite`, division-free reverse recurrence.

    v_k = c_k . g_rgb + g_A;  d c_k = alpha_k T_k g_rgb;  d alpha_k = T_k (v_k - S_k),
    S_k = alpha_{k+1} v_{k+1} + (1 - alpha_{k+1}) S_{k+1},  S_last = 0.
    """
    s = samples.astype(dtype)
    g = grad_rgba.astype(dtype)
    alpha = s[..., 3]
    S = alpha.shape[-1]
    T = np.ones(alpha.shape, dtype=dtype)
    for k in range(1, S):
        T[..., k] = T[..., k - 1] * (1.0 - alpha[..., k - 1])
    v = (s[..., :3] * g[..., None, :3]).sum(-1) + g[..., None, 3]
    out = np.zeros_like(s)
    behind = np.zeros(alpha.shape[:-1], dtype=dtype)
    for k in range(S - 1, -1, -1):
        out[..., k, 3] = T[..., k] * (v[..., k] - behind)
        out[..., k, :3] = (alpha[..., k] * T[..., k])[..., None] * g[..., :3]
        behind = alpha[..., k] * v[..., k] + (1.0 - alpha[..., k]) * behind
    return out


# --------------------------------------------------------------------------- fused restatement of the step
def render_forward(grid, origins, dirs, num_samples, delta_step, gmin, points_distance,
                   mode: str = "nearest", clamp: bool = True, dtype=F32):
    """rays -> rgba (N,4), depth (N,), count (N,) int32, lin_idx (N,S) int64 (-1 = out of bounds).

    Sequence of scripts/train.py:130-151 (or src/visualization.py:125-146 with clamp done by the caller):
    sample placement, normalisation, `grid.clip(0,1)` lookup, mask multiply, compositing.
    """
    grid = np.asarray(grid, dtype=F32)
    dims = grid.shape[:3]
    pos = sample_positions(origins, dirs, num_samples, delta_step)
    ns = normalize_positions(pos, gmin, points_distance)
    g = np.clip(grid, F32(0), F32(1)) if clamp else grid
    if mode == "nearest":
        vals, inb = gather_nearest(ns, g)
        vals = (vals * inb[..., None]).astype(F32)                                   # scripts/train.py:147
        idx, _ = nearest_indices(ns, dims)
        lin = (idx[..., 0] * dims[1] + idx[..., 1]) * dims[2] + idx[..., 2]
        lin = np.where(inb, lin, -1)
    elif mode == "trilinear":
        vals, inb = trilinear_lookup(ns, g)
        lin = np.where(inb, 0, -1)
    else:
        raise ValueError(mode)
    steps = sample_steps(num_samples, delta_step)
    if num_samples == 0:
        N = origins.shape[0]
        return np.zeros((N, 4), dtype), np.zeros(N, dtype), np.zeros(N, np.int32), lin
    rgba, depth = composite(vals, steps, dtype=dtype)
    return rgba, depth, inb.sum(-1).astype(np.int32), lin


def mse_loss(pixels: np.ndarray, targets: np.ndarray, n_global: int | None = None):
    """mean over N*4 elements incl. alpha (scripts/train.py:156, A8); returns loss and d loss / d pixels."""
    n = pixels.shape[0] if n_global is None else n_global
    diff = pixels.astype(np.float64) - targets.astype(np.float64)
    return float((diff ** 2).sum() / (4 * n)), (2.0 * diff / (4 * n))


def render_backward(grid, origins, dirs, num_samples, delta_step, gmin, points_distance, grad_rgba,
                    mode: str = "nearest", clamp: bool = True, beta: float = 0.0, dtype=np.float64):
    """Gradient of sum(rgba * grad_rgba) [+ beta term] w.r.t. the raw grid, (X,Y,Z,4).

    Autograd of scripts/train.py:146-181 restated: composite backward -> mask -> scatter-add at the
    nearest cell (index_put accumulate) -> clip pass-mask 0 <= raw <= 1 inclusive (A6).
    beta term (scripts/train.py:170-177): beta * mean_M(log(a+eps) - log(1-a+eps)) over ALL M samples
    (masked samples have a = 0 and no gradient path).
    """
    grid = np.asarray(grid, dtype=F32)
    dims = grid.shape[:3]
    pos = sample_positions(origins, dirs, num_samples, delta_step)
    ns = normalize_positions(pos, gmin, points_distance)
    g = np.clip(grid, F32(0), F32(1)) if clamp else grid
    passmask = ((grid >= 0) & (grid <= 1)) if clamp else np.ones(grid.shape, bool)
    out = np.zeros(grid.shape, dtype=np.float64)
    flat = out.reshape(-1, 4)
    ncell = flat.shape[0]
    if mode == "nearest":
        vals, inb = gather_nearest(ns, g)
        vals = vals * inb[..., None]
        dvals = composite_backward(vals, grad_rgba, dtype=dtype)
        if beta:
            eps = 1e-4
            a = vals[..., 3].astype(np.float64)
            dvals[..., 3] += beta / a.size * (1.0 / (a + eps) + 1.0 / (1.0 - a + eps))
        idx, _ = nearest_indices(ns, dims)
        lin = ((idx[..., 0] * dims[1] + idx[..., 1]) * dims[2] + idx[..., 2])[inb]
        dv = dvals[inb]
        for c in range(4):
            flat[:, c] += np.bincount(lin, weights=dv[:, c], minlength=ncell)
    elif mode == "trilinear":
        vals, inb = trilinear_lookup(ns, g)
        dvals = composite_backward(vals, grad_rgba, dtype=dtype)
        if beta:
            eps = 1e-4
            a = vals[..., 3].astype(np.float64)
            dvals[..., 3] += beta / a.size * (1.0 / (a + eps) + 1.0 / (1.0 - a + eps))
        nsf = ns[inb].astype(np.float64)
        dv = dvals[inb]
        fr = nsf - np.trunc(nsf)
        cl = [np.mod(np.ceil(nsf[:, a]).astype(np.int64), dims[a]) for a in range(3)]
        fl = [np.mod(np.floor(nsf[:, a]).astype(np.int64), dims[a]) for a in range(3)]
        for cx in (0, 1):
            for cy in (0, 1):
                for cz in (0, 1):
                    ix = cl[0] if cx == 0 else fl[0]
                    iy = cl[1] if cy == 0 else fl[1]
                    iz = cl[2] if cz == 0 else fl[2]
                    w = ((fr[:, 0] if cx == 0 else 1 - fr[:, 0]) * (fr[:, 1] if cy == 0 else 1 - fr[:, 1])
                         * (fr[:, 2] if cz == 0 else 1 - fr[:, 2]))
                    lin = (ix * dims[1] + iy) * dims[2] + iz
                    for c in range(4):
                        flat[:, c] += np.bincount(lin, weights=dv[:, c] * w, minlength=ncell)
    else:
        raise ValueError(mode)
    return out * passmask


# --------------------------------------------------------------------------- regulariser
def tv_loss(grid, dtype=np.float64):
    """`tv_loss` of scripts/train.py:44-65: sqrt of the summed squared neighbour differences along the three grid axes
    (all four channels), and its gradient w.r.t. the grid (autograd of :63).  At an all-equal grid the reference's
    gradient is 0/0 = NaN; this restatement (and the kernel) return 0 there."""
    g = np.asarray(grid).astype(dtype)
    d1 = g[:, :-1] - g[:, 1:]
    d2 = g[:, :, :-1] - g[:, :, 1:]
    d0 = g[:-1] - g[1:]
    s = (d1 ** 2).sum() + (d2 ** 2).sum() + (d0 ** 2).sum()
    loss = float(np.sqrt(s))
    grad = np.zeros_like(g)
    if loss > 0:
        grad[:, :-1] += d1
        grad[:, 1:] -= d1
        grad[:, :, :-1] += d2
        grad[:, :, 1:] -= d2
        grad[:-1] += d0
        grad[1:] -= d0
        grad /= loss
    return loss, grad


# --------------------------------------------------------------------------- optimiser
def adam_step(p, g, m, v, gabs, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """One torch.optim.Adam step (torch/optim/adam.py `_single_tensor_adam`, non-capturable branch,
    as driven by scripts/train.py:89,:180-184) plus `grid_grad += |grad|`.  fp32 element ops, python-double
    scalars; FMA placement follows ATen's CPU vector kernels (m: fma(w, g-m, m); v: fma((1-b2)*g, g, v*b2);
    p: p + ((-lr/bc1)*m)/denom) — probed bit-exact for m and v; p agrees on 99.97 % of elements, the rest
    differ by 1 ulp because torch-CPU's sqrt (MKL VML) is not correctly rounded while numpy's / CUDA's is.  Returns new (p, m, v, gabs).
    """
    p, g, m, v, gabs = (np.asarray(a, dtype=F32) for a in (p, g, m, v, gabs))
    w = F32(1.0 - beta1)
    diff = (g - m).astype(F32)
    m2 = (m.astype(np.float64) + np.float64(w) * diff.astype(np.float64)).astype(F32)      # fma(w, g-m, m)
    v2 = (v * F32(beta2)).astype(F32)
    vg = (F32(1.0 - beta2) * g).astype(F32)
    v2 = (vg.astype(np.float64) * g.astype(np.float64) + v2.astype(np.float64)).astype(F32)   # fma(val*g, g, v*b2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = bc2 ** 0.5
    denom = ((np.sqrt(v2).astype(F32) / F32(bc2_sqrt)).astype(F32) + F32(eps)).astype(F32)
    upd = ((F32(-step_size) * m2).astype(F32) / denom).astype(F32)
    p2 = (p + upd).astype(F32)
    return p2, m2, v2, (gabs + np.abs(g)).astype(F32)


def train_step(grid, m, v, gabs, origins, dirs, targets, num_samples, delta_step, gmin, points_distance,
               lr, step, mode="nearest", n_global=None):
    """One full step of scripts/train.py:130-184 (tv = beta = 0): returns loss, grad, and the new state."""
    rgba, _, _, _ = render_forward(grid, origins, dirs, num_samples, delta_step, gmin, points_distance, mode)
    loss, gpix = mse_loss(rgba, targets, n_global)
    grad = render_backward(grid, origins, dirs, num_samples, delta_step, gmin, points_distance, gpix, mode)
    p2, m2, v2, ga2 = adam_step(grid, grad.astype(F32), m, v, gabs, lr, step)
    return loss, grad, p2, m2, v2, ga2
