"""Drop-in for the reference's `src/grid_functions.py`: same names, arguments, return layouts — CUDA kernels underneath.

Only the functions on the hot path launch kernels (`get_nearest_voxels`, `trilinear_interpolation` and friends);
`generate_grid` and `average_pool3d_grid` are tensor construction / a library pooling call that SURVEY.md §8f leaves
for a later round.  Every compute entry point refuses CPU tensors: there is no fallback path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L
from . import ops


class GridCoords(torch.Tensor):
    """The (.., 3) world-coordinate tensors `generate_grid` returns, remembering which cells they hold.

    `normalize_samples_for_indecies` needs `grid_indices.min(0)[0]` (src/ray_sampling.py:13) — a reduction over G^3 x 3
    values every step for a constant.  For pd > 0 that minimum is the first cell's coordinate, which the host can form
    exactly as `generate_grid` does (fl32(i - ceil(s/2) + 1) * fl32(pd)), and it survives the two things the reference
    does to these tensors: positive-step slicing of the three leading axes (`grid_grid[start::stride, ...]`,
    scripts/train.py:114-115) and `reshape(-1, 3)` (:116, :122).  Any other operation returns a plain tensor and the
    reduction is then done for real; an operation that WRITES into a GridCoords (in-place / `out=` / slice assignment)
    drops the cached origin of every GridCoords argument, because views may share the edited storage.
    """

    @staticmethod
    def wrap(t: torch.Tensor, meta) -> "GridCoords":
        out = t.as_subclass(GridCoords)
        out._plx = meta
        return out

    @staticmethod
    def _writes(func, kwargs) -> bool:
        name = getattr(func, "__name__", "")
        if kwargs.get("out") is not None or name in ("__setitem__", "__set__", "set_", "resize_"):
            return True
        return (name.endswith("_") and not name.endswith("__")) or (name.startswith("__i") and name.endswith("__") and name != "__index__"
                                                                     and name != "__init__" and name != "__iter__" and name != "__invert__")

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if cls._writes(func, kwargs):
            for a in list(args) + ([kwargs["out"]] if isinstance(kwargs.get("out"), torch.Tensor) else []):
                if isinstance(a, GridCoords):
                    meta = a.__dict__.get("_plx")
                    if meta is not None:
                        meta["pd"] = -1.0          # shared by every view derived from the same tensor: all of them fall back
                    a.__dict__.pop("_plx", None)
        with torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **kwargs)
        if not isinstance(out, torch.Tensor):
            return out
        out = out.as_subclass(torch.Tensor)
        src = args[0] if args and isinstance(args[0], GridCoords) else None
        meta = getattr(src, "_plx", None) if src is not None else None
        if meta is None:
            return out
        name = getattr(func, "__name__", "")
        if name in ("reshape", "view", "contiguous", "detach", "clone") and out.numel() == src.numel():
            return GridCoords.wrap(out, meta)
        if name == "__getitem__" and src.dim() == 4:
            key = args[1] if isinstance(args[1], tuple) else (args[1],)
            if len(key) <= 3 and all(isinstance(k, slice) for k in key):
                start, step = list(meta["start"]), list(meta["step"])
                for a, k in enumerate(key):
                    b, _, st = k.indices(src.shape[a])
                    if st <= 0:
                        return out
                    start[a] += b * step[a]
                    step[a] *= st
                if out.numel():
                    return GridCoords.wrap(out, dict(meta, start=tuple(start), step=tuple(step), root=meta.get("root", meta)))
        return out


def coords_origin(grid_indices: torch.Tensor):
    """(gx, gy, gz) = grid_indices.min(0)[0] as python floats holding fp32 values."""
    meta = getattr(grid_indices, "_plx", None) if isinstance(grid_indices, GridCoords) else None
    if meta is not None and meta["pd"] > 0 and meta.get("root", meta)["pd"] > 0:
        return ops.grid_origin(meta["dims"], meta["pd"], meta["start"])
    flat = grid_indices.as_subclass(torch.Tensor).reshape(-1, grid_indices.shape[-1])
    return tuple(float(x) for x in flat.min(0)[0].detach().cpu().tolist())


def generate_grid(sx, sy, sz, points_distance=0.5, info_size=4, device="cuda"):
    """Grid with `points_distance` spacing centred on the origin — src/grid_functions.py:184-217.

    Returns (grid_coords (sx*sy*sz,3) f32, grid_cells (sx,sy,sz,info_size) f32 zeros requires_grad,
    meshgrid (sx,sy,sz,3) i64, grid_grid (sx,sy,sz,3) f32), as the reference does.
    """
    ix, iy, iz = (torch.arange(s, device=device) for s in (sx, sy, sz))
    cx, cy, cz = torch.meshgrid(ix, iy, iz, indexing="ij")
    meshgrid = torch.stack([cx, cy, cz], dim=-1)
    world = [(c - np.ceil(s / 2) + 1) * points_distance for c, s in ((cx, sx), (cy, sy), (cz, sz))]   # :205-209
    grid_grid = torch.stack(world, dim=-1)
    grid_coords = grid_grid.reshape(sx * sy * sz, 3)
    grid_cells = torch.zeros([sx, sy, sz, info_size], requires_grad=True, device=device)
    meta = dict(dims=(int(sx), int(sy), int(sz)), pd=float(points_distance), start=(0, 0, 0), step=(1, 1, 1))
    return GridCoords.wrap(grid_coords, meta), grid_cells, meshgrid, GridCoords.wrap(grid_grid, meta)


def find_out_of_bound(indecies, grid):
    """True where the (N,3) indices / coordinates lie inside the grid — src/grid_functions.py:47-63."""
    X, Y, Z, _ = grid.shape
    i0, i1, i2 = indecies[:, 0], indecies[:, 1], indecies[:, 2]
    return ((i0 < X) & (i0 >= 0)) & ((i1 < Y) & (i1 >= 0)) & ((i2 < Z) & (i2 >= 0))


def fix_out_of_bounds(indices, grid):
    """Periodic wrap, IN PLACE through the column views like the reference — src/grid_functions.py:66-79."""
    X, Y, Z, _ = grid.shape
    i0, i1, i2 = indices[:, 0], indices[:, 1], indices[:, 2]
    i0 %= X
    i1 %= Y
    i2 %= Z
    return i0, i1, i2


def get_nearest_voxels(normalized_samples_for_indices, grid, receptive_field=1):
    """Nearest-voxel values (M,4) at periodically wrapped indices + in-bounds mask (M,) — src/grid_functions.py:103-114.
    One gather kernel (round-half-even, mask, wrap, 16-byte cell read); differentiable w.r.t. `grid`."""
    L.require_cuda(normalized_samples_for_indices, grid)
    from . import lazy
    ns = normalized_samples_for_indices
    if isinstance(ns, lazy.LazyTensor):
        if ns._kind == "normalized" and ns._real is None and grid.dim() == 4 and grid.shape[3] == 4 and grid.dtype == torch.float32:
            return lazy.lazy_lookup(ns, grid)           # stays lazy: see plenoxels_b200/lazy.py
        ns = ns.materialize()
    return ops.gather_nearest(ns, grid)


def get_grid_points_indices(normalized_samples_for_indecies):
    """(N,8,3) int64 corner indices [ccc, ccf, cfc, cff, fcc, fcf, ffc, fff] — src/grid_functions.py:220-246."""
    ns = normalized_samples_for_indecies
    hi, lo = torch.ceil(ns), torch.floor(ns)
    corners = [torch.stack([x[:, 0], y[:, 1], z[:, 2]], dim=1) for x in (hi, lo) for y in (hi, lo) for z in (hi, lo)]
    return torch.stack(corners, dim=1).type(torch.long)


def trilinear_interpolation(normalized_samples_for_indecies, selected_points, grid_cells):
    """(N,info) trilinear values — src/grid_functions.py:7-44.  `selected_points` must be the (wrapped) corners of
    `get_grid_points_indices` for the same samples; the kernel recomputes them from the coordinates."""
    L.require_cuda(normalized_samples_for_indecies, grid_cells)
    from . import lazy
    if isinstance(normalized_samples_for_indecies, lazy.LazyTensor):
        normalized_samples_for_indecies = normalized_samples_for_indecies.materialize()
    vals, _ = ops.trilinear_lookup(normalized_samples_for_indecies, grid_cells, masked=False)
    return vals


def collect_cell_information_via_indices(normalized_samples_for_indices, B):
    """(N*8,info) corner cells + in-bounds mask — src/grid_functions.py:154-170."""
    inb = find_out_of_bound(normalized_samples_for_indices, B)
    pts = get_grid_points_indices(normalized_samples_for_indices)
    flat = pts.reshape(pts.shape[0] * pts.shape[1], pts.shape[2])
    i0, i1, i2 = fix_out_of_bounds(flat, B)
    return B[i0, i1, i2], inb


def average_pool3d_grid(tensor, receptive_field_size=3, stride=None):
    """Strided 3-D average pooling of an (X,Y,Z,4) grid — src/grid_functions.py:173-181.  On CUDA (X,Y,Z,4) fp32 grids this
    is three separable box passes (plx_avgpool3d_fwd/bwd) returning a contiguous (x,y,z,4) tensor; anything else takes
    the library call."""
    if tensor.is_cuda and tensor.dim() == 4 and tensor.shape[3] == 4 and tensor.dtype == torch.float32:
        return ops.avgpool3d_grid(tensor, receptive_field_size, stride)
    inp = tensor.permute(3, 0, 1, 2).unsqueeze(0)
    out = F.avg_pool3d(inp, (receptive_field_size,) * 3, stride=stride)
    return out.squeeze().permute(1, 2, 3, 0)


def filled_circle_kernel_3d(size: int = 3, radius: float = 1.0):
    """src/grid_functions.py:260-267."""
    k = np.linspace(-(size // 2), size // 2, size)
    x, y, z = np.meshgrid(k, k, k)
    kern = np.where(np.sqrt(x * x + y * y + z * z) <= radius, 1, 0)
    kern[size // 2, size // 2, size // 2] = 0
    return torch.from_numpy(kern).float()


def gaussian_kernel_3d(size: int = 3, sigma: float = 1.0):
    """src/grid_functions.py:249-257."""
    k = np.linspace(-(size // 2), size // 2, size)
    x, y, z = np.meshgrid(k, k, k)
    kern = np.exp(-(x * x + y * y + z * z) / (2.0 * sigma ** 2))
    kern[size // 2, size // 2, size // 2] = 0
    return torch.from_numpy(kern / kern.sum()).float()


def convolve_grid_to_remove_noise(grid_cells, kernel_size=3, radius=1.0, threshold=2, repeats=20):
    """Zero voxels with fewer than `threshold` occupied neighbours — src/grid_functions.py:270-280 (viewer prep, off path)."""
    w = filled_circle_kernel_3d(kernel_size, radius).unsqueeze(0).unsqueeze(0).to(grid_cells.device)
    w /= w.max()
    for _ in range(repeats):
        occ = F.conv3d(grid_cells[..., -1].unsqueeze(0), w, padding="same")[0]
        grid_cells[occ < threshold] = 0
    return grid_cells
