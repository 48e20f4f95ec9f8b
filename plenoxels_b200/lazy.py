"""Lazy handles that let the reference's UNFUSED call sequence reach the FUSED kernels.

`fit()` (scripts/train.py:130-151) and `visulize_3d_in_2d` (src/visualization.py:125-146) call six functions and do
tensor arithmetic in between:

    samples, targets, cam, dirs = sample_camera_rays_batched(...)          # (M,3) positions
    ns = normalize_samples_for_indecies(grid_indices, samples, pd)         # (M,3)
    nearest, mask = get_nearest_voxels(ns, grid_cells.clip(0, 1))          # (M,4), (M,)
    nearest = nearest * mask.unsqueeze(-1)
    nearest = nearest.reshape(C, R, S, 4)
    pixels = compute_alpha_weighted_pixels(nearest)                        # (C,R,4)

Materialising every line costs ~25 bytes of HBM traffic per sample and per line.  Instead the first four functions
return `LazyTensor`s: storage-less torch.Tensor subclasses (correct shape / dtype / device) that only remember how they
would be computed.  The operations above are recognised and stay lazy; `compute_alpha_weighted_pixels` on a masked,
reshaped lazy lookup launches ONE fused march (ops.render_rays, K1 forward / K2 backward through autograd).  Any other
operation — e.g. the beta-loss slice `nearest[:, :, :, -1]` (scripts/train.py:172) — materialises the operand with the
eager kernels first and then proceeds as a normal tensor op, so semantics never change, only speed.
"""
from __future__ import annotations

import torch

from . import ops


class LazyTensor(torch.Tensor):
    """Storage-less handle; `kind` in {"samples", "normalized", "lookup", "mask"}."""

    @staticmethod
    def __new__(cls, shape, dtype, device, kind, spec):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=dtype, device=device, requires_grad=False)
        t._kind, t._spec, t._real = kind, spec, None
        return t

    def __init__(self, *a, **k):
        pass

    # ------------------------------------------------------------------------------------------------ materialise
    def materialize(self) -> torch.Tensor:
        """The real tensor this handle stands for, computed with the eager kernels (cached)."""
        if self._real is not None:
            return self._real
        s = self._spec
        if self._kind == "samples":
            out = ops.sample_points(s["origins"], s["dirs"], s["S"], s["delta"], rays_per_origin=s["R"])
        elif self._kind == "normalized":
            out = ops.normalize_points(s["samples"].materialize(), s["gmin"], s["pd"])
        elif self._kind == "mask":
            _, out = ops.gather_nearest(s["ns"].materialize(), s["grid"])
            out = out.reshape(self.shape)
        elif self._kind == "lookup":
            vals, inb = ops.gather_nearest(s["ns"].materialize(), s["grid"])
            if s["masked"]:
                vals = vals * inb.unsqueeze(-1)
            out = vals.reshape(self.shape)
        else:  # pragma: no cover
            raise RuntimeError(self._kind)
        self._real = out
        return out

    def _derive(self, shape, **changes) -> "LazyTensor":
        spec = dict(self._spec)
        spec.update(changes)
        return LazyTensor(shape, self.dtype, self.device, self._kind, spec)

    # ------------------------------------------------------------------------------------------------ interception
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        a0 = args[0] if args else None
        if isinstance(a0, LazyTensor):
            # metadata: answered from the wrapper, nothing to compute
            if name in ("size", "dim", "numel", "__len__", "ndimension", "is_contiguous", "__get__", "stride"):
                with torch._C.DisableTorchFunctionSubclass():
                    return func(*args, **kwargs)
            if a0._kind == "mask" and name == "unsqueeze" and (args[1] if len(args) > 1 else kwargs.get("dim")) in (-1, a0.dim()):
                return a0._derive(tuple(a0.shape) + (1,))
            if a0._kind == "lookup" and name in ("reshape", "view"):
                shape = args[1:] if not isinstance(args[1], (tuple, list, torch.Size)) else tuple(args[1])
                shape = tuple(int(x) for x in shape)
                if -1 not in shape and _numel(shape) == a0.numel() and shape[-1] == 4:
                    return a0._derive(shape)
            if a0._kind == "lookup" and name in ("mul", "__mul__", "multiply") and len(args) == 2 and _is_own_mask(a0, args[1]):
                return a0._derive(a0.shape, masked=True)
        if name in ("mul", "__rmul__", "__mul__") and len(args) == 2 and isinstance(args[1], LazyTensor) \
                and args[1]._kind == "lookup" and _is_own_mask(args[1], args[0]):
            return args[1]._derive(args[1].shape, masked=True)
        # anything else: become real tensors and carry on
        real_args = _materialize_tree(args)
        real_kwargs = _materialize_tree(kwargs)
        with torch._C.DisableTorchFunctionSubclass():
            return func(*real_args, **real_kwargs)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # an ATen op reached a handle without going through the Python API: make it real and run the op
        return func(*_materialize_tree(args), **_materialize_tree(kwargs or {}))

    def __repr__(self):
        return f"LazyTensor(kind={self._kind}, shape={tuple(self.shape)}, device={self.device})"


def _numel(shape):
    n = 1
    for x in shape:
        n *= x
    return n


def _is_own_mask(lookup: "LazyTensor", m) -> bool:
    """`m` is the in-bounds mask that belongs to this lookup, broadcast over the 4 channels."""
    return (isinstance(m, LazyTensor) and m._kind == "mask" and m._spec["ns"] is lookup._spec["ns"]
            and m._spec["grid"] is lookup._spec["grid"] and m.dim() == lookup.dim() and m.shape[-1] == 1
            and m.numel() * 4 == lookup.numel())


def _materialize_tree(x):
    if isinstance(x, LazyTensor):
        return x.materialize()
    if isinstance(x, (list, tuple)):
        return type(x)(_materialize_tree(v) for v in x)
    if isinstance(x, dict):
        return {k: _materialize_tree(v) for k, v in x.items()}
    return x


# ---------------------------------------------------------------------------------------------------- constructors
def lazy_samples(origins, dirs, rays_per_origin, num_samples, delta_step):
    n = dirs.shape[0]
    return LazyTensor((n * num_samples, 3), torch.float32, dirs.device, "samples",
                      dict(origins=origins, dirs=dirs, R=int(rays_per_origin), S=int(num_samples), delta=float(delta_step)))


def lazy_normalized(samples: LazyTensor, gmin, points_distance):
    return LazyTensor(samples.shape, torch.float32, samples.device, "normalized",
                      dict(samples=samples, gmin=tuple(float(g) for g in gmin), pd=float(points_distance)))


def lazy_lookup(ns: LazyTensor, grid):
    m = ns.shape[0]
    vals = LazyTensor((m, 4), torch.float32, ns.device, "lookup", dict(ns=ns, grid=grid, masked=False))
    mask = LazyTensor((m,), torch.bool, ns.device, "mask", dict(ns=ns, grid=grid))
    return vals, mask


def fused_composite(lookup: LazyTensor):
    """compute_alpha_weighted_pixels on a lazy, masked (C,R,S,4) lookup -> one fused march, or None if the handle does
    not have exactly that shape (the caller then materialises)."""
    if lookup._kind != "lookup" or not lookup._spec["masked"] or lookup._real is not None or lookup.dim() != 4:
        return None
    ns = lookup._spec["ns"]
    if ns._kind != "normalized" or ns._real is not None:
        return None
    smp = ns._spec["samples"]
    if smp._kind != "samples" or smp._real is not None:
        return None
    sp = smp._spec
    C_, R, S, _ = lookup.shape
    n_rays = sp["dirs"].shape[0]
    if S != sp["S"] or C_ * R != n_rays:
        return None
    grid = lookup._spec["grid"]
    # the grid handed to get_nearest_voxels is already clipped by the caller (scripts/train.py:146): no clamp here, the
    # gradient flows back through the caller's clip
    pix = ops.render_rays(grid, sp["origins"], sp["dirs"], S, sp["delta"], ns._spec["gmin"], ns._spec["pd"], mode="nearest",
                          clamp=False, rays_per_origin=sp["R"])
    return pix.reshape(C_, R, 4)
