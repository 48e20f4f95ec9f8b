"""Drop-in for the reference's `src/data_processing.py`: the NeRF-synthetic loader (host I/O, runs once, off the hot path).

A dataset is a `transforms_*.json` (one `camera_angle_x`, a list of frames with `file_path`, `rotation` and a 4x4
camera-to-world `transform_matrix`) next to a folder of RGBA PNGs named after the last component of each `file_path`.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch


def read_data(data_path):
    """Parsed transforms json — src/data_processing.py:32-35."""
    return json.loads(Path(data_path).read_text())


def get_data_from_index(data, index):
    """(pose (4,4) fp32, rotation, file_path, camera_angle_x) of one frame — src/data_processing.py:7-15."""
    frame = data["frames"][index]
    pose = torch.tensor(frame["transform_matrix"])
    return pose, frame["rotation"], frame["file_path"], data["camera_angle_x"]


def load_data(data):
    """(poses (C,4,4), file paths, camera_angle_x) — src/data_processing.py:18-29 (0.0 for an empty frame list)."""
    frames = data["frames"]
    if not frames:
        return torch.cat([], 0), [], 0.0
    poses = torch.stack([torch.tensor(f["transform_matrix"]) for f in frames], 0)
    return poses, [f["file_path"] for f in frames], data["camera_angle_x"]


def _png_bytes(folder, frames) -> np.ndarray:
    """(C,H,W,4) uint8 as decoded from `<folder>/<basename of file_path>.png`."""
    from PIL import Image
    names = [Path(folder) / (f["file_path"].split("/")[-1] + ".png") for f in frames]
    return np.stack([np.asarray(Image.open(n)) for n in names], 0)


def _png_stack(folder, frames) -> torch.Tensor:
    """(C,H,W,4) fp32 in [0,1] — the reference's conversion, src/data_processing.py:58."""
    return torch.tensor(_png_bytes(folder, frames), dtype=torch.float) / 255


def load_image_data_from_path(path, transformpath):
    """(transforms dict, imgs (C,H,W,4) fp32) — src/data_processing.py:51-60."""
    data = read_data(transformpath)
    return data, _png_stack(path, data["frames"])


def load_image_bytes_from_path(path, transformpath):
    """(transforms dict, imgs (C,H,W,4) uint8): the same dataset with the pixels left as the PNGs store them.  The training
    kernels convert a target pixel when they fetch it (fp32(u8) / 255, bit-equal to the line above), so the resident image set
    is a quarter of the size (100 views of 800x800: 256 MB instead of 1.02 GB) and the host never touches a float."""
    data = read_data(transformpath)
    px = _png_bytes(path, data["frames"])
    if px.dtype != np.uint8 or px.ndim != 4 or px.shape[3] != 4:
        raise ValueError(f"expected 8-bit RGBA PNGs, got {px.dtype} {px.shape}")
    return data, torch.from_numpy(np.ascontiguousarray(px))


def load_image_data(data_folder, object_folder, split="train"):
    """Same, addressed as <data_folder>/<object_folder>/{transforms_<split>.json, <split>/} — src/data_processing.py:38-48."""
    root = Path(data_folder) / object_folder
    return load_image_data_from_path(root / split, root / f"transforms_{split}.json")
