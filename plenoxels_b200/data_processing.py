"""Drop-in for the reference's `src/data_processing.py` (NeRF-synthetic loader; host I/O, off the hot path)."""
from __future__ import annotations

import json

import numpy as np
import torch


def get_data_from_index(data, index):
    """(transform_matrix (4,4), rotation, file_path, camera_angle_x) of frame `index` — src/data_processing.py:7-15."""
    frame = data["frames"][index]
    return torch.tensor(frame["transform_matrix"]), frame["rotation"], frame["file_path"], data["camera_angle_x"]


def load_data(data):
    """(poses (C,4,4), file_paths, camera_angle_x) — src/data_processing.py:18-29."""
    poses, paths, fov = [], [], 0.0
    for i in range(len(data["frames"])):
        pose, _, path, fov = get_data_from_index(data, i)
        poses.append(pose.unsqueeze(0))
        paths.append(path)
    return torch.cat(poses, 0), paths, fov


def read_data(data_path):
    """src/data_processing.py:32-35."""
    with open(data_path, "r") as f:
        return json.load(f)


def _load_images(paths):
    from PIL import Image
    imgs = np.array([np.array(Image.open(p)) for p in paths])
    return torch.tensor(imgs, dtype=torch.float) / 255


def load_image_data(data_folder, object_folder, split="train"):
    """src/data_processing.py:38-48."""
    data = read_data(f"{data_folder}/{object_folder}/transforms_{split}.json")
    stem = f"{data_folder}/{object_folder}/{split}"
    return data, _load_images([f'{stem}/{fr["file_path"].split("/")[-1]}.png' for fr in data["frames"]])


def load_image_data_from_path(path, transformpath):
    """(transforms dict, imgs (C,H,W,4) fp32 in [0,1]) — src/data_processing.py:51-60."""
    data = read_data(transformpath)
    return data, _load_images([f'{path}/{fr["file_path"].split("/")[-1]}.png' for fr in data["frames"]])
