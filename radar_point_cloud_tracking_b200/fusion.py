"""Multi-gain fusion modes of ``5_gain_fusion_ply_builder.py`` on the GPU.

``fuse_gains_absolute`` (T5:193-219) is the concatenation the tracker uses (see ``tracker.build_frame``);
``fuse_gains_max`` (T5:222-273) pools all gains on a grid and keeps the maximum intensity per cell.
Both take already-converted per-gain points so they can be fed from files or from device batches."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from . import device as dev
from ._lib import RadarB200Error

Points = Tuple[np.ndarray, np.ndarray, np.ndarray]


def fuse_points_absolute(per_gain: Dict[int, Points]):
    """Concatenate in ascending gain order with an int32 gain label per point (T5:202-219)."""
    xs, ys, zs, gs = [], [], [], []
    for gain in sorted(per_gain):
        x, y, z = per_gain[gain]
        if len(x) == 0:
            continue
        xs.append(x), ys.append(y), zs.append(z), gs.append(np.full(len(x), gain, dtype=np.int32))
    if not xs:
        return np.array([]), np.array([]), np.array([]), np.array([])
    return np.concatenate(xs), np.concatenate(ys), np.concatenate(zs), np.concatenate(gs)


def fuse_points_max(per_gain: Dict[int, Points], grid_resolution: float = 1.0) -> Points:
    """Grid max-pooling of all gains (T5:222-273): returns ``(x, y, max_intensity)`` — cell centres
    (float64) of the occupied cells in y-major order and their float32 maximum intensity."""
    if not torch.cuda.is_available():
        raise RadarB200Error("no CUDA device: fuse_points_max is GPU only (no CPU fallback)")
    parts = [per_gain[g] for g in sorted(per_gain) if len(per_gain[g][0])]
    if not parts:
        return np.array([]), np.array([]), np.array([])
    d = torch.device("cuda", torch.cuda.current_device())
    x = torch.from_numpy(np.concatenate([p[0] for p in parts]).astype(np.float32, copy=False)).to(d)
    y = torch.from_numpy(np.concatenate([p[1] for p in parts]).astype(np.float32, copy=False)).to(d)
    z = torch.from_numpy(np.concatenate([p[2] for p in parts]).astype(np.float32, copy=False)).to(d)
    b4 = dev.bounds(x, y).cpu().numpy()
    x_min, x_max, y_min, y_max = (np.float32(v) for v in b4)
    nx = int(np.ceil((x_max - x_min) / grid_resolution)) + 1            # T5:255-256 (float32 arithmetic)
    ny = int(np.ceil((y_max - y_min) / grid_resolution)) + 1
    ix, iy, mx = dev.fuse_max_cells(x, y, z, float(x_min), float(y_min), float(np.float32(grid_resolution)), nx, ny)
    cx, cy = ix.cpu().numpy().astype(np.int64), iy.cpu().numpy().astype(np.int64)
    out_x = x_min + cx * grid_resolution + grid_resolution / 2          # T5:269-270 (float64 on the host)
    out_y = y_min + cy * grid_resolution + grid_resolution / 2
    return out_x, out_y, mx.cpu().numpy()
