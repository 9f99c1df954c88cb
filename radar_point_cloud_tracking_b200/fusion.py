"""Multi-gain fusion modes of ``5_gain_fusion_ply_builder.py`` on the GPU.

``fuse_gains_absolute`` (T5:193-219) is the concatenation the tracker uses (see ``tracker.build_frame``);
``fuse_gains_max`` (T5:222-273) pools all gains on a grid and keeps the maximum intensity per cell.
``fuse_points_*`` take already-converted per-gain points so they can be fed from files or from device batches;
``fuse_gains_*`` are the reference's file-level functions (same names and arguments), and :func:`install` swaps them
- with the loader, the colour helpers and the PLY writers - into an imported ``5_gain_fusion_ply_builder`` module."""
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


# ---- the reference's file-level surface (T5) -------------------------------------------------------------------
# T5's own configuration globals (T5:52-63); a module passed to :func:`install` keeps using ITS values, read at call time
NUM_ECHO_COLUMNS = 1024
INTENSITY_THRESHOLD = 5.0
POINT_STRIDE = 8
_CFG_NAMES = ("NUM_ECHO_COLUMNS", "INTENSITY_THRESHOLD", "POINT_STRIDE")


def _config(module=None):
    import sys
    from types import SimpleNamespace

    m = module or sys.modules[__name__]
    return SimpleNamespace(**{k: getattr(m, k) for k in _CFG_NAMES})


def load_radar_csv(path, _cfg=None):
    """``(x, y, intensity, gain)`` of one sweep file (T5:84-121 - the tracker's loader with T5's threshold / stride)."""
    from . import tracker

    return tracker.load_radar_csv(path, _cfg=_cfg or _config())


def _load_all(frame_files, cfg) -> Dict[int, Points]:
    per_gain = {}
    for gain, path in sorted(frame_files.items()):
        x, y, z, _ = load_radar_csv(path, _cfg=cfg)
        per_gain[gain] = (x, y, z)                    # labelled with the directory's gain, not the file's (T5:202,209)
    return per_gain


def fuse_gains_absolute(frame_files, _cfg=None):
    """``(x, y, intensity, gain_labels)`` of all gains of a frame, ascending gain order (T5:193-219)."""
    return fuse_points_absolute(_load_all(frame_files, _cfg or _config()))


def fuse_gains_max(frame_files, grid_resolution: float = 1.0, _cfg=None) -> Points:
    """``(x, y, max_intensity)``: all gains pooled on a grid, maximum per occupied cell (T5:222-273)."""
    return fuse_points_max(_load_all(frame_files, _cfg or _config()), grid_resolution)


def install(ref_module) -> None:
    """Replace the hot and the output functions of an imported ``5_gain_fusion_ply_builder`` module: the loader and
    both fusion modes (GPU), the colour helpers and the PLY writers (byte-identical, native formatter). File
    discovery, grouping, plots and the CLI stay the module's own; its configuration globals are read at call time."""
    from . import plyio

    cfg = lambda: _config(ref_module)
    ref_module.load_radar_csv = lambda path: load_radar_csv(path, _cfg=cfg())
    ref_module.fuse_gains_absolute = lambda frame_files: fuse_gains_absolute(frame_files, _cfg=cfg())
    ref_module.fuse_gains_max = lambda frame_files, grid_resolution=1.0: fuse_gains_max(frame_files, grid_resolution, _cfg=cfg())
    for name in ("normalize_intensity", "intensity_to_rgb", "gain_to_rgb", "write_ply", "write_ply_fast"):
        setattr(ref_module, name, getattr(plyio, name))
