"""Flat-label ST-DBSCAN: mirror of ``radar_pipeline.processors.clustering.st_dbscan``
(PKG clustering.py:49-115) and of ``3_stdbscan_point_clouds.py::st_dbscan`` (T3:101-136).

Same signature and return value (``labels[N]`` int32, -1 = noise, clusters numbered exactly as the
reference numbers them). Runs on the GPU through ``rb_stdbscan``; there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import device as dev
from ._lib import RadarB200Error


def _as_float32_exact(a: np.ndarray, what: str) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype == np.float32:
        return a
    b = a.astype(np.float32)
    if not np.array_equal(b.astype(np.float64), a.astype(np.float64)):
        raise RadarB200Error(f"{what}: values are not representable in float32; the CUDA path works on the "
                             "float32 data the reference pipeline produces")
    return b


def st_dbscan(coords: np.ndarray, times: np.ndarray, eps_space: float, eps_time: float, min_samples: int,
              min_frames: int = None) -> np.ndarray:
    """Spatio-temporal DBSCAN: neighbours are within ``eps_space`` (Euclidean, float64 test on the
    float32 coordinates, inclusive) AND within ``eps_time``. ``coords`` is ``[N, D]`` with D in 1..3.

    ``min_frames=None``: the function of ``3_stdbscan_point_clouds.py:101`` / ``radar_pipeline.processors.clustering``.
    ``min_frames=k``: the ``PointCloudWorkF/stdbscan_denoising_pipeline.py:264`` variant (core points must also have
    neighbours in at least k distinct frames; its FIFO border rule), labels identical to that function's."""
    coords = np.asarray(coords)
    if coords.ndim == 1:
        coords = coords.reshape(-1, 1)
    n, dim = coords.shape
    if n == 0:
        return np.full(0, -1, dtype=np.int32)
    if not 1 <= dim <= 3:
        raise RadarB200Error("coords must have 1 to 3 columns")
    if not torch.cuda.is_available():
        raise RadarB200Error("no CUDA device: st_dbscan is GPU only (no CPU fallback)")
    coords = np.ascontiguousarray(_as_float32_exact(coords, "coords"))
    times = np.asarray(times)
    if times.dtype != np.float32:
        # float64/int times: only integer-valued ones give identical arithmetic in float32
        t32 = _as_float32_exact(times, "times")
        if not np.array_equal(t32, np.rint(t32)) or np.abs(t32).max(initial=0) >= 2 ** 24:
            raise RadarB200Error("times must be float32 (or integer valued)")
        times = t32
    d = torch.device("cuda", torch.cuda.current_device())
    flat = torch.from_numpy(coords).to(d).view(-1)
    t = torch.from_numpy(np.ascontiguousarray(times)).to(d)
    labels, _ = dev.stdbscan(flat, flat[1:] if dim > 1 else None, flat[2:] if dim > 2 else None, t,
                             eps_space, eps_time, min_samples, stride=dim, n=n, min_frames=min_frames)
    return labels.cpu().numpy()
