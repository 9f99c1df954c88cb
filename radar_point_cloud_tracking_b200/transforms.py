"""Mirror of ``radar_pipeline.core.transforms.{polar_to_cartesian, sweep_to_point_cloud}``
(PKG transforms.py:13-79) on the GPU. Same signatures; ``sweep`` / ``config`` are duck-typed
(``RadarSweep`` / ``ProcessingConfig`` of the reference package work as they are)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import device as dev
from ._lib import RadarB200Error


@dataclass
class PointCloud:
    """Same fields as the reference's ``PointCloud`` (PKG loaders.py:28-43)."""
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    colors: Optional[np.ndarray] = None

    @property
    def size(self) -> int:
        return self.x.size

    def to_coords(self) -> np.ndarray:
        return np.column_stack((self.x, self.y, self.z))


def _cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RadarB200Error("no CUDA device: the radar-b200 detection path is GPU only (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _trig(angles_rad: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    col = np.asarray(angles_rad)[:, None]                 # the reference evaluates the column view
    return (np.ascontiguousarray(np.cos(col)[:, 0], dtype=np.float32),
            np.ascontiguousarray(np.sin(col)[:, 0], dtype=np.float32))


def polar_to_cartesian(angles_rad: np.ndarray, ranges: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """``x = ranges * cos(angles)[:, None]``, ``y = ranges * sin(angles)[:, None]`` (float32)."""
    angles_rad = np.asarray(angles_rad)
    ranges = np.asarray(ranges)
    if angles_rad.dtype != np.float32 or ranges.dtype != np.float32:
        raise RadarB200Error("polar_to_cartesian: the CUDA path takes float32 angles and ranges")
    d = _cuda()
    c, s = _trig(angles_rad)
    x, y = dev.polar_to_cartesian(torch.from_numpy(np.ascontiguousarray(ranges)).to(d),
                                  torch.from_numpy(c).to(d), torch.from_numpy(s).to(d))
    return x.cpu().numpy(), y.cpu().numpy()


def sweep_to_point_cloud(sweep, config=None, radar_config=None) -> PointCloud:
    """Threshold (strict ``>``), row-major compaction and stride of one sweep (PKG transforms.py:37-79).
    Defaults as ``ProcessingConfig``: threshold 0.0, stride 16 (PKG config/models.py:43-44)."""
    thr = float(getattr(config, "intensity_threshold", 0.0)) if config is not None else 0.0
    stride = int(getattr(config, "point_stride", 16)) if config is not None else 16
    d = _cuda()
    echo = np.ascontiguousarray(sweep.intensities, dtype=np.float32)
    ranges = np.ascontiguousarray(sweep.ranges, dtype=np.float32)
    if np.asarray(sweep.intensities).dtype != np.float32 or np.asarray(sweep.ranges).dtype != np.float32:
        raise RadarB200Error("sweep_to_point_cloud: the CUDA path takes float32 sweeps")
    c, s = _trig(np.asarray(sweep.angles_rad))
    gain = torch.tensor([int(sweep.gain or 0)], dtype=torch.int32, device=d)
    batch = dev.spoke_to_points(torch.from_numpy(echo).to(d)[None], torch.from_numpy(c).to(d),
                                torch.from_numpy(s).to(d), None, gain, thr, stride,
                                ranges=torch.from_numpy(ranges).to(d)[None])
    n = batch.n
    return PointCloud(x=batch.x[:n].cpu().numpy(), y=batch.y[:n].cpu().numpy(), z=batch.inten[:n].cpu().numpy())
