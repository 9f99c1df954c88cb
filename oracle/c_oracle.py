"""ctypes front end of ``stdbscan_ref.c`` — test infrastructure only."""
from __future__ import annotations

import ctypes
from typing import Tuple

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(str(_build.build()))
        lib.oracle_stdbscan.restype = ctypes.c_int64
        lib.oracle_stdbscan.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_double, ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib = lib
    return _lib


def st_dbscan_c(coords: np.ndarray, times: np.ndarray, eps_space: float, eps_time: float,
                min_samples: int) -> Tuple[np.ndarray, np.ndarray]:
    """Canonical ST-DBSCAN labels + core flags from the C oracle."""
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    if coords.ndim == 1:
        coords = coords.reshape(-1, 1)
    times = np.ascontiguousarray(times, dtype=np.float32)
    n, dim = coords.shape
    labels = np.empty(n, dtype=np.int32)
    core = np.empty(n, dtype=np.uint8)
    rc = _load().oracle_stdbscan(coords.ctypes.data, dim, times.ctypes.data, n,
                                 float(eps_space), float(np.float32(eps_time)),
                                 int(min_samples), labels.ctypes.data, core.ctypes.data)
    if rc < 0:
        raise RuntimeError("oracle_stdbscan failed")
    return labels, core.astype(bool)


def st_dbscan_wf_c(coords: np.ndarray, times: np.ndarray, eps_space: float, eps_time: float,
                   min_samples: int, min_frames: int = 2) -> Tuple[np.ndarray, np.ndarray]:
    """Labels + core flags of the ``PointCloudWorkF`` variant (``min_frames`` core test, FIFO border rule)."""
    lib = _load()
    if not hasattr(lib, "_wf_bound"):
        lib.oracle_stdbscan_wf.restype = ctypes.c_int64
        lib.oracle_stdbscan_wf.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_double, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib._wf_bound = True
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    if coords.ndim == 1:
        coords = coords.reshape(-1, 1)
    times = np.ascontiguousarray(times, dtype=np.float32)
    n, dim = coords.shape
    labels = np.empty(n, dtype=np.int32)
    core = np.empty(n, dtype=np.uint8)
    rc = lib.oracle_stdbscan_wf(coords.ctypes.data, dim, times.ctypes.data, n, float(eps_space), float(np.float32(eps_time)),
                                int(min_samples), int(min_frames), labels.ctypes.data, core.ctypes.data)
    if rc < 0:
        raise RuntimeError("oracle_stdbscan_wf failed")
    return labels, core.astype(bool)
