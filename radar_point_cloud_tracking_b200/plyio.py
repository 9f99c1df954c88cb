"""PLY output (SURVEY.md section 8 f, rank 4): the reference's writers and colour helpers with the same names,
arguments and bytes on disk.

* :func:`write_ply_fast` / :func:`write_ply` — ``5_gain_fusion_ply_builder.py:370-403`` / ``:345-367`` (ASCII);
  the body is written by the library's host function ``rb_ply_append_ascii`` (exact ``%.4f``, see ``csrc/plyfmt.cu``)
  instead of a float64 table formatted row by row in Python;
* :func:`write_ply_binary` — the binary branch of ``PointCloudWorkF/stdbscan_denoising_pipeline.py:797-827``;
* :func:`normalize_intensity`, :func:`intensity_to_rgb`, :func:`gain_to_rgb` — ``T5:276-342``.

Coordinates / colours may be numpy arrays or torch tensors (device tensors are brought back once).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Dict, Tuple

import numpy as np

from . import _lib

GAIN_COLORS: Dict[int, Tuple[int, int, int]] = {40: (0, 114, 255), 50: (0, 200, 83), 70: (255, 165, 0), 75: (255, 87, 34)}   # T5:45-50
INTENSITY_PERCENTILE = 99                                                                                   # T5:63

_HEADER = ("ply\nformat {fmt} 1.0\nelement vertex {n}\nproperty float x\nproperty float y\nproperty float z\n"
           "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")


def _host(a) -> np.ndarray:
    if hasattr(a, "detach"):                      # torch tensor, any device
        a = a.detach().cpu().numpy()
    return np.asarray(a)


# ---- colours ------------------------------------------------------------------------------------------
def normalize_intensity(intensity) -> np.ndarray:
    """Intensities scaled to 0..255 between their minimum and their 99th percentile, clipped (T5:276-289)."""
    v = _host(intensity)
    if len(v) == 0:
        return v
    top, low = np.percentile(v, INTENSITY_PERCENTILE), np.min(v)
    if top <= low:
        return np.zeros_like(v)
    return np.clip((v - low) / (top - low) * 255.0, 0, 255)


def intensity_to_rgb(intensity) -> np.ndarray:
    """Blue -> cyan -> green -> yellow -> red over four quarters of 0..255 (T5:292-327): in every quarter one channel
    ramps with ``t = 4 * (v/255 - quarter start)``, the others sit at 0 or 255; ramps are truncated to uint8."""
    u = _host(intensity) / 255.0
    rgb = np.zeros((len(u), 3), dtype=np.uint8)
    quarter = [u < 0.25, (u >= 0.25) & (u < 0.5), (u >= 0.5) & (u < 0.75), u >= 0.75]
    ramp = lambda sel, start: (u[sel] - start) * 4 if start else u[sel] * 4
    q = quarter[0]
    rgb[q, 1] = (ramp(q, 0) * 255).astype(np.uint8); rgb[q, 2] = 255
    q = quarter[1]
    rgb[q, 1] = 255; rgb[q, 2] = ((1 - ramp(q, 0.25)) * 255).astype(np.uint8)
    q = quarter[2]
    rgb[q, 0] = (ramp(q, 0.5) * 255).astype(np.uint8); rgb[q, 1] = 255
    q = quarter[3]
    rgb[q, 0] = 255; rgb[q, 1] = ((1 - ramp(q, 0.75)) * 255).astype(np.uint8)
    return rgb


def gain_to_rgb(gains) -> np.ndarray:
    """One fixed colour per gain label, black for anything else (T5:330-342)."""
    g = _host(gains)
    rgb = np.zeros((len(g), 3), dtype=np.uint8)
    for gain, colour in GAIN_COLORS.items():
        rgb[g == gain] = colour
    return rgb


# ---- writers ------------------------------------------------------------------------------------------
def _append_ascii(path: Path, x: np.ndarray, y: np.ndarray, z: np.ndarray, colors: np.ndarray) -> None:
    """Vertex lines ``%.4f %.4f %.4f %d %d %d``: native for float32 coordinates + uint8 colours, numpy otherwise."""
    if all(a.dtype == np.float32 for a in (x, y, z)) and colors.dtype == np.uint8 and colors.ndim == 2 and colors.shape[1] == 3:
        lib = _lib.load()
        arrs = [np.ascontiguousarray(a) for a in (x, y, z, colors)]
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = lib.rb_ply_append_ascii(str(path).encode(), ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]), ptr(arrs[3]), len(x))
        if rc != 0:
            raise _lib.RadarB200Error(f"rb_ply_append_ascii: {lib.rb_last_error().decode()}")
        return
    data = np.column_stack([x, y, z, colors[:, 0].astype(int), colors[:, 1].astype(int), colors[:, 2].astype(int)])
    with path.open("a", encoding="utf-8") as fh:
        np.savetxt(fh, data, fmt="%.4f %.4f %.4f %d %d %d")


def write_ply_fast(path: Path, x, y, z, colors) -> None:
    """ASCII PLY, coordinates cast to float32 first (T5:370-403)."""
    path = Path(path)
    x, y, z = (_host(a).astype(np.float32) for a in (x, y, z))
    colors = _host(colors)
    with path.open("w", encoding="utf-8") as fh:
        fh.write(_HEADER.format(fmt="ascii", n=len(x)))
    _append_ascii(path, x, y, z, colors)
    print(f"  Wrote {len(x):,} points to {path.name}")


def write_ply(path: Path, x, y, z, colors) -> None:
    """ASCII PLY, every value formatted as it is (T5:345-367: ``f"{xp:.4f}"`` formats the value's own precision, so
    float64 input is NOT rounded to float32 first - only float32 input takes the native writer)."""
    path = Path(path)
    x, y, z, colors = _host(x), _host(y), _host(z), _host(colors)
    with path.open("w", encoding="utf-8") as fh:
        fh.write(_HEADER.format(fmt="ascii", n=len(x)))
    if all(a.dtype == np.float32 for a in (x, y, z)) and colors.dtype == np.uint8:
        _append_ascii(path, x, y, z, colors)
    else:
        with path.open("a", encoding="utf-8") as fh:
            for xp, yp, zp, (r, g, b) in zip(x, y, z, colors):
                fh.write(f"{xp:.4f} {yp:.4f} {zp:.4f} {r} {g} {b}\n")
    print(f"  Wrote {len(x):,} points to {path.name}")


def write_ply_binary(path: Path, x, y, z, colors) -> None:
    """Binary little-endian PLY: 15-byte records ``<f4 x, y, z; u1 r, g, b`` (WF:797-827). Colours are the caller's
    (the reference derives them from cluster labels / intensities with matplotlib colour maps)."""
    path = Path(path)
    x, y, z, colors = _host(x), _host(y), _host(z), _host(colors)
    rec = np.empty(len(x), dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1")]))
    rec["x"], rec["y"], rec["z"] = x.astype(np.float32), y.astype(np.float32), z.astype(np.float32)
    rec["r"], rec["g"], rec["b"] = colors[:, 0], colors[:, 1], colors[:, 2]
    with path.open("wb") as fh:
        fh.write(_HEADER.format(fmt="binary_little_endian", n=len(x)).encode("ascii"))
        rec.tofile(fh)
