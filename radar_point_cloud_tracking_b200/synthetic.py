"""Seeded synthetic radar sweeps: test/bench input, not part of the path.

Counter-based, so the numpy generator here and the device generator in
``csrc/synth.cu`` (``rb_synth_echo``) produce the *same* echo tensor bit for bit from
the same :class:`SweepSpec` — the CPU oracle and the CUDA path always see identical
inputs without shipping gigabytes of fixtures.

Data model (mirrors the radar CSV of the reference, ``PIPELINE_DOCUMENTATION.txt:41-50``:
``Status,Scale,Range,Gain,Angle,Echo_0..Echo_1023`` — one row per spoke):

* ``echo[F, G, S, E]`` float32, integer valued 0..255;
* ``angle_units[S]`` raw angle codes (0..8195), ``scale[S]`` max range in metres;
* background noise ``U{0..9}`` everywhere, sparse clutter ``U{11..255}`` with a per-gain
  probability, plus painted rectangles (in spoke x bin space): persistent land blobs,
  flickering static buoys and linearly moving boats.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Sequence, Tuple

import numpy as np

GAINS_DEFAULT: Tuple[int, ...] = (40, 50, 75)
ANGLE_UNITS_PER_TURN = 8196  # reference: ANGLE_SCALE = 360/8196 (T4:66)

_M1 = np.uint32(0x85EBCA6B)
_M2 = np.uint32(0xC2B2AE35)
_GOLD = np.uint32(0x9E3779B9)
_SALT_CLUTTER = np.uint32(0x6A09E667)
_SALT_VALUE = np.uint32(0xBB67AE85)
_SALT_RECT = np.uint32(0x3C6EF372)

#: per-gain multiplier of the clutter probability and of painted-object intensity
GAIN_CLUTTER = {40: 0.6, 50: 1.0, 70: 1.4, 75: 1.6}
GAIN_LEVEL = {40: 0.70, 50: 0.85, 70: 0.95, 75: 1.0}


def mix32(x: np.ndarray) -> np.ndarray:
    """murmur3 fmix32 on a uint32 array (wrapping arithmetic); same as ``mix32`` in synth.cu."""
    x = np.asarray(x, dtype=np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= _M1
    x ^= x >> np.uint32(13)
    x *= _M2
    x ^= x >> np.uint32(16)
    return x


def _mix32_scalar(v: int) -> int:
    v &= 0xFFFFFFFF
    v ^= v >> 16
    v = (v * 0x85EBCA6B) & 0xFFFFFFFF
    v ^= v >> 13
    v = (v * 0xC2B2AE35) & 0xFFFFFFFF
    v ^= v >> 16
    return v


@dataclass
class SweepSpec:
    """Everything that determines a synthetic data set."""
    seed: int = 1234
    frames: int = 16
    spokes: int = 2048
    bins: int = 1024
    gains: Tuple[int, ...] = GAINS_DEFAULT
    scale_m: float = 231.5            # 0.125 NM data set
    clutter_p: float = 0.003          # at gain 50
    land_blobs: int = 3
    buoys: int = 4
    boats: int = 3
    land_presence: float = 0.95
    buoy_presence: float = 0.9

    @property
    def sweeps(self) -> int:
        return self.frames * len(self.gains)

    def angle_units(self) -> np.ndarray:
        s = np.arange(self.spokes, dtype=np.int64)
        return ((s * ANGLE_UNITS_PER_TURN) // self.spokes).astype(np.int32)

    def scale(self) -> np.ndarray:
        return np.full(self.spokes, self.scale_m, dtype=np.float32)

    def clutter_threshold(self, gain: int) -> int:
        """uint32 threshold: a cell is clutter when its hash < threshold."""
        p = min(max(self.clutter_p * GAIN_CLUTTER.get(gain, 1.0), 0.0), 1.0)
        return min(int(p * 4294967296.0), 0xFFFFFFFF)

    def sweep_key(self, frame: int, gain_index: int) -> int:
        w = frame * len(self.gains) + gain_index
        return _mix32_scalar((self.seed & 0xFFFFFFFF) ^ _mix32_scalar((w + 0x9E3779B9) & 0xFFFFFFFF))


@dataclass
class Rect:
    """Painted rectangle, half-open in spoke and bin: value = base + hash % span."""
    frame: int
    s0: int
    s1: int
    j0: int
    j1: int
    base: int
    span: int


def build_rects(spec: SweepSpec) -> List[Rect]:
    """Host-side deterministic object list (land / buoys / boats) for every frame."""
    rng = np.random.default_rng(spec.seed + 7919)
    S, E, F = spec.spokes, spec.bins, spec.frames
    res = spec.scale_m / E
    rects: List[Rect] = []

    def clip_rect(f, sc, jc, hs, hj, base, span):
        s0, s1 = max(sc - hs, 0), min(sc + hs + 1, S)
        j0, j1 = max(jc - hj, 1), min(jc + hj + 1, E)
        if s1 > s0 and j1 > j0:
            rects.append(Rect(f, s0, s1, j0, j1, base, span))

    for _ in range(spec.land_blobs):
        sc = int(rng.integers(S // 16, S - S // 16))
        jc = int(rng.integers(E // 3, E - E // 8))
        hs = int(rng.integers(max(S // 128, 1), max(S // 48, 2)))
        hj = int(rng.integers(max(E // 64, 1), max(E // 24, 2)))
        present = rng.random(F) < spec.land_presence
        for f in range(F):
            if present[f]:
                clip_rect(f, sc, jc, hs, hj, 190, 60)
    for _ in range(spec.buoys):
        sc = int(rng.integers(0, S))
        jc = int(rng.integers(E // 8, E - E // 8))
        present = rng.random(F) < spec.buoy_presence
        for f in range(F):
            if present[f]:
                clip_rect(f, sc, jc, max(S // 512, 1), 3, 150, 90)
    for _ in range(spec.boats):
        r0 = float(rng.uniform(0.25, 0.8)) * spec.scale_m
        th0 = float(rng.uniform(0, 2 * np.pi))
        x0, y0 = r0 * np.cos(th0), r0 * np.sin(th0)
        speed = float(rng.uniform(1.0, 3.0))
        hd = float(rng.uniform(0, 2 * np.pi))
        vx, vy = speed * np.cos(hd), speed * np.sin(hd)
        for f in range(F):
            x, y = x0 + vx * f, y0 + vy * f
            r = float(np.hypot(x, y))
            th = float(np.arctan2(y, x)) % (2 * np.pi)
            sc = int(round(th / (2 * np.pi) * S)) % S
            jc = int(round(r / res))
            if jc < E - 4:
                clip_rect(f, sc, jc, max(S // 400, 2), 4, 140, 100)
    return rects


MAX_GAINS = 4
RECT_COLS = 6 + MAX_GAINS


def rects_to_array(spec: SweepSpec, rects: Sequence[Rect]) -> Tuple[np.ndarray, np.ndarray]:
    """Device painter tables: ``int32[n, 10]`` = (frame, s0, s1, j0, j1, span, base@gain0..3),
    sorted by frame (stable, so "last rectangle wins" is preserved), and ``int32[F+1]`` offsets."""
    assert len(spec.gains) <= MAX_GAINS
    order = sorted(range(len(rects)), key=lambda i: rects[i].frame)
    arr = np.zeros((len(rects), RECT_COLS), dtype=np.int32)
    for k, i in enumerate(order):
        r = rects[i]
        arr[k, :6] = (r.frame, r.s0, r.s1, r.j0, r.j1, r.span)
        for gi, gain in enumerate(spec.gains):
            arr[k, 6 + gi] = int(r.base * GAIN_LEVEL.get(gain, 1.0))
    offs = np.zeros(spec.frames + 1, dtype=np.int32)
    if len(rects):
        np.add.at(offs, arr[:, 0] + 1, 1)
    return arr, np.cumsum(offs).astype(np.int32)


def _cell_hash(key: int, s0: int, s1: int, bins: int) -> np.ndarray:
    s = np.arange(s0, s1, dtype=np.uint32)[:, None]
    j = np.arange(bins, dtype=np.uint32)[None, :]
    cell = s * np.uint32(bins) + j
    return mix32(np.uint32(key) ^ (cell * _GOLD))


def synth_sweep(spec: SweepSpec, frame: int, gain_index: int,
                rects: Sequence[Rect] | None = None) -> np.ndarray:
    """One ``[S, E]`` float32 echo matrix."""
    S, E = spec.spokes, spec.bins
    gain = spec.gains[gain_index]
    key = spec.sweep_key(frame, gain_index)
    h1 = _cell_hash(key, 0, S, E)
    h2 = mix32(h1 + _SALT_CLUTTER)
    h3 = mix32(h2 ^ _SALT_VALUE)
    val = (h1 % np.uint32(10)).astype(np.int32)
    clutter = h2 < np.uint32(spec.clutter_threshold(gain))
    val = np.where(clutter, 11 + (h3 % np.uint32(245)).astype(np.int32), val)
    if rects is None:
        rects = [r for r in build_rects(spec) if r.frame == frame]
    level = GAIN_LEVEL.get(gain, 1.0)
    for r in rects:
        if r.frame != frame:
            continue
        hr = mix32(h1[r.s0:r.s1, r.j0:r.j1] ^ _SALT_RECT)
        base = int(r.base * level)
        val[r.s0:r.s1, r.j0:r.j1] = np.minimum(base + (hr % np.uint32(r.span)).astype(np.int32), 255)
    return val.astype(np.float32)


def synth_echo(spec: SweepSpec) -> np.ndarray:
    """``[F, G, S, E]`` float32 (host). Use the device generator for big batches."""
    rects = build_rects(spec)
    by_frame: Dict[int, List[Rect]] = {}
    for r in rects:
        by_frame.setdefault(r.frame, []).append(r)
    out = np.empty((spec.frames, len(spec.gains), spec.spokes, spec.bins), dtype=np.float32)
    for f in range(spec.frames):
        for g in range(len(spec.gains)):
            out[f, g] = synth_sweep(spec, f, g, by_frame.get(f, []))
    return out


def write_csv_tree(spec: SweepSpec, root: Path, echo: np.ndarray | None = None,
                   start: str = "20250813_142602", period_ms: int = 2500) -> List[Dict[int, Path]]:
    """Write ``gain_XX/YYYYMMDD_HHMMSS_mmm.csv`` files the reference CLI can discover
    (layout per T4:235-309). Returns the per-frame ``{gain: path}`` mapping."""
    from datetime import datetime, timedelta

    if echo is None:
        echo = synth_echo(spec)
    t0 = datetime.strptime(start, "%Y%m%d_%H%M%S")
    header = "Status,Scale,Range,Gain,Angle," + ",".join(f"Echo_{i}" for i in range(spec.bins))
    ang = spec.angle_units()
    scale = spec.scale()
    frames: List[Dict[int, Path]] = []
    for f in range(spec.frames):
        entry: Dict[int, Path] = {}
        for gi, gain in enumerate(spec.gains):
            ts = t0 + timedelta(milliseconds=f * period_ms + gi * 100)
            name = ts.strftime("%Y%m%d_%H%M%S") + f"_{ts.microsecond // 1000:03d}.csv"
            d = root / f"gain_{gain}"
            d.mkdir(parents=True, exist_ok=True)
            rows = [header]
            e = echo[f, gi].astype(np.int64)
            for s in range(spec.spokes):
                rows.append(f"1,{scale[s]:g},3,{gain},{int(ang[s])}," + ",".join(map(str, e[s].tolist())))
            (d / name).write_text("\n".join(rows) + "\n")
            entry[gain] = d / name
        frames.append(entry)
    return frames
