"""Golden files of the PLY writers and colour helpers, produced by the UNMODIFIED reference functions of
``PointCloudWork/5_gain_fusion_ply_builder.py`` (T5:276-403), imported by path in the build container
(``/root/reference`` does not exist on the GPU box).

    python tests/golden/make_golden_ply.py        ->  tests/golden/ply_golden.npz
"""
import contextlib
import importlib.util
import io
import tempfile
from pathlib import Path

import numpy as np

REF = "/root/reference/PointCloudWork/5_gain_fusion_ply_builder.py"
spec = importlib.util.spec_from_file_location("t5_ref", REF)
t5 = importlib.util.module_from_spec(spec)
spec.loader.exec_module(t5)


def cloud(rng, n):
    """float32 coordinates with the awkward values in: exact %.4f ties, negatives that round to zero, -0.0, big."""
    x = (rng.normal(0, 800, n)).astype(np.float32)
    y = (rng.normal(0, 800, n)).astype(np.float32)
    z = (rng.random(n) * 255).astype(np.float32)
    special = np.array([0.03125, -0.03125, 0.00005, -0.00001, -0.0, 0.0, 1234567.875, -98765.4321, 1e-5, 2.5e-5, 0.99995, 16777216.0,
                        3.0e9, -7.5e12], dtype=np.float32)
    x[:len(special)] = special
    y[:len(special)] = special[::-1]
    x[20:60] = (rng.integers(-4000, 4000, 40) / 32.0 + 1 / 65536).astype(np.float32)      # near-ties
    y[20:60] = (rng.integers(-4000, 4000, 40) / 32768.0).astype(np.float32)               # exact ties at the 4th decimal and beyond
    return x, y, z


rng = np.random.default_rng(20251018)
out = {}
with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
    for case, n in enumerate((400, 1, 0, 5000)):
        x, y, z = cloud(rng, max(n, 64))
        x, y, z = x[:n], y[:n], z[:n]
        inten = (rng.random(n) ** 3 * 300).astype(np.float32)
        gains = rng.choice([40, 50, 70, 75, 60], n).astype(np.int32)
        norm = t5.normalize_intensity(inten)
        rgb_i = t5.intensity_to_rgb(norm) if n else np.zeros((0, 3), np.uint8)
        rgb_g = t5.gain_to_rgb(gains)
        fast, slow = Path(tmp) / f"fast{case}.ply", Path(tmp) / f"slow{case}.ply"
        t5.write_ply_fast(fast, x, y, z, rgb_i)
        t5.write_ply(slow, x, y, z, rgb_g)
        for k, v in (("x", x), ("y", y), ("z", z), ("inten", inten), ("gains", gains), ("norm", np.asarray(norm)), ("rgb_i", rgb_i), ("rgb_g", rgb_g),
                     ("fast", np.frombuffer(fast.read_bytes(), np.uint8)), ("slow", np.frombuffer(slow.read_bytes(), np.uint8))):
            out[f"c{case}_{k}"] = v
    # float64 coordinates: write_ply formats them as they are, write_ply_fast rounds to float32 first
    xd = rng.normal(0, 5e4, 50); yd = rng.normal(0, 5e4, 50); zd = rng.random(50) * 255    # float32 spacing ~0.004-0.008 here: the 4th decimal differs
    col = rng.integers(0, 256, (50, 3)).astype(np.uint8)
    f64a, f64b = Path(tmp) / "f64_slow.ply", Path(tmp) / "f64_fast.ply"
    t5.write_ply(f64a, xd, yd, zd, col); t5.write_ply_fast(f64b, xd, yd, zd, col)
    out.update(d_x=xd, d_y=yd, d_z=zd, d_col=col, d_slow=np.frombuffer(f64a.read_bytes(), np.uint8), d_fast=np.frombuffer(f64b.read_bytes(), np.uint8))
out["n_cases"] = np.array(4)
path = Path(__file__).resolve().parent / "ply_golden.npz"
np.savez_compressed(path, **out)
print("wrote", path, path.stat().st_size, "bytes")
