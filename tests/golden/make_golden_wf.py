"""Golden vectors of the PointCloudWorkF ST-DBSCAN variant (min_frames core test), produced by the UNMODIFIED
reference function ``PointCloudWorkF/stdbscan_denoising_pipeline.py::st_dbscan`` (WF:264-369), imported by path in
the build container (``/root/reference`` does not exist on the GPU box).

    python tests/golden/make_golden_wf.py        ->  tests/golden/wf_stdbscan.npz
"""
import importlib.util
from pathlib import Path

import numpy as np

REF = "/root/reference/PointCloudWorkF/stdbscan_denoising_pipeline.py"
spec = importlib.util.spec_from_file_location("wf_ref", REF)
wf = importlib.util.module_from_spec(spec)
spec.loader.exec_module(wf)

rng = np.random.default_rng(20251018)
out = {}
cases = 0
for trial in range(20):
    dim = 3 if trial % 5 == 4 else 2
    n = int(rng.integers(300, 1400))
    span = 70.0
    coords = (rng.random((n, dim)) * span).astype(np.float32)
    k = n * 2 // 3
    centres = (rng.random((7, dim)) * span).astype(np.float32)
    coords[:k] = (centres[rng.integers(0, 7, k)] + rng.normal(0, 2.2, (k, dim))).astype(np.float32)
    coords = coords[rng.permutation(n)]
    frames = int(rng.integers(2, 9))
    if trial % 4 == 3:
        times = (rng.random(n) * frames).astype(np.float32)            # fractional times: int32 truncation matters
    else:
        times = rng.integers(0, frames, n).astype(np.float32)
    eps_s = float(rng.choice([3.0, 4.5, 6.0, 8.0]))
    eps_t = float(rng.choice([1.0, 2.0, 0.5, 1.5, 3.0]))
    ms = int(rng.choice([4, 6, 10, 15]))
    mf = int(rng.choice([1, 2, 3, 4]))
    labels = wf.st_dbscan(coords, times, eps_s, eps_t, ms, mf)
    out[f"c{cases}_coords"], out[f"c{cases}_times"] = coords, times
    out[f"c{cases}_params"] = np.array([eps_s, eps_t, ms, mf], dtype=np.float64)
    out[f"c{cases}_labels"] = labels.astype(np.int32)
    cases += 1
out["n_cases"] = np.array(cases)
path = Path(__file__).resolve().parent / "wf_stdbscan.npz"
np.savez_compressed(path, **out)
print("wrote", path, cases, "cases;", sum(int((out[f"c{i}_labels"] >= 0).sum()) for i in range(cases)), "clustered points")
