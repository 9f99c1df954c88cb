"""Per-sweep load time: pandas (the reference's parser, T4:192) vs the device CSV ingest, on full-size synthetic sweeps.

    python tools/time_csv_ingest.py [n_files]
"""
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from radar_point_cloud_tracking_b200 import synthetic as syn, tracker as trk

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 6
spec = syn.SweepSpec(seed=3, frames=(n_files + 2) // 3)
with tempfile.TemporaryDirectory() as tmp:
    files = [p for entry in syn.write_csv_tree(spec, Path(tmp)) for p in entry.values()][:n_files]
    size = sum(p.stat().st_size for p in files) / len(files) / 1e6
    trk.read_sweep_csv_device(files[0]); trk.read_sweep_csv(files[0])          # warm-up (imports, contexts, page cache)
    t0 = time.perf_counter()
    host = [trk.read_sweep_csv(p) for p in files]
    t1 = time.perf_counter()
    devs = [trk.read_sweep_csv_device(p) for p in files]
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    same = all(np.array_equal(h[2], d[2].cpu().numpy().astype(np.float32)) and np.array_equal(h[0], d[0]) and np.array_equal(h[1], d[1]) and h[3] == d[3]
               for h, d in zip(host, devs))
    # whole load_radar_csv (parse + spoke-to-point + points back on the host)
    t3 = time.perf_counter()
    for p in files:
        trk.load_radar_csv(p)
    t4 = time.perf_counter()
    print(f"{len(files)} sweeps of {spec.spokes} x {spec.bins}, {size:.1f} MB each: pandas parse {1e3 * (t1 - t0) / len(files):.1f} ms/sweep, "
          f"device ingest {1e3 * (t2 - t1) / len(files):.1f} ms/sweep (file read + upload + kernels + leading fields through pandas), "
          f"identical: {same}; load_radar_csv end to end {1e3 * (t4 - t3) / len(files):.1f} ms/sweep")
