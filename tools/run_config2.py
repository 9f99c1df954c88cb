"""BASELINE config 2 at scale: 3_stdbscan_point_clouds.py-style flat ST-DBSCAN on a stacked single-gain cloud
(coords = x, y, intensity; time = gain index = 0 for all; eps 5 / 1 / 10) built from `frames` synthetic sweeps.

    python tools/run_config2.py [frames] [stride]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from radar_point_cloud_tracking_b200 import device as dev, synthetic as syn
from radar_point_cloud_tracking_b200.tracker import sweep_tables

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 500
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 4
spec = syn.SweepSpec(seed=2, frames=frames, gains=(75,))
d = torch.device("cuda:0")
echo = dev.synth_echo(spec, device=d).view(frames, spec.spokes, spec.bins)
c, s, r = sweep_tables(spec.angle_units(), spec.scale(), spec.bins)
rep = lambda t: torch.from_numpy(np.ascontiguousarray(np.broadcast_to(t, (frames, len(t)))).copy()).to(d)
b = dev.spoke_to_points(echo, rep(c), rep(s), rep(r), torch.full((frames,), 75, dtype=torch.int32, device=d), 10.0, stride)
n = b.n
times = torch.zeros(n, dtype=torch.float32, device=d)
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lab, ncl = dev.stdbscan(b.x, b.y, b.inten, times, 5.0, 1.0, 10, stride=1, n=n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = dev.stdbscan_stats(0)
    print(f"rep {i}: {n} points (3-D), {ncl} clusters, noise {int((lab < 0).sum())}, {dt * 1e3:.1f} ms = {n / dt / 1e6:.1f} M points/s; "
          f"tight={st['tight']} dims={st['dims']} tests={st['pair_tests_count']}/{st['pair_tests_union']}/{st['pair_tests_border']}")
