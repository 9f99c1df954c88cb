"""Runs only the spoke-to-point stage on synthetic sweeps (for ncu captures and quick timing).

    python tools/run_spoke.py [frames] [reps] [thr] [stride]
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from radar_point_cloud_tracking_b200 import _lib, device as dev, synthetic as syn
from radar_point_cloud_tracking_b200.tracker import sweep_tables

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
stride = int(sys.argv[4]) if len(sys.argv) > 4 else 4
spec = syn.SweepSpec(seed=2025, frames=frames)
d = torch.device("cuda:0")
W = frames * 3
echo = dev.synth_echo(spec, device=d).view(W, spec.spokes, spec.bins)
if os.environ.get('RB_U8'):
    echo = echo.to(torch.uint8)
c, s, r = sweep_tables(spec.angle_units(), spec.scale(), spec.bins)
rep = lambda t: torch.from_numpy(np.ascontiguousarray(np.broadcast_to(t, (W, len(t)))).copy()).to(d)
c, s, r = rep(c), rep(s), rep(r)
gains = torch.tensor([40, 50, 75] * frames, dtype=torch.int32, device=d)
cap = dev.default_capacity(W, spec.spokes, spec.bins, stride)
out = None
ctx = _lib.context(0)
ctx.set_option("spoke_profile", 1)
import os
ctx.set_option("spoke_mask_variant", int(os.environ.get("RB_MASK_VARIANT", "0")))
rows = []
for i in range(reps):
    x, y, inten, gain, base = dev.spoke_to_points_raw(echo, c, s, r, gains, thr, stride, cap, out=out)
    out = (x, y, inten, gain, base)
    rows.append([ctx.info(k) * 1e-3 for k in ("spoke_mask_ns", "spoke_offsets_ns", "spoke_emit_ns")])
torch.cuda.synchronize()
n = int(base[-1].item())
gb = echo.numel() * echo.element_size() / 1e9
for i, (a, b, e) in enumerate(rows):
    print(f"rep {i}: mask {a:8.1f} us ({gb / a * 1e6:7.1f} GB/s)  offsets {b:6.1f} us  emit {e:8.1f} us  "
          f"stage {(gb + 16e-9 * n) / (a + b + e) * 1e6:7.1f} GB/s")
print(f"mask variant used: {ctx.info('spoke_last_variant')}")
print(f"frames {frames} thr {thr} stride {stride}: {n} points kept, echo {gb:.3f} GB")
