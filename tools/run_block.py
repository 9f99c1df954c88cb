"""One bench-size block (reference defaults) on one GPU: counters and wall time; the target of per-kernel ncu captures.

    python tools/run_block.py [frames] [reps]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from radar_point_cloud_tracking_b200 import device as dev, synthetic as syn
from radar_point_cloud_tracking_b200.pipeline import DetectionConfig, DetectionPipeline

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = syn.SweepSpec(seed=2025, frames=frames, clutter_p=0.003)
d = torch.device("cuda:0")
pipe = DetectionPipeline(DetectionConfig(), 0)
echo = dev.synth_echo(spec, device=d)
c, s, r = pipe.spoke_tables(spec.angle_units(), spec.scale(), frames, spec.bins)
tabs = [torch.from_numpy(t).to(d) for t in (c, s, r)]
for i in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = pipe.run_device(echo, *tabs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = dev.stdbscan_stats(0)
    print(f"rep {i}: {frames} frames, {res.raw.n} raw points, {res.points.n} after land filter, {res.n_clusters} clusters, "
          f"{dt * 1e3:.2f} ms; tight={st['tight']} cells={st['n_cells']} dims={st['dims']} "
          f"tests count/union/border = {st['pair_tests_count']}/{st['pair_tests_union']}/{st['pair_tests_border']}")
lab = res.labels[:res.points.n]
print("noise:", int((lab < 0).sum()), "largest clusters:", torch.bincount(lab[lab >= 0]).topk(min(5, int(res.n_clusters))).values.tolist())
