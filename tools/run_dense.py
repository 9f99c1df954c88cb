"""Dense-clutter stress (BASELINE config 4): thr 2.0, stride 2, eps_space 12 -> ~2.2 M points per frame.

    python tools/run_dense.py [frames] [reps]
Prints per-stage CUDA-event times of the device pipeline and the ST-DBSCAN work counters.
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from radar_point_cloud_tracking_b200 import device as dev, synthetic as syn
from radar_point_cloud_tracking_b200.pipeline import DetectionConfig, DetectionPipeline

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spec = syn.SweepSpec(seed=404, frames=frames)
d = torch.device("cuda:0")
cfg = DetectionConfig(intensity_threshold=2.0, point_stride=2, eps_space=12.0)
pipe = DetectionPipeline(cfg, 0)
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
    print(f"rep {i}: {frames} frames, {res.raw.n} points ({res.raw.n / frames / 1e6:.2f} M/frame), after land filter {res.points.n}, "
          f"{res.n_clusters} clusters, {dt * 1e3:.1f} ms -> {frames / dt:.1f} frames/s, {res.raw.n / dt / 1e6:.1f} M points/s; "
          f"tight={st['tight']} cells={st['n_cells']} tests count/union/border = {st['pair_tests_count']}/{st['pair_tests_union']}/{st['pair_tests_border']}")
lab = res.labels[:res.points.n]
print("label histogram (top 5):", torch.bincount(lab[lab >= 0]).topk(min(5, int(res.n_clusters))).values.tolist() if res.n_clusters else [], "noise:", int((lab < 0).sum()))
