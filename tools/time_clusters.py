"""Host cost of the per-frame Cluster records (T4:511-534) for one bench block, before / after rb_cluster_records:
   python tools/time_clusters.py [frames]   (GPU box)"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from radar_point_cloud_tracking_b200 import device as dev, synthetic as syn
from radar_point_cloud_tracking_b200.pipeline import DetectionConfig, DetectionPipeline

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.cuda.set_device(0)
spec = syn.SweepSpec(seed=2025, frames=F)
pipe = DetectionPipeline(DetectionConfig(), 0)
echo = dev.synth_echo(spec)
tabs = [torch.from_numpy(t).cuda() for t in pipe.spoke_tables(spec.angle_units(), spec.scale(), F, spec.bins)]
res = pipe.run_device(echo, *tabs)
torch.cuda.synchronize()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


t_host, a = timed(lambda: res.clusters_by_frame_host())
t_dev, b = timed(lambda: res.clusters_by_frame())
t_rec, rec = timed(lambda: dev.cluster_records(res.points, res.labels, res.n_clusters))
n = sum(len(v) for v in a.values())
same = list(a) == list(b) and all(np.array_equal(x.centroid, y.centroid) and x.cluster_id == y.cluster_id and x.mean_intensity == y.mean_intensity
                                  for f in a for x, y in zip(a[f], b[f]))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ctx = pipe.ctx
l0 = ctx.launch_count()
ev0.record()
dev.cluster_records(res.points, res.labels, res.n_clusters)
ev1.record()
torch.cuda.synchronize()
print(f"{F}-frame block: {res.points.n} points, {res.n_clusters} clusters, {n} (frame, cluster) records; identical: {same}")
print(f"  host loop (reference's masks + np.mean, incl. D2H of points/labels): {t_host * 1e3:8.2f} ms")
print(f"  device records + Cluster objects (views):                          {t_dev * 1e3:8.2f} ms")
print(f"    of which rb_cluster_records + one packed D2H:                     {t_rec * 1e3:8.2f} ms "
      f"(GPU time {ev0.elapsed_time(ev1):.3f} ms, {ctx.launch_count() - l0} launches)")
