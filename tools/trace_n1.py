"""Kernel timeline (torch.profiler) of the single-GPU overlapped pipeline at bench size: GPU busy fraction and
kernel concurrency across the worker streams.

    python tools/trace_n1.py [frames] [workers]
"""
import collections
import json
import re
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import ProfilerActivity, profile

from radar_point_cloud_tracking_b200 import device as dev, synthetic as syn
from radar_point_cloud_tracking_b200.pipeline import DetectionConfig, DetectionPipeline, OverlappedPipeline

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = syn.SweepSpec(seed=2025, frames=frames, clutter_p=0.003)
d = torch.device("cuda:0")
cfg = DetectionConfig()
base = DetectionPipeline(cfg, 0)
echo = dev.synth_echo(spec, device=d)
c, s, r = base.spoke_tables(spec.angle_units(), spec.scale(), frames, spec.bins)
tabs = [torch.from_numpy(t).to(d) for t in (c, s, r)]
ov = OverlappedPipeline(cfg, 0, workers=workers)
blk = ((echo, *tabs), {})
ov.map([blk] * (2 * workers), keep=False)
torch.cuda.synchronize()
n_blocks = 4 * workers
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ov.map([blk] * n_blocks, keep=False)
    torch.cuda.synchronize()
Path("gpurun_out").mkdir(exist_ok=True)
path = f"gpurun_out/n1_trace_w{workers}.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
t0 = min(e["ts"] for e in ev); t1 = max(e["ts"] + e["dur"] for e in ev)
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
busy = 0; cs, ce = iv[0]
for a, b in iv[1:]:
    if a > ce: busy += ce - cs; cs, ce = a, b
    else: ce = max(ce, b)
busy += ce - cs
tot = sum(e["dur"] for e in ev)
mask = [(e["ts"], e["ts"] + e["dur"]) for e in ev if "spoke_mask" in e["name"]]
mask_t = sum(b - a for a, b in mask)
# time of other kernels that runs while a mask kernel is running
ovl = 0.0
for e in ev:
    if "spoke_mask" in e["name"]: continue
    a, b = e["ts"], e["ts"] + e["dur"]
    for ma, mb in mask:
        lo, hi = max(a, ma), min(b, mb)
        if hi > lo: ovl += hi - lo
print(f"{n_blocks} blocks of {frames} frames, {workers} in flight: span {(t1 - t0) / 1e3:.2f} ms = {(t1 - t0) / 1e3 / n_blocks:.3f} ms/block; "
      f"GPU busy {100 * busy / (t1 - t0):.1f} %; kernel time {tot / 1e3 / n_blocks:.3f} ms/block of which mask {mask_t / 1e3 / n_blocks:.3f}; "
      f"non-mask kernel time overlapping a mask kernel {ovl / 1e3 / n_blocks:.3f} ms/block")
agg = collections.defaultdict(float)
def kname(e):
    m = re.search(r"(\w+)(<[^(]*>)?\(", e["name"])
    return m.group(1) if m else e["name"][:40]


for e in ev: agg[kname(e)] += e["dur"]
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:40]: print(f"  {v / 1e3 / n_blocks:7.3f} ms/block  {k}")
# where the time goes when NO mask kernel is running: the exposed tail
gaps = []
ms = sorted(mask)
cur_end = t0
for a, b in ms:
    if a > cur_end: gaps.append((cur_end, a))
    cur_end = max(cur_end, b)
if cur_end < t1: gaps.append((cur_end, t1))
exposed = sum(b - a for a, b in gaps)
agg2 = collections.defaultdict(float)
for e in ev:
    if "spoke_mask" in e["name"]: continue
    a, b = e["ts"], e["ts"] + e["dur"]
    for ga, gb in gaps:
        lo, hi = max(a, ga), min(b, gb)
        if hi > lo: agg2[kname(e)] += hi - lo
print(f"time with no mask kernel running: {exposed / 1e3 / n_blocks:.3f} ms/block; kernels running then (ms/block):")
for k, v in sorted(agg2.items(), key=lambda x: -x[1])[:10]: print(f"  {v / 1e3 / n_blocks:7.3f}  {k}")
conc = sum(min(b, mb) - max(a, ma) for i, (a, b) in enumerate(ms) for (ma, mb) in ms[i + 1:] if min(b, mb) > max(a, ma))
print(f"two mask kernels running at once: {conc / 1e3 / n_blocks:.3f} ms/block")
