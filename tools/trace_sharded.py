"""Kernel/host timeline (torch.profiler -> chrome trace) of the time-sharded step at bench size; run under torchrun.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/trace_sharded.py [frames] [in_flight]
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

from radar_point_cloud_tracking_b200 import device as dev, synthetic as syn
from radar_point_cloud_tracking_b200.pipeline import DetectionConfig
from radar_point_cloud_tracking_b200.sharded import ShardedDetection

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spec = syn.SweepSpec(seed=7, frames=B * world)
sd = ShardedDetection(DetectionConfig(), rank, world, local)
first = rank * B
echo = dev.synth_echo(spec, first_frame=first, n_frames=B, device=device)
tabs = tuple(torch.from_numpy(t).to(device) for t in sd.base.spoke_tables(spec.angle_units(), spec.scale(), B, spec.bins))
blk = (echo, *tabs, np.arange(first, first + B))
sd.run_blocks([blk] * 6, keep=False, in_flight=K)
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    sd.run_blocks([blk] * 8, keep=False, in_flight=K)
    torch.cuda.synchronize()
if rank == 0:
    import collections
    import json
    import re
    Path("gpurun_out").mkdir(exist_ok=True)
    path = f"gpurun_out/shard_trace_k{K}.json"
    prof.export_chrome_trace(path)
    n_blocks = 8
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    os.remove(path)

    def kname(e):
        m = re.search(r"(\w+)(<[^(]*>)?\(", e["name"])
        return m.group(1) if m else e["name"][:40]
    t0 = min(e["ts"] for e in ev); t1 = max(e["ts"] + e["dur"] for e in ev)
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
    busy = 0; cs, ce = iv[0]
    for a, b in iv[1:]:
        if a > ce: busy += ce - cs; cs, ce = a, b
        else: ce = max(ce, b)
    busy += ce - cs
    mask = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev if "spoke_mask" in e["name"])
    gaps, cur = [], t0
    for a, b in mask:
        if a > cur: gaps.append((cur, a))
        cur = max(cur, b)
    if cur < t1: gaps.append((cur, t1))
    print(f"rank 0 of {world}: {n_blocks} blocks of {B} frames, {K} in flight: span {(t1 - t0) / 1e3 / n_blocks:.3f} ms/block; GPU busy {100 * busy / (t1 - t0):.1f} %; "
          f"kernel time {sum(e['dur'] for e in ev) / 1e3 / n_blocks:.3f} ms/block of which mask {sum(b - a for a, b in mask) / 1e3 / n_blocks:.3f}; "
          f"no mask kernel running {sum(b - a for a, b in gaps) / 1e3 / n_blocks:.3f} ms/block; {len(ev) / n_blocks:.0f} GPU activities per block")
    agg = collections.defaultdict(float)
    for e in ev: agg[kname(e)] += e["dur"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:16]: print(f"  {v / 1e3 / n_blocks:7.3f} ms/block  {k}")
dist.barrier()
dist.destroy_process_group()
