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
    Path("gpurun_out").mkdir(exist_ok=True)
    prof.export_chrome_trace(f"gpurun_out/shard_trace_k{K}.json")
dist.barrier()
dist.destroy_process_group()
