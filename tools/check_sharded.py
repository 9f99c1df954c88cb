"""Parity of the time-sharded path on real GPUs: run under torchrun with N ranks; every rank runs its shard,
rank 0 also runs the whole recording on one GPU and compares labels / points id for id.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded.py [frames_per_rank] [eps_time]
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from radar_point_cloud_tracking_b200 import device as dev, synthetic as syn
from radar_point_cloud_tracking_b200.pipeline import DetectionConfig, DetectionPipeline
from radar_point_cloud_tracking_b200.sharded import ShardedDetection

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
spec = syn.SweepSpec(seed=31, frames=B * world, spokes=512, bins=1024, clutter_p=0.004)
cfg = DetectionConfig(eps_time=float(sys.argv[2])) if len(sys.argv) > 2 else DetectionConfig()
sd = ShardedDetection(cfg, rank, world, local)
sd._cap_hint = 1000 + 7 * rank          # far too small: the first block must repeat its spoke stage on every rank
first = rank * B
echo = dev.synth_echo(spec, first_frame=first, n_frames=B, device=device)
c, s, r = sd.base.spoke_tables(spec.angle_units(), spec.scale(), B, spec.bins)
res = sd.run_device(echo, *(torch.from_numpy(t).to(device) for t in (c, s, r)), np.arange(first, first + B))
host = res.to_host()
gathered = [None] * world
dist.gather_object({"labels": host["labels"], "points": host["points"], "ncl": res.n_clusters, "halo": res.halo_points},
                   gathered if rank == 0 else None, dst=0)
ok = True
if rank == 0:
    full = dev.synth_echo(spec, device=device)
    pipe = DetectionPipeline(cfg, local)
    c, s, r = pipe.spoke_tables(spec.angle_units(), spec.scale(), spec.frames, spec.bins)
    ref = pipe.run_device(full, *(torch.from_numpy(t).to(device) for t in (c, s, r)))
    want = ref.to_host()
    got_l = np.concatenate([g["labels"] for g in gathered])
    got_p = np.concatenate([g["points"] for g in gathered])
    ok = np.array_equal(got_p, want["points"]) and np.array_equal(got_l, want["labels"]) and \
        all(g["ncl"] == ref.n_clusters for g in gathered)
    print(f"sharded x{world} (eps_time {cfg.eps_time}): {len(got_l)} points, {ref.n_clusters} clusters, halo points per rank "
          f"{[g['halo'] for g in gathered]} -> {'IDENTICAL to single GPU' if ok else 'MISMATCH'}")
# interleaved blocks (run_blocks: two generators in flight per rank): same labels per block
tabs = tuple(torch.from_numpy(t).to(device) for t in sd.base.spoke_tables(spec.angle_units(), spec.scale(), B, spec.bins))
many = sd.run_blocks([(echo, *tabs, np.arange(first, first + B))] * 4)
torch.cuda.synchronize()
same = all(torch.equal(m.labels, res.labels) and m.n_clusters == res.n_clusters and m.points.n == res.points.n for m in many)
t_same = torch.tensor([1 if same else 0], device=device)
dist.all_reduce(t_same, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"4 pipelined blocks per rank: {'IDENTICAL' if int(t_same.item()) else 'MISMATCH'}")
ok = ok and bool(int(t_same.item()))
flag = torch.tensor([1 if ok else 0], device=device)
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
