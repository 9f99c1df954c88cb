"""Host-side profile (cProfile) of the time-sharded step at bench size; run under torchrun.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/profile_sharded.py [frames] [in_flight]
"""
import cProfile
import io
import os
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

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
sd.run_blocks([blk] * (2 * K), keep=False, in_flight=K)          # every block slot warm (streams, contexts, allocator pools)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
sd.run_blocks([blk] * 16, keep=False, in_flight=K)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 16 * 1e3
dist.barrier()
pr = cProfile.Profile()
pr.enable()
sd.run_blocks([blk] * 16, keep=False, in_flight=K)
torch.cuda.synchronize()
pr.disable()
if rank == 0:
    out = io.StringIO()
    st = pstats.Stats(pr, stream=out)
    st.sort_stats("tottime").print_stats(45)
    st.sort_stats("cumulative").print_stats(45)
    Path("gpurun_out").mkdir(exist_ok=True)
    # host time that is NOT waiting for the GPU: everything except event queries / synchronisations / blocking copies
    waits = ("query", "synchronize", "'cpu' of", "'item' of", "<genexpr>", "builtins.any")
    total = sum(v[2] for v in st.stats.values())
    waiting = sum(v[2] for k, v in st.stats.items() if any(w in k[2] for w in waits))
    head = (f"wall per block without profiler: {wall:.3f} ms (in_flight={K}, {B} frames/rank, world {world})\n"
            f"under cProfile, 16 blocks: {total * 1e3:.1f} ms of host time, of which {waiting * 1e3:.1f} ms polling / waiting for the GPU "
            f"(event.query, synchronize) -> host WORK per block {(total - waiting) / 16 * 1e3:.3f} ms (cProfile inflates Python-level work ~2x)\n")
    Path(f"gpurun_out/shard_hostprof_k{K}.txt").write_text(head + out.getvalue())
    print(head)
dist.destroy_process_group()
