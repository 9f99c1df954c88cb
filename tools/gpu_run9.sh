#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/profile_sharded.py 1024 3 2>&1 | grep -v -i "warn\|^\*\|OMP\|^$" | tail -5
head -75 gpurun_out/shard_hostprof_k3.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/trace_sharded.py 1024 3 2>&1 | grep -v -i "warn\|^\*\|OMP\|^$" | tail -24 | tee gpurun_out/r9_trace_sharded.txt
