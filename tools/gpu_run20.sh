#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/profile_sharded.py 1024 6 2>&1 | grep -v -i "warn\|^\*\|OMP\|^$" | tail -6
timeout 600 python -m pytest tests/test_gpu_sharded.py -q 2>&1 | tail -3
