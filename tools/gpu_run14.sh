#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 4 6 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2953$k bench.py --gpus 4 --steps 24 --warmup 3 --no-cpu-baseline --no-e2e --shard-in-flight $k 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=4 in flight $k:', round(d['value']), round(d['ms_per_step'],3), d.get('sharded_labels_identical'))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 4 --workload config5 --steps 16 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=4 config5:', round(d['value']), round(d['ms_per_step'],3), d.get('sharded_labels_identical'))"
