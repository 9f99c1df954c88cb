#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 8 6; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$k bench.py --gpus 8 --steps 32 --warmup 3 --no-cpu-baseline --no-e2e --shard-in-flight $k 2>/dev/null | tail -1 > gpurun_out/r24_n8_k$k.json
python -c "
import json; d=json.loads(open('gpurun_out/r24_n8_k$k.json').read()); print('N=8 in flight $k:', round(d['value']), round(d['ms_per_step'],3), d.get('sharded_labels_identical'))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --workload config4 --steps 8 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r24_config4_n8.json
python -c "
import json; d=json.loads(open('gpurun_out/r24_config4_n8.json').read()); print('N=8 config4:', round(d['value']), round(d['ms_per_step'],3), d.get('sharded_labels_identical'), d['points_per_s'])"
