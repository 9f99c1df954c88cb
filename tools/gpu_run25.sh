#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_shard_pack.py tests/test_gpu_sharded.py -q -x 2>&1 | tail -4 | tee gpurun_out/r25_pytest.txt
timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --workload config4 --steps 4 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r25_config4_n2.json
python -c "
import json; d=json.loads(open('gpurun_out/r25_config4_n2.json').read()); print('N=2 config4:', round(d['value']), round(d['ms_per_step'],3), d.get('sharded_labels_identical'))"
