#!/bin/bash
# 2 GPUs: single-GPU suite part for the pack kernels, NCCL parity (check_sharded), bench at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard_pack.py tests/test_gpu_sharded.py -q 2>&1 | tail -15 | tee gpurun_out/r10_pytest.txt
for eps in 2.0 5.0; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded.py 12 $eps 2>&1 | grep -v -i "warn\|^$" | tail -8 | tee gpurun_out/r10_check_sharded_x2_eps$eps.txt
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench_n2.json 2> gpurun_out/r10_bench_n2.err
tail -3 gpurun_out/r10_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r10_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2 value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'identical', d.get('sharded_labels_identical'), 'e2e', d.get('e2e', {}).get('value'), 'u8', d.get('e2e_uint8_echoes', {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 16 --warmup 3 --no-cpu-baseline --no-e2e --shard-in-flight 4 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2 4 in flight', round(d['value']), d['ms_per_step'])"
