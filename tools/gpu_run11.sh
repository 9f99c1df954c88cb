#!/bin/bash
# 8 GPUs: NCCL parity at 8 ranks, bench at N=8 (with e2e) and N=4 (device resident only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r11_topo.txt 2>&1
for eps in 2.0 5.0; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded.py 12 $eps 2>&1 | grep -v -i "warn\|^$\|^\*\|OMP" | tail -4 | tee gpurun_out/r11_check_sharded_x8_eps$eps.txt
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r11_bench_n8.json 2> gpurun_out/r11_bench_n8.err
tail -3 gpurun_out/r11_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --steps 16 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r11_bench_n4.json 2> gpurun_out/r11_bench_n4.err
tail -3 gpurun_out/r11_bench_n4.err
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads(open(f'gpurun_out/r11_bench_n{n}.json').read().strip().splitlines()[-1])
        print(f'N={n} value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'identical', d.get('sharded_labels_identical'), 'e2e', d.get('e2e', {}).get('value'), 'u8', d.get('e2e_uint8_echoes', {}).get('value'), 'numa cpus', d['run'].get('numa_bound_cpus'))
    except Exception as e:
        print('bench parse failed', n, e)
PY
