#!/bin/bash
# 8 GPUs, final build: config 3 at N=8 (full line with e2e), config 5 and config 4 at N=8 (device resident), config 3 at N=4 and N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { n=$1; w=$2; shift 2; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$((40+RANDOM%50)) bench.py --gpus $n --workload $w --no-cpu-baseline "$@" > gpurun_out/r18_bench_${w}_n$n.json 2> gpurun_out/r18_bench_${w}_n$n.err; tail -2 gpurun_out/r18_bench_${w}_n$n.err | grep -i -E "error|Traceback" ; }
run 8 config3 --steps 24 --warmup 3
run 8 config5 --steps 16 --warmup 3 --no-e2e
run 8 config4 --steps 8 --warmup 3 --no-e2e
run 4 config3 --steps 24 --warmup 3
run 2 config3 --steps 24 --warmup 3
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r18_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e', {})
        print(f.split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'identical', d.get('sharded_labels_identical'), 'e2e', round(e.get('value', 0)), 'u8', round(d.get('e2e_uint8_echoes', {}).get('value', 0)), 'copy-only', round(e.get('h2d_copy_only', {}).get('frames_per_s', 0)))
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
