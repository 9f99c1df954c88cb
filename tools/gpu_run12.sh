#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 2000 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r12_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r12_pytest.txt
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r12_pytest.txt | tail -15
b() { label="$1"; shift; out=$(env "$@" timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 32 $EXTRA 2>/dev/null | tail -1); python -c "
import json,sys;d=json.loads(sys.argv[2]);print(f'{sys.argv[1]:46s}', round(d['value']), round(d['ms_per_step'],3), 'mask', round(d['roofline']['kernel_ms'],3), d['stdbscan']['pair_tests'])" "$label" "$out"; }
EXTRA="--streams 3" b "3 in flight" A=1
EXTRA="--streams 4" b "4 in flight" A=1
EXTRA="--streams 5" b "5 in flight" A=1
EXTRA="--streams 6" b "6 in flight" A=1
EXTRA="--streams 8" b "8 in flight" A=1
echo "=== trace 1 block in flight (kernel times alone)"
timeout 300 python tools/trace_n1.py 1024 1 2>&1 | grep -v -i warn | head -12 | tee gpurun_out/r12_trace_alone.txt
echo "=== trace 4 in flight"
timeout 300 python tools/trace_n1.py 1024 4 2>&1 | grep -v -i warn | tee gpurun_out/r12_trace_w4.txt | head -14
grep -A12 "no mask" gpurun_out/r12_trace_w4.txt
rm -f gpurun_out/n1_trace_w*.json
timeout 600 python bench.py > gpurun_out/r12_bench_config3.json 2> gpurun_out/r12_bench_config3.err
python -c "
import json;d=json.loads(open('gpurun_out/r12_bench_config3.json').read().strip().splitlines()[-1]);print('full bench', round(d['value']), d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'u8', d['e2e_uint8_echoes']['value'], 'cpu', d['cpu_baseline']['value'], 'roof', d['roofline']['frac'])"
