#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 2000 python -m pytest tests -m gpu -q ) > gpurun_out/r19_pytest.txt 2>&1
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r19_pytest.txt | tail -8
out=$(timeout 400 python bench.py --workload config4 --steps 12 2>/dev/null | tail -1); echo "$out" > gpurun_out/r19_bench_config4.json; python -c "
import json,sys;d=json.loads(sys.argv[1]);print('config4', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d['e2e']['value'], 'u8', d['e2e_uint8_echoes']['value'], 'cpu', d['cpu_baseline']['value'], d['points_per_step'])" "$out"
