#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 2000 python -m pytest tests -m gpu -q ) > gpurun_out/r17_pytest.txt 2>&1
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r17_pytest.txt | tail -8
for w in config4 config3 config2; do
out=$(timeout 400 python bench.py --workload $w --no-e2e --no-cpu-baseline --steps 12 2>/dev/null | tail -1); python -c "
import json,sys;d=json.loads(sys.argv[1]);print('$w', round(d['value'],1), round(d['ms_per_step'],3), d['run']['parallelism'])" "$out"
done
