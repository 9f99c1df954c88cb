#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 2000 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r4_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r4_pytest.txt
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r4_pytest.txt | tail -15
for ring in 0 2; do
  echo "=== ring $ring" 
  RB_OPT_SPOKE_RING=$ring timeout 300 python tools/trace_n1.py 1024 3 2>&1 | grep -v Warning | tee gpurun_out/r4_trace_ring$ring.txt
  RB_OPT_SPOKE_RING=$ring timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 24 2>/dev/null | tail -1 > gpurun_out/r4_bench_ring$ring.json
  python -c "
import json;d=json.load(open('gpurun_out/r4_bench_ring$ring.json'));print('bench ring $ring', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['stdbscan']['pair_tests'])"
done
rm -f gpurun_out/n1_trace_w3.json
