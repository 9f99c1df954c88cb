#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 2000 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r7_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r7_pytest.txt
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r7_pytest.txt | tail -15
timeout 900 bash tools/sweep_ring.sh 2>&1 | tee gpurun_out/r7_sweep_ring.txt
echo "=== trace 1 block in flight (kernel times alone)"
timeout 300 python tools/trace_n1.py 1024 1 2>&1 | grep -v -i warn | head -12 | tee gpurun_out/r7_trace_alone.txt
rm -f gpurun_out/n1_trace_w*.json
