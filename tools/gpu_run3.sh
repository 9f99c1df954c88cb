#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 2000 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/r3_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r3_pytest.txt
timeout 300 python tools/time_clusters.py 1024 > gpurun_out/r3_time_clusters.txt 2>&1
timeout 900 bash tools/sweep_ring.sh > gpurun_out/r3_sweep_ring.txt 2>&1
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r3_pytest.txt | tail -15
cat gpurun_out/r3_time_clusters.txt
cat gpurun_out/r3_sweep_ring.txt
