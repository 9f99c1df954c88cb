#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 2000 python -m pytest tests -m gpu -q --durations=25 ) > gpurun_out/r2_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest.txt
timeout 900 bash tools/sweep_ring.sh > gpurun_out/r2_sweep_ring.txt 2>&1
tail -40 gpurun_out/r2_pytest.txt
cat gpurun_out/r2_sweep_ring.txt
