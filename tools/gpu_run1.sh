#!/bin/bash
# GPU call 1 of round 2: the whole GPU suite, the default bench line, the carve-out sweep, the other workloads.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/r1_gpu.txt 2>&1
nproc >> gpurun_out/r1_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r1_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r1_pytest.txt
timeout 600 python bench.py > gpurun_out/r1_bench_config3.json 2> gpurun_out/r1_bench_config3.err
timeout 900 bash tools/sweep_overlap.sh > gpurun_out/r1_sweep_overlap.txt 2>&1
for w in config5 config4 config2; do
  timeout 600 python bench.py --workload $w --steps 8 > gpurun_out/r1_bench_$w.json 2> gpurun_out/r1_bench_$w.err
done
tail -5 gpurun_out/r1_pytest.txt
cat gpurun_out/r1_sweep_overlap.txt
