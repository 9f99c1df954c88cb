#!/bin/bash
# ncu evidence of round 2 (one GPU): --set full of the hot kernels on the 1024-frame bench block, and the launch list of
# the bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/run_block.py 1024 2 > gpurun_out/ncu_plain_block.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"spoke_mask|dbt_count|dbt_union|dbt_border|spoke_emit|land_accumulate" -s 6 -c 6 \
    -o gpurun_out/r02_ncu_block python tools/run_block.py 1024 2 > gpurun_out/ncu_ncu_block.log 2>&1
echo "ncu block rc=$?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
tail -3 gpurun_out/ncu_plain_block.log
ls -la gpurun_out/r02_ncu_block.ncu-rep gpurun_out/r02_launches.csv
