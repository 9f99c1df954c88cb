cd /root/repo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dbt_union_kernel" -s 2 -c 1 -o gpurun_out/r01_ncu_union_v3 python tools/run_block.py 512 3 > gpurun_out/ncu_db.log 2>&1; echo ncu rc=$?
