# Round-end verification on one B200 (run through gpurun): tests, smoke, both bench arms, then the ncu evidence.
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=600 --timeout-method=thread > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/final_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/final_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spoke_mask_tma_kernel -s 3 -c 1 -o gpurun_out/r01_ncu_spoke_final python tools/run_spoke.py 32 6 > gpurun_out/final_ncu_spoke.log 2>&1; echo "ncu spoke rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dbt_union_kernel|dbt_count_kernel|dbt_border_kernel" -s 6 -c 3 -o gpurun_out/r01_ncu_dbscan_final python tools/run_block.py 512 3 > gpurun_out/final_ncu_db.log 2>&1; echo "ncu dbscan rc=$?"
