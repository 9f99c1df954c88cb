cd /root/repo
timeout 120 python tools/run_block.py 512 4 2>&1 | tail -2
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout=600 --timeout-method=thread 2>&1 | tail -2
timeout 120 python tools/run_dense.py 8 3 2>&1 | tail -2 | cut -c1-250
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_block.csv python tools/run_block.py 512 3 > /dev/null 2>&1; echo ncu rc=$?
