#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 2000 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r5_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r5_pytest.txt
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r5_pytest.txt | tail -15
b() { label="$1"; shift; out=$(env "$@" timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 24 2>/dev/null | tail -1); python -c "
import json,sys;d=json.loads(sys.argv[2]);print(f'{sys.argv[1]:46s}', round(d['value']), round(d['ms_per_step'],3), 'mask', round(d['roofline']['kernel_ms'],3), d['stdbscan']['pair_tests'])" "$label" "$out"; }
b "ring0 gate"                 RB_OPT_SPOKE_RING=0
b "ring0 nogate"               RB_OPT_SPOKE_RING=0 RB_OPT_MASK_GATE=0
b "ring2 gate"                 RB_OPT_SPOKE_RING=2
b "ring2 nogate"               RB_OPT_SPOKE_RING=2 RB_OPT_MASK_GATE=0
b "ring2 gate oldcount"        RB_OPT_SPOKE_RING=2 RB_OPT_COUNT_VARIANT=1
b "ring2 gate l2hint"          RB_OPT_SPOKE_RING=2 RB_OPT_SPOKE_L2_HINT=1
b "ring6 (48Kx2) gate"         RB_OPT_SPOKE_RING=6
b "ring1 (32Kx4) gate"         RB_OPT_SPOKE_RING=1
for ring in 2; do
  echo "=== trace ring $ring gate" 
  RB_OPT_SPOKE_RING=$ring timeout 300 python tools/trace_n1.py 1024 3 2>&1 | grep -v -i warn | tee gpurun_out/r5_trace_ring${ring}_gate.txt
  echo "=== trace ring $ring gate oldcount" 
  RB_OPT_COUNT_VARIANT=1 RB_OPT_SPOKE_RING=$ring timeout 300 python tools/trace_n1.py 1024 3 2>&1 | grep -v -i warn | head -30 | tee gpurun_out/r5_trace_ring${ring}_gate_oldcount.txt
done
echo "=== trace 1 block in flight (kernel times alone)"
RB_OPT_SPOKE_RING=2 timeout 300 python tools/trace_n1.py 1024 1 2>&1 | grep -v -i warn | head -44 | tee gpurun_out/r5_trace_alone.txt
rm -f gpurun_out/n1_trace_w*.json
