#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 2000 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r6_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r6_pytest.txt
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r6_pytest.txt | tail -15
b() { label="$1"; shift; out=$(env "$@" timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 24 $EXTRA 2>/dev/null | tail -1); python -c "
import json,sys;d=json.loads(sys.argv[2]);print(f'{sys.argv[1]:46s}', round(d['value']), round(d['ms_per_step'],3), 'mask', round(d['roofline']['kernel_ms'],3), d['stdbscan']['pair_tests'])" "$label" "$out"; }
b "ring2 gate prio"            RB_OPT_SPOKE_RING=2
b "ring2 gate noprio"          RB_OPT_SPOKE_RING=2 RB_OPT_MASK_PRIORITY=0
b "ring2 nogate prio"          RB_OPT_SPOKE_RING=2 RB_OPT_MASK_GATE=0
b "ring0 gate prio"            RB_OPT_SPOKE_RING=0
b "ring1 (32Kx4) gate prio"    RB_OPT_SPOKE_RING=1
b "ring2 gate prio l2hint"     RB_OPT_SPOKE_RING=2 RB_OPT_SPOKE_L2_HINT=1
EXTRA="--streams 4" b "ring2 gate prio 4 in flight"   RB_OPT_SPOKE_RING=2
EXTRA="--streams 2" b "ring2 gate prio 2 in flight"   RB_OPT_SPOKE_RING=2
echo "=== trace ring 2 gate prio"
RB_OPT_SPOKE_RING=2 timeout 300 python tools/trace_n1.py 1024 3 2>&1 | grep -v -i warn | tee gpurun_out/r6_trace_ring2_gate_prio.txt
echo "=== trace 1 block in flight (kernel times alone)"
RB_OPT_SPOKE_RING=2 timeout 300 python tools/trace_n1.py 1024 1 2>&1 | grep -v -i warn | head -24 | tee gpurun_out/r6_trace_alone.txt
rm -f gpurun_out/n1_trace_w*.json
