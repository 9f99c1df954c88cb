#!/bin/bash
cd "$(dirname "$0")/.."
b() { label="$1"; shift; out=$(env "$@" timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 32 $EXTRA 2>/dev/null | tail -1); python -c "
import json,sys;d=json.loads(sys.argv[2]);print(f'{sys.argv[1]:46s}', round(d['value']), round(d['ms_per_step'],3), 'mask', round(d['roofline']['kernel_ms'],3))" "$label" "$out"; }
b "default (32Kx3, driver carve-outs)" A=1
b "32Kx4 + carveout 58 on every launch" RB_OPT_SPOKE_RING=1 RB_OPT_CARVEOUT=58
b "32Kx4 + carveout 60 on every launch" RB_OPT_SPOKE_RING=1 RB_OPT_CARVEOUT=60
b "64Kx3 + carveout 100 on every launch" RB_OPT_SPOKE_RING=0 RB_OPT_CARVEOUT=100
b "32Kx3 + carveout 44 on every launch" RB_OPT_SPOKE_RING=2 RB_OPT_CARVEOUT=44
