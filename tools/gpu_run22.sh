#!/bin/bash
cd "$(dirname "$0")/.."
b() { label="$1"; shift; out=$(env "$@" timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 32 $EXTRA 2>/dev/null | tail -1); python -c "
import json,sys;d=json.loads(sys.argv[2]);print(f'{sys.argv[1]:46s}', round(d['value']), round(d['ms_per_step'],3), 'mask', round(d['roofline']['kernel_ms'],3))" "$label" "$out"; }
b "default" A=1
b "tail_smem 16K" RB_OPT_TAIL_SMEM=16384
b "tail_smem 24K" RB_OPT_TAIL_SMEM=24576
b "tail_smem 32K" RB_OPT_TAIL_SMEM=32768
b "tail_smem 40K" RB_OPT_TAIL_SMEM=40960
b "tail_smem 48K" RB_OPT_TAIL_SMEM=49152
EXTRA="--streams 6" b "tail_smem 32K, 6 in flight" RB_OPT_TAIL_SMEM=32768
EXTRA="--streams 6" b "tail_smem 48K, 6 in flight" RB_OPT_TAIL_SMEM=49152
