#!/bin/bash
# Round-2 experiment: can the latency-bound clustering kernels of block k be RESIDENT next to the HBM-bound mask kernel of
# block k+1? Sweeps the mask kernel's TMA ring (rb option spoke_ring), a common shared-memory carve-out on every launch
# (carveout) and the L2 evict-first hint on the echo stream (spoke_l2_hint) over `bench.py` (device-resident value only).
# Usage (GPU box): bash tools/sweep_overlap.sh > gpurun_out/sweep_overlap.txt
cd "$(dirname "$0")/.."
run() {
    label="$1"; shift
    out=$(env "$@" python bench.py --no-e2e --no-cpu-baseline --steps ${STEPS:-24} --warmup 3 ${EXTRA} 2>/dev/null | tail -1)
    python - "$label" "$out" <<'PY'
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    r = d["roofline"]
    print(f"{label:58s} value {d['value']:9.0f} frames/s  step {d['ms_per_step']:6.3f} ms  mask {r['kernel_ms']:6.3f} ms ({r['achieved']:6.0f} GB/s)  "
          f"emit {r['stage']['kernels_ms']['spoke_emit_kernel']:5.3f} ms  sm {d['clocks']['sm_mhz']}", flush=True)
except Exception as e:
    print(f"{label:58s} FAILED: {e}: {line[:200]}", flush=True)
PY
}
run "baseline (64 KiB x 3, driver carve-outs)"               RB_OPT_SPOKE_RING=0
run "l2 evict-first only"                                    RB_OPT_SPOKE_RING=0 RB_OPT_SPOKE_L2_HINT=1
run "64Kx3 + carveout 100 (round-1 experiment)"              RB_OPT_SPOKE_RING=0 RB_OPT_CARVEOUT=100
run "32Kx4, driver carve-outs"                               RB_OPT_SPOKE_RING=1
run "32Kx4 + carveout 58"                                    RB_OPT_SPOKE_RING=1 RB_OPT_CARVEOUT=58
run "32Kx4 + carveout 58 + l2 hint"                          RB_OPT_SPOKE_RING=1 RB_OPT_CARVEOUT=58 RB_OPT_SPOKE_L2_HINT=1
run "32Kx3, driver carve-outs"                               RB_OPT_SPOKE_RING=2
run "32Kx3 + carveout 44"                                    RB_OPT_SPOKE_RING=2 RB_OPT_CARVEOUT=44
run "32Kx3 + carveout 44 + l2 hint"                          RB_OPT_SPOKE_RING=2 RB_OPT_CARVEOUT=44 RB_OPT_SPOKE_L2_HINT=1
run "64Kx2 + carveout 58 + l2 hint"                          RB_OPT_SPOKE_RING=3 RB_OPT_CARVEOUT=58 RB_OPT_SPOKE_L2_HINT=1
run "16Kx4 + carveout 44 + l2 hint"                          RB_OPT_SPOKE_RING=4 RB_OPT_CARVEOUT=44 RB_OPT_SPOKE_L2_HINT=1
EXTRA="--streams 4" run "32Kx4 + carveout 58 + l2 hint, 4 blocks in flight"   RB_OPT_SPOKE_RING=1 RB_OPT_CARVEOUT=58 RB_OPT_SPOKE_L2_HINT=1
EXTRA="--streams 4" run "32Kx3 + carveout 44 + l2 hint, 4 blocks in flight"   RB_OPT_SPOKE_RING=2 RB_OPT_CARVEOUT=44 RB_OPT_SPOKE_L2_HINT=1
EXTRA="--streams 2" run "32Kx4 + carveout 58 + l2 hint, 2 blocks in flight"   RB_OPT_SPOKE_RING=1 RB_OPT_CARVEOUT=58 RB_OPT_SPOKE_L2_HINT=1
