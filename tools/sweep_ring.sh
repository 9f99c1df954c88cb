#!/bin/bash
# Round-2 experiment, part 2: the mask kernel's TMA ring with the DRIVER's carve-outs (part 1 showed that a forced common
# carve-out costs the other kernels their L1 and gains nothing, while a ring that leaves shared memory free lets their
# CTAs become resident beside it). Usage (GPU box): bash tools/sweep_ring.sh > gpurun_out/sweep_ring.txt
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
    print(f"{label:44s} value {d['value']:9.0f} frames/s  step {d['ms_per_step']:6.3f} ms  mask alone {r['kernel_ms']:6.3f} ms ({r['achieved']:6.0f} GB/s)  "
          f"emit {r['stage']['kernels_ms']['spoke_emit_kernel']:5.3f} ms  sm {d['clocks']['sm_mhz']}", flush=True)
except Exception as e:
    print(f"{label:44s} FAILED: {e}: {line[:200]}", flush=True)
PY
}
run "32Kx3 ( 98 KB, default)"        RB_OPT_SPOKE_RING=2
run "48Kx2 ( 98 KB)"                 RB_OPT_SPOKE_RING=3
run "32Kx4 (131 KB)"                 RB_OPT_SPOKE_RING=1
run "64Kx3 (197 KB, round 1)"        RB_OPT_SPOKE_RING=0
run "32Kx3, no L2 evict-first hint"  RB_OPT_SPOKE_RING=2 RB_OPT_SPOKE_L2_HINT=0
run "32Kx3, no mask gate"            RB_OPT_SPOKE_RING=2 RB_OPT_MASK_GATE=0
run "32Kx3, no launch priority"      RB_OPT_SPOKE_RING=2 RB_OPT_MASK_PRIORITY=0
run "32Kx3 + carveout 44 everywhere" RB_OPT_SPOKE_RING=2 RB_OPT_CARVEOUT=44
for k in 1 2 3 4 6; do EXTRA="--streams $k" run "32Kx3, $k block(s) in flight" RB_OPT_SPOKE_RING=2; done
