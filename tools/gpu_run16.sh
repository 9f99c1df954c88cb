#!/bin/bash
cd "$(dirname "$0")/.."
b() { label="$1"; shift; out=$(env "$@" timeout 300 python bench.py --workload config4 --no-e2e --no-cpu-baseline --steps 6 $EXTRA 2>/dev/null | tail -1); python -c "
import json,sys;d=json.loads(sys.argv[2]);print(f'{sys.argv[1]:40s}', round(d['value'],1), round(d['ms_per_step'],2), 'mask', round(d['roofline']['kernel_ms'],3), d['stdbscan']['pair_tests'])" "$label" "$out"; }
EXTRA="--streams 1" b "config4 1 in flight" A=1
EXTRA="--streams 2" b "config4 2 in flight" A=1
EXTRA="--streams 3" b "config4 3 in flight" A=1
EXTRA="--streams 4" b "config4 4 in flight" A=1
EXTRA="--streams 3" b "config4 3 in flight nogate noprio" RB_OPT_MASK_GATE=0 RB_OPT_MASK_PRIORITY=0
EXTRA="--streams 3" b "config4 3 in flight ring0" RB_OPT_SPOKE_RING=0
python - <<'PY'
# per-kernel times of one config-4 block run alone
import sys, re, json, collections
sys.path.insert(0, '.')
import numpy as np, torch
from torch.profiler import ProfilerActivity, profile
from radar_point_cloud_tracking_b200 import device as dev, synthetic as syn
from radar_point_cloud_tracking_b200.pipeline import DetectionConfig, DetectionPipeline
spec = syn.SweepSpec(seed=2025, frames=64, clutter_p=0.003)
cfg = DetectionConfig(intensity_threshold=2.0, point_stride=2, eps_space=12.0)
pipe = DetectionPipeline(cfg, 0)
echo = dev.synth_echo(spec)
tabs = [torch.from_numpy(t).cuda() for t in pipe.spoke_tables(spec.angle_units(), spec.scale(), 64, spec.bins)]
pipe.run_device(echo, *tabs); pipe.run_device(echo, *tabs)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pipe.run_device(echo, *tabs)
    torch.cuda.synchronize()
prof.export_chrome_trace('/tmp/t.json')
ev = [e for e in json.load(open('/tmp/t.json'))['traceEvents'] if e.get('ph') == 'X' and e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
agg = collections.defaultdict(float)
for e in ev:
    m = re.search(r"(\w+)(<[^(]*>)?\(", e['name']); agg[m.group(1) if m else e['name'][:30]] += e['dur']
print('config-4 block alone: kernel time', round(sum(agg.values()) / 1e3, 2), 'ms')
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:14]: print(f'  {v/1e3:8.3f} ms  {k}')
PY
