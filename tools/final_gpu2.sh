#!/bin/bash
# last GPU call of round 2: GPU tests, then the default bench line and the same with pageable read-backs (A/B of the e2e path)
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final2_pytest.txt
python bench.py --no-cpu-baseline 2>gpurun_out/final2_bench.err | tail -1 > gpurun_out/final2_bench.json
RB_HOST_READBACK=pageable python bench.py --no-cpu-baseline --steps 8 2>/dev/null | tail -1 > gpurun_out/final2_bench_pageable.json
RB_HOST_READBACK=pinned python bench.py --no-cpu-baseline --steps 8 2>/dev/null | tail -1 > gpurun_out/final2_bench_pinned.json
cat gpurun_out/final2_pytest.txt
python - <<'P'
import json
for f in ("final2_bench.json", "final2_bench_pageable.json", "final2_bench_pinned.json"):
    try:
        d = json.load(open("gpurun_out/" + f)); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 1), round(d["e2e_uint8_echoes"]["value"]))
    except Exception as e:
        print(f, "unreadable", e)
P
