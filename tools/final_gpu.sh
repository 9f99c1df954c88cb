#!/bin/bash
# Round-end check on one B200: GPU test suite, smoke(), the default bench line, the reference arm, the other workloads.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 2000 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest.log
grep -E "passed|failed|FAILED|ERROR|rc=" gpurun_out/final_pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
for w in config5 config4 config2; do
  timeout 900 python bench.py --workload $w --steps 12 > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err
done
python - <<'PY'
import json
for n in ("final_bench", "final_bench_ref", "final_bench_config5", "final_bench_config4", "final_bench_config2"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        e = d.get("e2e", {})
        print(n, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(e.get("value", 0), 1), "u8", round(d.get("e2e_uint8_echoes", {}).get("value", 0), 1),
              "cpu", round(d.get("cpu_baseline", {}).get("value", 0), 4), "cores", d.get("cpu_baseline", {}).get("cores"), "roof", d.get("roofline", {}).get("frac"),
              "copy-only", e.get("h2d_copy_only", {}).get("frames_per_s"), "launches", d.get("gpu_launches"), "clocks", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
    except Exception as ex:
        print(n, "FAILED", ex)
PY
