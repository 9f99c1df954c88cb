#!/bin/bash
cd "$(dirname "$0")/.."
( timeout 2000 python -m pytest tests -m gpu -q ) > gpurun_out/r23_pytest.txt 2>&1
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r23_pytest.txt | tail -5
b() { label="$1"; shift; out=$(timeout 300 python bench.py --no-cpu-baseline --steps 8 "$@" 2>/dev/null | tail -1); python -c "
import json,sys;d=json.loads(sys.argv[2]);e=d['e2e'];u=d['e2e_uint8_echoes'];print(f'{sys.argv[1]:36s}', 'e2e', round(e['value']), round(e['ms_per_step'],2), 'copy-only', round(e['h2d_copy_only']['frames_per_s']), 'u8', round(u['value']), round(u['ms_per_step'],2))" "$label" "$out"; }
b "e2e 128 frames, 2 in flight" --e2e-frames 128 --e2e-in-flight 2
b "e2e 128 frames, 3 in flight" --e2e-frames 128 --e2e-in-flight 3
b "e2e 256 frames, 2 in flight" --e2e-frames 256 --e2e-in-flight 2
b "e2e 256 frames, 3 in flight" --e2e-frames 256 --e2e-in-flight 3
b "e2e 64 frames, 4 in flight" --e2e-frames 64 --e2e-in-flight 4
