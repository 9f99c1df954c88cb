"""Turn an `ncu -i X.ncu-rep --page raw --csv` export into the per-launch DRAM traffic record bench.py reads.

    python tools/ncu_traffic.py RAW.csv KERNEL_SUBSTRING FRAMES_PER_LAUNCH ALGORITHMIC_BYTES "how it was captured" > profiles/..._traffic.json
"""
import csv
import json
import sys

raw, kernel, frames, alg, how = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
rows = list(csv.reader(open(raw)))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h, units = rows[hdr], rows[hdr + 1]
row = next(r for r in rows[hdr + 2:] if kernel in r[h.index("Kernel Name")])


def metric(name, want_unit):
    i = h.index(name)
    v, u = float(row[i].replace(",", "")), units[i]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "%": 1}.get(u, 1)
    return v * scale


print(json.dumps({
    "kernel": kernel, "capture": how, "frames_per_launch": frames, "algorithmic_bytes": alg,
    "dram_bytes_read": int(metric("dram__bytes_read.sum", "byte")), "dram_bytes_write": int(metric("dram__bytes_write.sum", "byte")),
    "gpu_time_us": round(metric("gpu__time_duration.sum", "us"), 3),
    "dram_throughput_pct_of_ncu_peak": round(metric("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%"), 2)}))
