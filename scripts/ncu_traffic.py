#!/usr/bin/env python
"""Extracts the per-launch DRAM traffic (and a few companions) of one kernel from an `ncu --set full` report into the small JSON
file bench.py cites for `roofline.traffic`:

    ncu -i gpurun_out/<report>.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/ncu_traffic.py /tmp/raw.csv k_heat_euler_fast "<the profiled command>" profiles/r3j_ncu_bench_fast_kernel.json
"""
import csv
import json
import sys

raw, pattern, command, out = sys.argv[1:5]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale_units=True):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    if scale_units:
        v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    return v


sel = [r for r in data if pattern in r[col["Kernel Name"]] and "nan" not in r[col["smsp__inst_executed.sum"]]]
if not sel:
    raise SystemExit(f"no fully captured launch of a kernel matching {pattern!r}")
r = max(sel, key=lambda r: val(r, "gpu__time_duration.sum", False))       # the longest launch = the bench step
d = {
    "kernel": r[col["Kernel Name"]], "command": command,
    "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
    "duration_under_ncu": f"{r[col['gpu__time_duration.sum']]} {units[col['gpu__time_duration.sum']]}",
    "registers_per_thread": int(float(r[col["launch__registers_per_thread"]])),
    "fp64_pipe_pct": float(r[col["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]]),
    "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
    "inst_executed": float(r[col["smsp__inst_executed.sum"]]),
    "launches_captured": len(sel),
}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d, indent=1))
