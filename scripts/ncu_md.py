import csv, sys
fn, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(fn)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name): return hdr.index(name) if name in hdr else None
metrics = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
           "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
kn = col("Kernel Name")
# every fully captured launch, in launch order
chk = col("smsp__inst_executed.sum")
keep = {}
for k, r in enumerate(data):
    if "nan" in r[chk]: continue
    keep[f"#{k + 1} " + r[kn]] = r
names = list(keep)
with open(out, "w") as f:
    f.write(f"# {title}\n\n")
    f.write("| metric | " + " | ".join(n.replace("void cgb::", "").replace("void ", "").split("(")[0] for n in names) + " | unit |\n")
    f.write("|---|" + "---|" * (len(names) + 1) + "\n")
    for m in metrics:
        i = col(m)
        if i is None: continue
        f.write(f"| {m} | " + " | ".join(keep[n][i] for n in names) + f" | {units[i]} |\n")
    f.write("\nStall reasons per issued instruction (warp-cycles stalled per issue):\n\n")
    f.write("| stall | " + " | ".join(n.replace("void cgb::", "").replace("void ", "").split("(")[0] for n in names) + " |\n")
    f.write("|---|" + "---|" * len(names) + "\n")
    for s in stalls:
        i = col(s)
        vals = [float(keep[n][i]) for n in names]
        if max(vals) < 0.1: continue
        f.write("| " + s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "") + " | " + " | ".join(f"{v:.2f}" for v in vals) + " |\n")
print(open(out).read())
