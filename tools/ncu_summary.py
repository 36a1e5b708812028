"""Summarise an `ncu --page raw --csv` export of the step kernel into a markdown table (development aid).
usage: python tools/ncu_summary.py raw.csv "title" "command" > profiles/xxx.md"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
print("# %s\n" % sys.argv[2])
print("Command (after the same command exited 0 without ncu): `%s`\n" % sys.argv[3])
print("| metric | unit | " + " | ".join("launch %d" % (i + 1) for i in range(len(data))) + " |")
print("|---|---|" + "---|" * len(data))
for h in want + stall:
    if h in hdr:
        i = hdr.index(h)
        vals = [r[i] for r in data]
        if h in stall and all(float(v or 0) < 0.05 for v in vals):
            continue
        print("| %s | %s | %s |" % (h, units[i], " | ".join(vals)))
