"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line (development aid).
usage: python tools/ncu_lines.py export.csv [topN]"""
import collections, csv, os, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
res = collections.defaultdict(lambda: collections.defaultdict(float))
cur = hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if not r or cur is None or hdr is None or not r[0].isdigit():
        continue
    d = dict(zip(hdr[4:], r[4:]))
    for k, v in d.items():
        try:
            res[(cur, int(r[0]))][k] += float(v)
        except ValueError:
            pass
tot_i = sum(v["Instructions Executed"] for v in res.values())
tot_s = sum(v["# Samples"] for v in res.values())
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
keys = ["stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_branch_resolving", "stall_selected", "stall_math", "stall_mio", "stall_lg", "stall_dispatch"]
print("stall totals:", {k: int(sum(v[k] for v in res.values())) for k in keys})
src = {}
for (fn, ln), v in sorted(res.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    if fn not in src:
        src[fn] = open(fn).read().split("\n") if os.path.exists(fn) else []
    s = src[fn]
    print("%5.2f%% smp %5.2f%% inst  noi=%5d lsb=%5d ssb=%5d wait=%5d  %s:%d  %s" % (
        100 * v["# Samples"] / tot_s, 100 * v["Instructions Executed"] / tot_i, v["stall_no_inst"], v["stall_long_sb"],
        v["stall_short_sb"], v["stall_wait"], os.path.basename(fn), ln, s[ln - 1].strip()[:80] if ln - 1 < len(s) else ""))
