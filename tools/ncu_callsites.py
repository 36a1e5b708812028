"""Join ncu's per-SASS-instruction counts with nvdisasm's inline call chains: dynamic instructions and stall samples
per call site inside a source range (development aid).
usage: python tools/ncu_callsites.py ncu_sass.csv all_gi.sass <kernel substring> <file> <lo> <hi>"""
import collections, csv, re, sys
ncu, dis, kern, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
rows = list(csv.reader(open(ncu)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
dyn = []
for r in rows[2:]:
    if len(r) > isamp and r[ia].startswith("0x"):
        dyn.append((int(r[ia], 16), float(r[ii] or 0), float(r[isamp] or 0)))
base = dyn[0][0]
by_off = {a - base: (n, s) for a, n, s in dyn}
inside, chain, fresh = False, [], True
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for ln in open(dis):
    if ln.startswith("//---") and ".text." in ln:
        inside = all(k in ln for k in kern.split(","))
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        fresh = True
        off = int(m.group(1), 16)
        key = None
        for f, l in chain:
            if f == fname and lo <= l <= hi:
                key = l
        if key is None:
            key = "outside:%s:%s" % (chain[-1] if chain else ("?", 0))
        n, s = by_off.get(off, (0, 0))
        a = agg[key]
        a[0] += 1; a[1] += n; a[2] += s
tn = sum(a[1] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print("static %d  dynamic warp-instr %d  samples %d" % (sum(a[0] for a in agg.values()), tn, ts))
src = open("/root/repo/imitation-learning-rl_b200/csrc/" + fname).read().split("\n")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:int(sys.argv[7]) if len(sys.argv) > 7 else 40]:
    print("%5.1f%% smp %5.1f%% dyn  static %5d  %s  %s" % (100 * a[2] / ts, 100 * a[1] / tn, a[0], k, src[k - 1].strip()[:90] if isinstance(k, int) else ""))
