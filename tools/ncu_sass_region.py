"""Print the SASS of one call site (inline chain containing file:line) with ncu's per-instruction execution counts and
stall samples (development aid).  usage: python tools/ncu_sass_region.py ncu_sass.csv all_gi.sass <kernel substrings> <file> <line>"""
import csv, re, sys
ncu, dis, kern, fname, line = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
rows = list(csv.reader(open(ncu)))
hdr = rows[1]
ia, ii, isamp, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
dyn = [(int(r[ia], 16), float(r[ii] or 0), float(r[isamp] or 0), r[isrc]) for r in rows[2:] if len(r) > isamp and r[ia].startswith("0x")]
base = dyn[0][0]
by_off = {a - base: (n, s, t) for a, n, s, t in dyn}
inside, chain, fresh = False, [], True
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
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
    if m:
        fresh = True
        off = int(m.group(1), 16)
        if any(f == fname and l == line for f, l in chain):
            n, s, t = by_off.get(off, (0, 0, ""))
            inner = [c for c in chain if not (c[0] == fname and c[1] == line)]
            print("%06x %9d %4d  %-70s  %s" % (off, n, s, m.group(2)[:70], inner[0] if inner else ""))
