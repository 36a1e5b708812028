"""Static SASS size of a kernel per call site inside a chosen source range (development aid).
usage: nvdisasm -gi -c x.cubin > all.sass; python tools/sass_regions.py all.sass <kernel substring> <file> <lo> <hi>"""
import collections, re, sys
path, kern, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
inside = False
chain, fresh = [], True
cnt = collections.Counter()
total = 0
for ln in open(path):
    if ln.startswith("//---") and ".text." in ln:
        inside = all(k in ln for k in kern.split(","))
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        fresh = True
        total += 1
        key = None
        for f, l in chain:
            if f == fname and lo <= l <= hi:
                key = l
        cnt[key if key is not None else ("outside", chain[-1] if chain else None)] += 1
print("total", total)
src = open("/root/repo/imitation-learning-rl_b200/csrc/" + fname).read().split("\n")
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:45]:
    print("%6d  %s  %s" % (v, k, src[k - 1].strip()[:100] if isinstance(k, int) else ""))
