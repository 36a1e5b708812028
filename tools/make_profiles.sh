#!/bin/bash
# Regenerate the committed profile artefacts of a kernel generation from a gpurun_out/ capture (development aid).
# usage: tools/make_profiles.sh <ncu-rep stem in gpurun_out/> <profiles prefix, e.g. r1_v3> <bench stem> <launch-list stem>
set -e
REP=gpurun_out/$1.ncu-rep; P=profiles/$2
ncu -i $REP --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
ncu -i $REP --page source --csv --print-source cuda,sass --launch-skip 0 --launch-count 1 > gpurun_out/$1_cs.csv 2>/dev/null
ncu -i $REP --page source --csv --print-source sass --launch-skip 0 --launch-count 1 > gpurun_out/$1_sass.csv 2>/dev/null
rm -rf /tmp/cub && mkdir -p /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all /root/repo/imitation-learning-rl_b200/libilrl_b200.so >/dev/null && for f in *.cubin; do nvdisasm -gi -c $f 2>/dev/null >> all_gi.sass; done)
python tools/ncu_summary.py gpurun_out/$1_raw.csv "$2 — ncu --set full of \`step_kernel<0>\`, 3 steady-state launches (skip 400), N = 4096 envs, B200" "ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 400 -c 3 python bench.py --steps 500 --warmup 5 --no-cpu-baseline" > ${P}_step_kernel_ncu.md
python tools/ncu_lines.py gpurun_out/$1_cs.csv 50 > ${P}_step_kernel_lines.txt
L=$(grep -n "^__device__ __forceinline__ int substep(" imitation-learning-rl_b200/csrc/ilrl_chain.cuh | cut -d: -f1)
H=$(awk -v l=$L 'NR>=l && /^}/ {print NR; exit}' imitation-learning-rl_b200/csrc/ilrl_chain.cuh)   # closing brace of substep()
python tools/ncu_callsites.py gpurun_out/$1_sass.csv /tmp/cub/all_gi.sass "step_kernelILi0,LayoutILi42E" ilrl_chain.cuh $L $H 30 > ${P}_step_kernel_callsites.txt
cp gpurun_out/$3.json ${P}_bench.json; cp gpurun_out/$3_ref.json ${P}_bench_reference_arm.json; cp gpurun_out/$4.csv ${P}_launches.csv
python - <<PY
import csv, json
rows = list(csv.reader(open('gpurun_out/$1_raw.csv')))
h, u, d = rows[0], rows[1], rows[2:]
sc = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6}
tot = [float(r[h.index('dram__bytes_read.sum')]) * sc[u[h.index('dram__bytes_read.sum')]] + float(r[h.index('dram__bytes_write.sum')]) * sc[u[h.index('dram__bytes_write.sum')]] for r in d]
wi = [float(r[h.index('smsp__inst_executed.sum')]) for r in d]
json.dump({"dram_bytes_per_launch": int(sum(tot) / len(tot)), "warp_inst_per_env_step": round(sum(wi) / len(wi) / 4096, 1), "source": "${P}_step_kernel_ncu.md (dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum; mean of %d steady-state launches, N=4096)" % len(tot)}, open('profiles/step_kernel_traffic.json', 'w'))
PY
