"""DRAM bytes and executed warp-instructions per launch of the step kernel from an `ncu --page raw --csv` export ->
profiles/step_kernel_traffic.json (read by bench.py for `roofline.traffic` of the contract workload).
usage: python tools/ncu_traffic.py raw.csv "source description" [n_envs]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
h, u, d = rows[0], rows[1], rows[2:]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
sc = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
ir, iw, ii = h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum'), h.index('smsp__inst_executed.sum')
tot = [float(r[ir]) * sc[u[ir]] + float(r[iw]) * sc[u[iw]] for r in d]
ins = [float(r[ii]) for r in d]
json.dump({"dram_bytes_per_launch": int(sum(tot) / len(tot)), "warp_inst_per_env_step": round(sum(ins) / len(ins) / n, 1),
           "n_envs": n, "source": sys.argv[2] + " (dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum; "
           "mean of %d steady-state launches, N=%d)" % (len(tot), n)}, open('profiles/step_kernel_traffic.json', 'w'), indent=1)
print(open('profiles/step_kernel_traffic.json').read())
