#!/bin/bash
# N-GPU pass: the multi-rank test, then the reference arm and our arm of bench.py under torchrun.  usage: gpu_multi.sh N
N=${1:-8}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/test_gpu_multi_${N}gpu.log 2>&1; tail -3 gpurun_out/test_gpu_multi_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}), 'e2e', d['e2e']['value'], 'seq', (d.get('sequence') or {}).get('value'))
for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
