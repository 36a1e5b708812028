#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/b20.json 2> gpurun_out/b20.err; echo "rc=$?"; tail -c 600 gpurun_out/b20.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b20.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','vs_cpu_port')}))
print('e2e', json.dumps({k:v for k,v in d['e2e'].items() if k!='api' and k!='sync_api'}))
print('sequence', json.dumps({k:v for k,v in (d.get('sequence') or {}).items() if k!='api'}))
print('clocks', json.dumps(d['clocks']))
print('roofline', json.dumps({k:v for k,v in d['roofline'].items() if k not in('note',)}))
print('config', json.dumps({k:v for k,v in d['config'].items() if k in('blocks','block_ms_min','block_ms_max','timing')}))
print('cpu', json.dumps(d['cpu_baseline']))
for k,v in (d.get('configs') or {}).items(): print(k, json.dumps({kk:vv for kk,vv in v.items() if kk not in ('workload','sample_batch_columns')}))
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/b20_ref.json 2>> gpurun_out/b20.err; tail -c 900 gpurun_out/b20_ref.json
