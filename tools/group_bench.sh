#!/bin/bash
# cost grouping (K8) on / off and its period, device-timed, on the multi-wave workloads
for wl in hier16384 multiclip65536 hier2_16384; do
  for ev in 0 1 2 4 8; do
    echo -n "$wl every=$ev: "
    ILRL_GROUP_EVERY=$ev python bench.py --workload $wl --steps 20 --warmup 10 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.2f M  %.1f us' % (d['value']/1e6, d['ms_per_step']*1e3))"
  done
done
