#!/bin/bash
# Kernel experiments: lean Darcy / NS bench lines under combinations of the scratch switches exp0..exp2 (env UNO_B200_EXPn).
#   gpurun --timeout 1200 -- 'CONFIGS="EXP0=0 EXP0=1 EXP0=2,EXP1=3" WORKLOADS="darcy" bash tools/run_exp.sh'
set -u
mkdir -p gpurun_out
OUT=gpurun_out/${TAG:-exp}.log
: > $OUT
for wl in ${WORKLOADS:-darcy}; do
  for cfg in ${CONFIGS:-"EXP0=0"}; do
    envs=$(echo $cfg | tr ',' ' ' | sed 's/\([A-Z0-9_]*=\)/UNO_B200_\1/g')
    echo "=== $wl [$envs]" >> $OUT
    env $envs timeout 300 python bench.py --workload $wl --lean --steps 10 --warmup 3 2>>$OUT | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    kb = d.get('kernel_breakdown') or {}
    print(json.dumps({'ms_per_step': round(d['ms_per_step'], 3), 'value': round(d['value'], 1), 'e2e': round(d['e2e']['value'], 1),
                      'top': {k: round(v['ms_per_step'], 3) for k, v in list(kb.items())[:12]}}))" >> $OUT 2>&1
  done
done
cat $OUT
