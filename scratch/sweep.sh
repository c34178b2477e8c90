#!/bin/bash
# usage: scratch/sweep.sh TAG "ENV1=.. ENV2=.." ["ENV.." ...]   -> one bench line per environment, summarised
tag=$1; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-job --no-other-configs > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - "$envs" gpurun_out/${tag}_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    k = {n[len('kernel_ms_'):]: v for n, v in d['roofline'].items() if n.startswith('kernel_ms_')}
    print(f"{sys.argv[1]:40s} step {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms  " + ' '.join(f'{n} {v:.3f}' for n, v in k.items()))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
