#!/bin/bash
# usage: scratch/cfg_other.sh CFG FEATURES [lib.so]   -> one other_configs line of bench.py for that configuration
cfg=$1; feats=$2; lib=$3
env ${lib:+YC_LIB_PATH=$PWD/$lib} YC_BENCH_OTHER=$cfg YC_BENCH_CFG${cfg}_FEATURES=$feats python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-job > gpurun_out/other_tmp.json 2> gpurun_out/other_tmp.err
python - "$lib" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/other_tmp.json').read().strip().split('\n')[-1])
for k, v in d['other_configs'].items():
    if 'error' in v:
        print(sys.argv[1], k, v['error'])
    else:
        print(f"{sys.argv[1] or 'base':32s} {k} {v['positions_per_s'] / 1e6:.3f} M pos/s  {v['ms_per_step']:.1f} ms  " + ' '.join(f'{n} {x:.1f}' for n, x in v['kernel_ms'].items()))
PY
