#!/bin/bash
# A/B runs of bench.py on the GPU box: each argument is "LABEL|ENV..." ; prints one compact line per run.
#   gpurun -- 'bash tools/ab.sh "spec|" "nospec|RFM_LIB_PATH=$PWD/pvr.rtl.radiofm_b200/libradiofm_b200_exp.so RFM_LANES_NOSPEC=1"'
set -u
EXP=$PWD/pvr.rtl.radiofm_b200/libradiofm_b200_exp.so
for spec in "$@"; do
  label=${spec%%|*}; envs=${spec#*|}
  envs=${envs//@EXP/RFM_LIB_PATH=$EXP}
  env $envs python bench.py --no-e2e --no-cpu --no-extras ${BENCH_ARGS:-} --steps ${STEPS:-16} --warmup 4 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readlines()[-1]); k=d['kernel_ms_per_step']
    print('$label', 'ms_per_step', round(d['ms_per_step'],4), 'repairs', d.get('demod_repairs'), {n:v for n,v in k.items() if v>0.05})
except Exception as e: print('$label failed', e)"
done
