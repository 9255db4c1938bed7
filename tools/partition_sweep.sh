#!/bin/bash
# SM-partition sweep: run on a B200, e.g.  gpurun --timeout 400 -- 'bash tools/partition_sweep.sh'
# Each line: the knobs, the step time and the lanes / resampler / front-end kernel times of the roofline leg.
# Correctness first: the immediate-barrier lanes kernel through smoke() and the chain parity tests.
set -u
RFM_LANES_IMMBAR=1 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
RFM_LANES_IMMBAR=1 RFM_LANES_SMS=16 python -m pytest tests/test_gpu_parity.py -x -q -k "chain_every_stage or golden" 2>&1 | tail -1
run() { # SMS IMMBAR STREAMS
  RFM_LANES_SMS=$1 RFM_LANES_IMMBAR=$2 RFM_LANES_STREAMS=$3 python bench.py --no-e2e --no-cpu --steps 16 --warmup 4 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readlines()[-1]); k=d['kernel_ms_per_step']
    print('SMS=$1 IMMBAR=$2 STREAMS=$3', 'ms_per_step', round(d['ms_per_step'],4), {n:v for n,v in k.items() if n in ('k_bb_lanes','k_front','k_resample','k_rds_pll','k_audio_tail')})
except Exception as e: print('SMS=$1 IMMBAR=$2 STREAMS=$3 failed', e)"
}
run 0 0 AO
run 0 1 AO
for n in 8 16 24 32; do run $n 1 AO; done
run 16 1 AOP
run 16 1 AOPC
run 8 1 AOPC
# What the co-resident pilot warps contend for (one capture each; read with `ncu -i ... --page raw --csv`, look at
# smsp__inst_executed_pipe_fp64 / _xu, smsp__issue_active, launch__occupancy_limit_*, smsp__warp_issue_stalled_*):
if [ "${WITH_NCU:-0}" = 1 ]; then
  for n in 0 16 32; do
    RFM_LANES_SMS=$n RFM_LANES_IMMBAR=1 ncu --set full --clock-control none --import-source on -k regex:k_bb_lanes -s 6 -c 1 \
      -f -o gpurun_out/lanes_sms$n python bench.py --no-e2e --no-cpu --no-prof --steps 4 --warmup 3 > /dev/null 2>&1
  done
fi
