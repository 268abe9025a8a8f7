#!/bin/bash
# sweep-kernel experiments: CONFIGS is a ;-separated list of env assignments (MCLCAR_SWEEP_FUSED=0|1, MCLCAR_SWEEP_TR=<rows>,
# MCLCAR_SWEEP_THREADS=<threads>); prints the sweep time per launch.  EXTRA passes bench flags (e.g. "--chains 148"), NS the samples per GPU.
out=${OUT:-gpurun_out/sweep_exp.log}
: > $out
IFS=';' read -ra CFG <<< "${CONFIGS:-MCLCAR_SWEEP_FUSED=1;MCLCAR_SWEEP_FUSED=1 MCLCAR_SWEEP_TR=15;MCLCAR_SWEEP_FUSED=0}"
for cfg in "${CFG[@]}"; do
  env $cfg python bench.py --steps 2 --warmup 1 --samples-per-gpu ${NS:-1320} ${EXTRA} --no-cpu-baseline --no-fp64 2>>gpurun_out/sweep_exp.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('$cfg', 'chains', d['config']['chains_per_gpu'], 'avg_launch_ms', round(r['avg_launch_ms'], 5), 'frac', round(r['frac'], 4), 'site_updates/s', '%.4g' % d['phases']['sampler']['value'], 'value', '%.4g' % d['value'])
" >> $out
done
cat $out
