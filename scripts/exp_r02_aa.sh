set -u
O=gpurun_out
run() { # label env...
  local label=$1; shift
  env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64 > $O/r2_bench_c5_$label.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_c5_$label.json'))
print('$label', round(d['ms_per_step'],2), '%.4g'%d['value'], round(d['phases']['sampler']['ms_per_step'],2), d['result_check']['oracle_stats_max_rel'], d['clocks']['sm_mhz'])
PY
}
run base A=1
run w288 MCLCAR_EMIT_WIDE=288 MCLCAR_EMIT_TR=16
run w320 MCLCAR_EMIT_WIDE=320 MCLCAR_EMIT_TR=18
run base2 A=1
run w320b MCLCAR_EMIT_WIDE=320 MCLCAR_EMIT_TR=18
python -m pytest tests/test_gpu_glm.py -q -k "both_proposals" 2>&1 | tail -3
