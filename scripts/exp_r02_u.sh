set -u
O=gpurun_out
python -m pytest tests/test_gpu_sampler_dcar.py tests/test_gpu_full_size.py tests/test_gpu_scale_parity.py -x -q > $O/r2_pytest18.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2_pytest18.log
for v in 2 3; do
MCLCAR_EMIT3_MINB=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64 > $O/r2_bench_c5_emit3_$v.json 2> $O/r2_bench_c5_emit3_$v.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_c5_emit3_$v.json'))
print($v, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['phases']['sampler']['ms_per_step'], d['phases']['likelihood'], d['result_check'])
PY
done
MCLCAR_NO_KEPT_FUSED_STATS=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64 > $O/r2_bench_c5_emit1.json 2>/dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_c5_emit1.json'))
print('emit1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['phases']['sampler']['ms_per_step'], d['phases']['likelihood']['ms_per_step'])
PY
