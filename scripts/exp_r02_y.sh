set -u
O=gpurun_out
python -m pytest tests/test_gpu_glm.py tests/test_gpu_hcar.py tests/test_gpu_sampler_dcar.py tests/test_gpu_callers.py tests/test_gpu_estimates_scale.py tests/test_gpu_boundary.py tests/test_gpu_scale_parity.py tests/test_gpu_full_size.py -x -q > $O/r2_pytest21.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_pytest21.log
for c in c3 c4 c1; do
  for g in 1 0; do
    MCLCAR_CSR_REBALANCE=$g python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_bench_${c}_reb$g.json 2>/dev/null
    python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${c}_reb$g.json'))
print('$c rebalance=$g', d['ms_per_step'], d['value'], d['phases']['sampler'])
PY
  done
done
