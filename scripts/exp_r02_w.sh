set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2_pytest19.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_pytest19.log
python scripts/exp_host_algebra.py > $O/r2_host_algebra2.log 2>&1; echo "rc=$?"; tail -14 $O/r2_host_algebra2.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_c5_lazy.json 2> $O/r2_bench_c5_lazy.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_c5_lazy.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['result_check'])
PY
for c in c2 c3; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-330; done
