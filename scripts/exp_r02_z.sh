set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2_pytest24.log 2>&1; echo "pytest rc=$?"; tail -8 $O/r2_pytest24.log
for c in c3; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_bench_${c}_laplace.json 2>/dev/null;  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${c}_laplace.json'))
print('$c', d['ms_per_step'], d['value'], d['phases'], d.get('result_check'))
PY
done
