set -u
O=gpurun_out
python bench.py --debug --steps 5 --warmup 3 --no-cpu-baseline --no-fp64 > $O/r2_bench_c5_mu.json 2> $O/r2_bench_c5_mu.err; echo "rc=$?"
grep debug $O/r2_bench_c5_mu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c5_mu.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['phases']['sampler']['ms_per_step'], d['phases']['likelihood']['ms_per_step'], d['clocks'])
PY
python -m pytest tests -m gpu -x -q > $O/r2_pytest17.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_pytest17.log
