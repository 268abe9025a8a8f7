set -u
O=gpurun_out
python scripts/exp_host_algebra.py > $O/r2_host_algebra.log 2>&1; echo "rc=$?"; tail -45 $O/r2_host_algebra.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_c5_emit3.json 2> $O/r2_bench_c5_emit3.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_c5_emit3.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['phases']['likelihood'], d['likelihood_fp64_resident'])
PY
