# end-of-round check on one B200: GPU tests, smoke(), the default bench line, the small configurations, the reference arm
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2_pytest28.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_pytest28.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > $O/r2_bench_default.json 2> $O/r2_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_default.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['clocks'], d['steps'], d['warmup'], d['gpu_launches'])
PY
for c in c1 c3 c4 c4p; do python bench.py --config $c 2>/dev/null > $O/r2_bench_$c.json; python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_$c.json'))
print('$c', round(d['ms_per_step'],3), '%.4g'%d['value'], d['clocks'])
PY
done
python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-400
