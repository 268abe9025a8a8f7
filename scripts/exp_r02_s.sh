set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2_pytest16.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_pytest16.log
for c in c1 c2 c3 c4; do
  python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_bench_${c}_chains.json 2> $O/r2_bench_${c}_chains.err; echo "$c rc=$?"; cat $O/r2_bench_${c}_chains.json | cut -c1-600
  MCLCAR_CHAINS_FILL=1 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_bench_${c}_fill.json 2>/dev/null; cut -c1-300 $O/r2_bench_${c}_fill.json
done
