# how does the power-capped full-size step respond to fewer instructions? (results are meaningless under DRY: timing only)
set -u
O=gpurun_out
for dry in 0 1 2; do
  MCLCAR_SWEEP_DRY=$dry python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-fp64 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('DRY=$dry sweep_ms %.4f frac %.3f step %.2f ms clocks %s' % (d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['ms_per_step'], d['clocks']))"
done
