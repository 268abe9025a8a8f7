set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler_dcar.py tests/test_gpu_boundary.py tests/test_gpu_scale_parity.py -q -x --timeout 300 > $O/r2_pytest16.log 2>&1; tail -5 $O/r2_pytest16.log
for dbl in 0 1; do
MCLCAR_SWEEP_DOUBLE=$dbl python bench.py --steps 2 --warmup 1 --samples-per-gpu 1480 --chains 148 --no-cpu-baseline --no-fp64 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('short double=$dbl', 'launches', r['launches'], 'avg_launch_ms %.4f' % r['avg_launch_ms'], 'frac %.3f' % r['frac'], 'sampler ms/step %.3f' % d['phases']['sampler']['ms_per_step'], 'site-upd/s %.4g' % d['phases']['sampler']['value'])"
done
for dbl in 0 1; do
MCLCAR_SWEEP_DOUBLE=$dbl python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-fp64 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('full double=$dbl', 'value %.4g step %.2f ms' % (d['value'], d['ms_per_step']), 'launches', r['launches'], 'frac %.3f' % r['frac'], 'sampler ms/step %.3f' % d['phases']['sampler']['ms_per_step'], d['clocks'], d['result_check'].get('oracle_lr_max_abs'))"
done
