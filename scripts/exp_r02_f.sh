set -u
O=gpurun_out
python -m pytest tests/test_gpu_sampler_dcar.py tests/test_gpu_boundary.py tests/test_gpu_dcar_parity.py tests/test_gpu_scale_parity.py -q -x --timeout 900 > $O/r2_pytest9.log 2>&1; tail -3 $O/r2_pytest9.log
CONFIGS="MCLCAR_SWEEP_TR=14" NS=1480 EXTRA="--chains 148" OUT=$O/sweep_exp3.log bash scripts/sweep_exp.sh
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-fp64 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('full: value %.4g sweep_ms %.4f frac %.3f step %.2f ms clocks %s' % (d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['ms_per_step'], d['clocks']))"
