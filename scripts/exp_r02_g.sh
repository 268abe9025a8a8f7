set -u
O=gpurun_out
python -m pytest tests/test_gpu_sampler_dcar.py tests/test_gpu_boundary.py tests/test_gpu_glm.py -q -x --timeout 900 > $O/r2_pytest10.log 2>&1; tail -2 $O/r2_pytest10.log
CONFIGS="MCLCAR_SWEEP_TR=14;MCLCAR_SWEEP_TR=14 MCLCAR_SWEEP_THREADS=224;MCLCAR_SWEEP_TR=14 MCLCAR_SWEEP_THREADS=192;MCLCAR_SWEEP_TR=12;MCLCAR_SWEEP_TR=10" NS=1480 EXTRA="--chains 148" OUT=$O/sweep_exp4.log bash scripts/sweep_exp.sh
for c in c2 c1; do python bench.py --config $c 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$c', '%.3f ms/step' % d['ms_per_step'], d['kernels'], d['roofline'].get('frac'))"; done
python bench.py --steps 3 --warmup 3 --stats-only --no-cpu-baseline --no-fp64 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('stats-only: value %.4g step %.2f ms' % (d['value'], d['ms_per_step']))"
