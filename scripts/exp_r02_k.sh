set -u
O=gpurun_out
python -m pytest tests/test_gpu_sampler_dcar.py tests/test_gpu_scale_parity.py tests/test_gpu_full_size.py tests/test_gpu_dcar_parity.py tests/test_gpu_estimates_scale.py -q -x --timeout 900 > $O/r2_pytest13.log 2>&1; tail -3 $O/r2_pytest13.log
python bench.py --config c2 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('c2', '%.3f ms/step value %.4g' % (d['ms_per_step'], d['value']), d['kernels'], d['roofline'].get('frac'))"
