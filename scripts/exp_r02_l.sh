set -u
O=gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2_pytest14.log 2>&1; tail -3 $O/r2_pytest14.log
for c in c3 c4 c1; do python bench.py --config $c 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$c', '%.3f ms/step value %.4g' % (d['ms_per_step'], d['value']), d['kernels'], d['phases']['sampler']['value'])"; done
