set -u
run() {  # label, env...
  env "$@" python bench.py --steps 2 --warmup 1 --samples-per-gpu 1480 --chains 148 --no-cpu-baseline --no-fp64 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('short', '$*', 'launches', r['launches'], 'sampler ms/step %.3f' % d['phases']['sampler']['ms_per_step'], 'site-upd/s %.4g' % d['phases']['sampler']['value'])"
}
run MCLCAR_SWEEP_DOUBLE=0
run MCLCAR_SWEEP_DOUBLE=1
run MCLCAR_SWEEP_DOUBLE=1 MCLCAR_SWEEP2_THREADS=384
run MCLCAR_SWEEP_DOUBLE=1 MCLCAR_SWEEP2_THREADS=416
run MCLCAR_SWEEP_DOUBLE=1 MCLCAR_SWEEP_TR2=18
run MCLCAR_SWEEP_DOUBLE=1 MCLCAR_SWEEP_TR2=18 MCLCAR_SWEEP2_THREADS=416
run MCLCAR_SWEEP_DOUBLE=1 MCLCAR_SWEEP_TR2=16 MCLCAR_SWEEP2_THREADS=384
runf() {
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fp64 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('full', '$*', 'value %.4g step %.2f ms' % (d['value'], d['ms_per_step']), 'sampler ms/step %.3f' % d['phases']['sampler']['ms_per_step'], d['clocks']['sm_mhz'])"
}
runf MCLCAR_SWEEP_DOUBLE=1
runf MCLCAR_SWEEP_DOUBLE=1 MCLCAR_SWEEP2_THREADS=384
runf MCLCAR_SWEEP_DOUBLE=1 MCLCAR_SWEEP_TR2=18 MCLCAR_SWEEP2_THREADS=416
runf MCLCAR_SWEEP_DOUBLE=0
