set -u
O=gpurun_out
python -m pytest tests/test_gpu_glm.py tests/test_gpu_scale_parity.py tests/test_gpu_scotcancer.py tests/test_gpu_callers.py -q -x --timeout 900 -k "glm or config3 or scot or beta or Exp" > $O/r2_pytest7.log 2>&1; tail -2 $O/r2_pytest7.log
for ch in 11 15 22 30 45 60; do MCLCAR_GLM_CHUNKS=$ch python bench.py --config c3 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('chunks $ch', '%.2f ms/step' % d['ms_per_step'], 'glm_terms %.3f ms' % d['kernels']['glm_terms']['ms_per_step'])"; done
python bench.py --config c3 --steps 1 --warmup 1 > /dev/null 2>&1 && ncu --clock-control none --set full --import-source on -k regex:glm_terms_kernel --launch-skip 1 -c 2 -f -o $O/r02_prof_glm_terms3 python bench.py --config c3 --steps 1 --warmup 1 > $O/ncu_glm3.log 2>&1; tail -2 $O/ncu_glm3.log | cut -c1-200
