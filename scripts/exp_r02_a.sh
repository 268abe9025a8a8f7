set -u
O=gpurun_out
python -m pytest tests/test_gpu_glm.py tests/test_gpu_scale_parity.py tests/test_gpu_sampler_dcar.py tests/test_gpu_scotcancer.py tests/test_gpu_callers.py tests/test_gpu_surface.py -q -x --timeout 900 > $O/r2_pytest5.log 2>&1; tail -3 $O/r2_pytest5.log
scripts/ubench/fp64_peaks > $O/r2_fp64_peaks2.log 2>&1
python bench.py --config c3 > $O/r2_bench_c3_tab.json 2> $O/r2_bench_c3_tab.err
CONFIGS="MCLCAR_SWEEP_CTAS=3;MCLCAR_SWEEP_CTAS=4;MCLCAR_SWEEP_CTAS=4 MCLCAR_SWEEP_THREADS=192;MCLCAR_SWEEP_CTAS=4 MCLCAR_SWEEP_THREADS=224;MCLCAR_SWEEP_CTAS=4 MCLCAR_SWEEP_TR=8;MCLCAR_SWEEP_CTAS=3 MCLCAR_SWEEP_TR=12" NS=1480 EXTRA="--chains 148" OUT=$O/sweep_exp2.log bash scripts/sweep_exp.sh
cat $O/r2_fp64_peaks2.log
python -c "
import json; d=json.load(open('$O/r2_bench_c3_tab.json')); print(d['value'], d['ms_per_step'], d['kernels'])"
