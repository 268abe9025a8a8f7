python scripts/exp_glm_site_mixing.py 2>&1 | tail -9
python -m pytest tests/test_gpu_glm.py tests/test_gpu_estimates_scale.py tests/test_gpu_scale_parity.py tests/test_gpu_full_size.py -q > gpurun_out/r2_pytest26.log 2>&1; tail -5 gpurun_out/r2_pytest26.log
python scripts/exp_glm_proposal.py 2>&1 | grep "laplace3\|prior" | tail -16
