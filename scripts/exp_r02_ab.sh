python -m pytest tests/test_gpu_glm.py -q -k "both_proposals" > gpurun_out/r2_pytest25.log 2>&1; tail -40 gpurun_out/r2_pytest25.log
