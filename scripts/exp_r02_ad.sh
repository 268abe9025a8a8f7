python -m pytest tests/test_gpu_glm.py -q > gpurun_out/r2_pytest27.log 2>&1; tail -5 gpurun_out/r2_pytest27.log
python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-260
