# two-GPU validation: one-process multi-GPU tests, then the bench under torchrun (weak and strong scaling) -- its
# result_check compares the NCCL-merged (lr, var) with a host merge of the per-rank tuples
set -u
O=gpurun_out
python -m pytest tests/test_gpu_multi.py -q --timeout 900 > $O/r2_pytest_multi.log 2>&1; tail -3 $O/r2_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_bench_2gpu.json 2> $O/r02_bench_2gpu.err; echo "weak rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 2 --warmup 3 --scaling strong --total-samples 25000 > $O/r02_bench_2gpu_strong.json 2> $O/r02_bench_2gpu_strong.err; echo "strong rc=$?"
python bench.py --steps 2 --warmup 3 --scaling strong --total-samples 25000 --no-cpu-baseline > $O/r02_bench_1gpu_strong.json 2> $O/r02_bench_1gpu_strong.err; echo "strong1 rc=$?"
python bench.py --config c3 > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err
python - <<'PY'
import json
for f in ["r02_bench_2gpu", "r02_bench_2gpu_strong", "r02_bench_1gpu_strong", "r02_bench_c3"]:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["n_gpus"], "%.4g" % d["value"], "%.2f ms" % d["ms_per_step"], d.get("scaling"), d.get("result_check"), d.get("kernels"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -5 $O/r02_bench_2gpu.err
