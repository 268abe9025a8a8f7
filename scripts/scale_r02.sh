# weak and strong scaling lines under torchrun (NG GPUs of one box, default 8); the bench checks the NCCL-merged results itself
set -u
O=gpurun_out
G=${NG:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --steps 3 --warmup 3 > $O/r02_bench_${G}gpu.json 2> $O/r02_bench_${G}gpu.err; echo "weak$G rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $G --steps 2 --warmup 3 --scaling strong --total-samples 100000 > $O/r02_bench_${G}gpu_strong.json 2> $O/r02_bench_${G}gpu_strong.err; echo "strong$G rc=$?"
python - <<PY
import json
for f in ["r02_bench_${G}gpu", "r02_bench_${G}gpu_strong"]:
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["n_gpus"], "%.4g" % d["value"], "%.2f ms" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d.get("scaling"), d.get("result_check"), d["clocks"])
PY
if [ "$G" = "2" ]; then python -m pytest tests/test_gpu_multi.py -q --timeout 600 2>&1 | tail -2; fi
