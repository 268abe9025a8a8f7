set -u
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > $O/r02_bench_8gpu.json 2> $O/r02_bench_8gpu.err; echo "weak8 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_8gpu.json"))
print(d["n_gpus"], "%.4g" % d["value"], "%.2f ms" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d.get("result_check"), d["clocks"])
PY
tail -3 $O/r02_bench_8gpu.err | cut -c1-300
