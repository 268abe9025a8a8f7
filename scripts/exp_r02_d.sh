set -u
O=gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2_pytest8.log 2>&1; tail -3 $O/r2_pytest8.log
python __graft_entry__.py smoke > $O/r2_smoke.log 2>&1; tail -2 $O/r2_smoke.log
python bench.py --steps 5 --warmup 3 > $O/r02b_bench_c5.json 2> $O/r02b_bench_c5.err
python bench.py --config c3 > $O/r02b_bench_c3.json 2> $O/r02b_bench_c3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02b_bench_c5.json"))
print("c5 value %.4g e2e %.4g ms %.2f e2e_ms %.2f sweep_ms %.4f frac %.3f like_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["phases"]["likelihood"]["frac_of_hbm_peak"], d["clocks"]))
d = json.load(open("gpurun_out/r02b_bench_c3.json"))
print("c3 value %.4g ms %.2f" % (d["value"], d["ms_per_step"]), d["kernels"])
PY
