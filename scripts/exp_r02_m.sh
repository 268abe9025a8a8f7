set -u
O=gpurun_out
bash scripts/profile_r02.sh
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_c5.json"))
print("c5 value %.4g e2e %.4g ms %.2f sweep_ms %.4f frac %.3f like_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["phases"]["likelihood"]["frac_of_hbm_peak"], d["clocks"]))
for c in ["c1", "c2", "c3", "c4"]:
    d = json.load(open(f"gpurun_out/r02_bench_{c}.json"))
    print(c, "%.3f ms" % d["ms_per_step"], "%.4g" % d["value"], d["kernels"])
PY
