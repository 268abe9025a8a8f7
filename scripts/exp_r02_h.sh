set -u
O=gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2_pytest11.log 2>&1; rc=$?; tail -3 $O/r2_pytest11.log
if [ $rc -ne 0 ]; then echo "tests failed: no profile"; exit 1; fi
python __graft_entry__.py smoke > $O/r2_smoke2.log 2>&1; tail -1 $O/r2_smoke2.log | cut -c1-200
CONFIGS="MCLCAR_SWEEP_TR=14" NS=1480 EXTRA="--chains 148" OUT=$O/sweep_exp5.log bash scripts/sweep_exp.sh
bash scripts/profile_r02.sh
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_c5.json"))
print("c5 value %.4g e2e %.4g ms %.2f sweep_ms %.4f frac %.3f like_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["phases"]["likelihood"]["frac_of_hbm_peak"], d["clocks"]))
PY
