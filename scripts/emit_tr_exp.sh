for tr in 14 18 22 26; do
MCLCAR_EMIT_TR=$tr python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('TR=$tr', round(d['ms_per_step'],2), round(d['roofline']['launch_kinds']['emitting']['avg_launch_ms'],4), round(d['roofline']['launch_kinds']['emitting']['frac'],3), d['result_check']['oracle_stats_max_rel'])"
done
