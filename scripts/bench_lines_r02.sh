python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; echo rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_c5.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print({k:r[k] for k in ('kernel','achieved','frac','algorithmic_bytes_per_launch','avg_launch_ms','launches','sweeps_per_launch','share_of_step','traffic')})
print(r['launch_kinds'])
print(r['all_sweep_launches'])
print(d['phases']['sampler'], d['phases']['emit'], d['clocks'])
PY
python bench.py --steps 5 --warmup 3 --stats-only --no-cpu-baseline --no-fp64 > gpurun_out/r02_bench_c5_statsonly.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c5_statsonly.json')); print('statsonly', d['value'], d['roofline']['frac'], d['roofline']['launch_kinds'].get('emitting',{}).get('frac'), d['roofline']['all_sweep_launches']['frac'])"
MCLCAR_SWEEP_DOUBLE=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r02_bench_c5_single.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c5_single.json')); print('single', d['value'], d['roofline']['frac'], d['roofline']['launch_kinds'].get('emitting',{}).get('frac'), d['roofline']['all_sweep_launches']['frac'], d['clocks'])"
python bench.py --config c2 > gpurun_out/r02_bench_c2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2.json')); print('c2', d['ms_per_step'], d['roofline'].get('frac'), d['kernels'])"
