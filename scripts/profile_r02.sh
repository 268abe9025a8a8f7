#!/bin/bash
# Round-2 evidence run (one B200): per-config bench lines, ncu launch lists (gpu__time_duration only), and one
# `ncu --set full` capture of each hot kernel.  Every profiled command first runs plain and must exit 0.
# Outputs under gpurun_out/r02_*; summaries are made here afterwards with scripts/ncu_summary.py and copied to profiles/.
set -u
O=gpurun_out
NCU="ncu --clock-control none"
small="--no-cpu-baseline --no-fp64 --steps 1 --warmup 1"
run() { echo "== $*" >> $O/r02_profile.log; "$@" >> $O/r02_profile.log 2>&1; echo "rc=$?" >> $O/r02_profile.log; }
: > $O/r02_profile.log

# 1. bench lines (numbers never come from a profiled run)
python bench.py --steps 5 --warmup 3 > $O/r02_bench_c5.json 2> $O/r02_bench_c5.err
python bench.py --steps 5 --warmup 3 --stats-only --no-cpu-baseline > $O/r02_bench_c5_statsonly.json 2> $O/r02_bench_c5_statsonly.err
MCLCAR_SWEEP_DOUBLE=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64 > $O/r02_bench_c5_single.json 2> $O/r02_bench_c5_single.err
for c in c1 c2 c3 c4; do python bench.py --config $c > $O/r02_bench_$c.json 2> $O/r02_bench_$c.err; done
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err

# 2. launch lists
for c in c5 c3 c4 c2; do
  extra=""; [ $c = c5 ] && extra="--samples-per-gpu 2980"
  if python bench.py --config $c $small $extra > /dev/null 2>> $O/r02_profile.log; then
    run $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/r02_launches_$c.csv python bench.py --config $c $small $extra
  fi
done

# 3. full captures, one or two launches per kernel
full() {  # name config kernel-regex skip count [extra bench flags]
  local name=$1 cfg=$2 k=$3 skip=$4 cnt=$5; shift 5
  run $NCU --set full --import-source on -k regex:$k --launch-skip $skip -c $cnt -f -o $O/r02_prof_$name python bench.py --config $cfg $small "$@"
}
full sweep        c5 'gibbs_torus_sweep_fused' 40 2 --samples-per-gpu 2980     # launch 40 = a two-sweep launch, 41 = an emitting sweep
MCLCAR_SWEEP_DOUBLE=0 full sweep_single c5 'gibbs_torus_sweep_fused' 40 2 --samples-per-gpu 2980   # single-sweep launches: plain, emitting
full sweep_stats  c5 'gibbs_torus_sweep_fused' 17 1 --samples-per-gpu 2980 --stats-only   # 6 burn-in launches, then (double, double, emit) per kept state
full stats_torus  c5 'stats_torus_kernel'       1 1 --samples-per-gpu 2980
full k3           c5 'k3_moments'               1 1 --samples-per-gpu 2980
full csr_sampler  c3 'gibbs_csr_kernel'         0 1
full glm_terms    c3 'glm_terms_kernel'         1 2
full stats_sm     c3 'stats_site_major_kernel'  0 1
full k3_glm       c3 'k3_'                      2 3
full hcar         c4 'gibbs_csr_kernel|stats_site_major_kernel|hcar_combine_kernel' 0 4
full c2_sweep     c2 'gibbs_torus'              20 1
ls -la $O | grep r02_ >> $O/r02_profile.log
