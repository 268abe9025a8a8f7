set -u
O=gpurun_out
small="--no-cpu-baseline --no-fp64 --steps 1 --warmup 1"
for c in c3 c4; do
  if python bench.py --config $c $small > /dev/null 2>&1; then
    ncu --clock-control none --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/r02_launches_$c.csv python bench.py --config $c $small > /dev/null 2>&1; echo "$c rc=$?"
  fi
done
