#!/bin/bash
# power while streaming from L2
for which in 0 1 2; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 100 > gpurun_out/pw_$which.csv &
  SMI=$!
  sleep 0.5
  timeout -s KILL 60 ab/l2probe 4864 20000 $which
  kill $SMI
  sort -t, -k2 -n gpurun_out/pw_$which.csv | awk -F, '{a[NR]=$2; c[NR]=$1} END{print "power W: median", a[int(NR/2)], "max", a[NR], " clock at max", c[NR], " samples", NR}'
done
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits
for mp in 74 66 56 44; do
  PNR_MAX_PAIRS=$mp timeout -s KILL 200 python bench.py --workload c2 --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pairs', $mp, round(d['value']), d['clocks'])"
done
