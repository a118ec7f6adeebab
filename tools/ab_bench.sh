#!/bin/bash
# A/B timing of library variants on one box: tools/ab_bench.sh "c3 c4" ab/libA.so ab/libB.so ...
# Interleaves the variants (two rounds) so that power-cap drift affects all of them alike.
WL="$1"; shift
for round in 1 2; do
  for so in "$@"; do
    for w in $WL; do
      PIXELNERF_B200_LIB="$PWD/$so" timeout -s KILL 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-legs 2>/dev/null | tail -1 |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', '$w', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'])"
    done
  done
done
