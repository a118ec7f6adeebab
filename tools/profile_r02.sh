#!/bin/bash
# ncu evidence of one build (run under gpurun; writes gpurun_out/<tag>_*): launch list of the bench command, --set full captures
# of the fused MLP kernel on c3 and c4, and of the HBM-bound per-ray / packing kernels.
#   tools/profile_r02.sh <tag>
set -u
TAG=${1:-r02}
O=gpurun_out
CMD3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-legs"
CMD4="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-legs"
$CMD3 > $O/${TAG}_plain_c3.json 2> $O/${TAG}_plain_c3.err || { echo "plain c3 run failed"; tail -5 $O/${TAG}_plain_c3.err; exit 1; }
$CMD4 > $O/${TAG}_plain_c4.json 2> $O/${TAG}_plain_c4.err || { echo "plain c4 run failed"; exit 1; }
# every launch of the timed c3 steps (warm-up: 3 steps x 3 batches x 6 launches + ray generation)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c3.csv $CMD3 > $O/${TAG}_ncu1.log 2>&1
# fused kernel, c3: the coarse and the fine launch of one 50 000-ray batch (after the 18 warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:mlp_fused -s 18 -c 2 -o $O/${TAG}_fused_c3 $CMD3 > $O/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_fused -s 18 -c 2 -o $O/${TAG}_fused_c4 $CMD4 > $O/${TAG}_ncu3.log 2>&1
# HBM-bound kernels around it
ncu --set full --clock-control none -k 'regex:composite|sample_fine|sample_coarse|pack_level|gen_rays' -s 8 -c 8 -o $O/${TAG}_small_c3 $CMD3 > $O/${TAG}_ncu4.log 2>&1
ls -la $O | tail -12
