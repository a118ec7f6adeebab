#!/bin/bash
# Quick A/B of a kernel change (run under gpurun): parity subset, c3/c4 bench, tensor-pipe activity of the fused kernel.
#   tools/quick_perf.sh <tag>
TAG=${1:-x}
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_parity.py tests/test_gpu_fp32_parity.py -m gpu -q -x 2>&1 | tail -2
for w in c3 c4; do
  python bench.py --workload $w --no-cpu-baseline --no-legs > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
done
python - <<PY
import json
for w in ("c3", "c4"):
    d = json.load(open("$O/${TAG}_bench_%s.json" % w))
    r = d["roofline"]
    print(w, "rays/s %.0f e2e %.0f | fused %.1f TF, %.3f ms/launch, sm %.0f MHz -> %.1f Mcycles/launch" % (
        d["value"], d["e2e"]["value"], r["achieved"], r["avg_launch_ms"], d["clocks"]["sm_mhz"], r["avg_launch_ms"] * d["clocks"]["sm_mhz"] / 1e3))
PY
CMD="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-legs"
$CMD > /dev/null 2>&1 && ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.avg --clock-control none -k regex:mlp_fused -s 18 -c 2 --csv --log-file $O/${TAG}_tensor.csv $CMD > /dev/null 2>&1
grep -v "^==" $O/${TAG}_tensor.csv | cut -d, -f13,15 | tail -6
