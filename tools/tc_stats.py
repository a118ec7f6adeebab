"""Per-role cycle counters of the fused gather + ResnetFC kernel (one 50 000-ray batch).  Needs a library built with
PNR_EXTRA_NVCC_FLAGS=-DPNR_TC_STATS=1 pixel_nerf_multiscale_b200/csrc/build.sh (the waits are then timed with clock64)."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
from pixel_nerf_multiscale_b200 import _native as N
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
dev = torch.device("cuda:0")
net, renderer, conf, cam = bench.build_scene(wl, dev, prec)
par = renderer.bind_parallel(net, [0], simple_output=True).eval()
rays = bench.orbit_rays(wl, cam, 2, dev)[:50000].contiguous()
lib = N.lib()
pairs = torch.cuda.get_device_properties(dev).multi_processor_count // 2
stats = torch.zeros(pairs * 16, dtype=torch.int64, device=dev)
with torch.no_grad():
    par(rays[None]); torch.cuda.synchronize()
    lib.pnr_tc_debug_stats(N.ptr(stats))
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); par(rays[None]); t1.record(); torch.cuda.synchronize()
    lib.pnr_tc_debug_stats(None)
    print("step ms", t0.elapsed_time(t1), "(counters below: the LAST launch = the fine pass)")
s = stats.cpu().reshape(pairs, 16).double()
names = ["mma_total", "mma_wait_bfull", "mma_wait_afull", "mma_wait_sx", "mma_wait_h", "mma_wait_xp", "prod_total",
         "prod_wait_bempty", "prod_wait_aempty", "epi_total", "epi_wait_xready", "epi_wait_net", "epi_X_work", "epi_H_work",
         "epi_pool_work", "commit_to_empty_mean"]
m = s.mean(0)
for n, v in zip(names, m):
    print("%-18s %14.0f  (%.1f%% of mma_total)" % (n, v, 100 * v / max(m[0], 1)))
