"""Per-role cycle counters of the phase-A kernel.  Needs a library built with
PNR_EXTRA_NVCC_FLAGS=-DPNR_TC_STATS=1 pixel_nerf_multiscale_b200/csrc/build.sh"""
import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import torch, bench
from pixel_nerf_multiscale_b200 import _native as N
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv)>1 else "c2"]
dev = torch.device("cuda:0")
net, renderer, conf, cam = bench.build_scene(wl, dev, "bf16")
par = renderer.bind_parallel(net, [0], simple_output=True).eval()
rays = bench.orbit_rays(wl, cam, 2, dev)[:50000].contiguous()
lib = N.lib()
stats = torch.zeros(74*20, dtype=torch.int64, device=dev)
fn = lib.pnr_tc_debug_stats; fn.argtypes=[C.c_void_p]; fn.restype=C.c_int
with torch.no_grad():
    par(rays[None]); torch.cuda.synchronize()
    fn(N.ptr(stats))
    t0=torch.cuda.Event(enable_timing=True); t1=torch.cuda.Event(enable_timing=True)
    t0.record(); par(rays[None]); t1.record(); torch.cuda.synchronize()
    print("step ms", t0.elapsed_time(t1))
extra = stats.cpu()[74*16:74*17].double(); extra2 = stats.cpu()[74*17:74*18].double(); ntile = stats.cpu()[74*18:74*19].double(); s = stats.cpu()[:74*16].reshape(74,16).double()
names = ["mma_total","mma_wait_bfull","mma_wait_afull","mma_wait_sx","mma_wait_h","mma_wait_xp","prod_total","prod_wait_bempty","prod_wait_aempty","epi_total","epi_wait_xready","epi_wait_net","epi_R_work","epi_H_work","epi_pool_work","commit_to_empty_mean"]
m = s.mean(0)
for n,v in zip(names,m): print("%-18s %12.0f  (%.1f%% of mma_total)" % (n, v, 100*v/max(m[0],1)))
print("sum issue->full-observed latency per pair (mean):", extra.mean().item())
print("pool: cycles in tcgen05.ld+wait per pair (mean):", extra2.mean().item())
print("tiles per pair (mean):", ntile.mean().item(), " pool cycles/tile:", (m[14]/ntile.mean()).item(), " mma_total/tile:", (m[0]/ntile.mean()).item())
