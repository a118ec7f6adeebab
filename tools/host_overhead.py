"""Host-side cost of one render_par(rays) call (no device sync inside the loop): wall time per call while the GPU
queue is kept short (4 096-ray batches of c1), and a cProfile breakdown.  Run on a GPU box."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c1"]
dev = torch.device("cuda:0")
net, renderer, conf, cam = bench.build_scene(wl, dev, "fp16")
par = renderer.bind_parallel(net, [0], simple_output=True).eval()
rays = bench.orbit_rays(wl, cam, 1, dev)[:4096].contiguous()[None]
with torch.no_grad():
    for _ in range(5):
        par(rays)
    torch.cuda.synchronize()
    for n in (1, 4):
        t0 = time.perf_counter()
        for _ in range(50):
            for _ in range(n):
                par(rays)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
        print("%d call(s) + sync: %.1f us per group" % (n, dt * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        par(rays)
        if _ % 4 == 3:
            torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print(s.getvalue()[:5000])
