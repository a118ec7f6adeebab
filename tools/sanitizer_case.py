"""Small render through the public API for compute-sanitizer (memcheck / racecheck) on pools where the tool is open:
    PNR_WAIT_TIMEOUT_MS=0 compute-sanitizer --tool memcheck python tools/sanitizer_case.py
(closed on this pool in rounds 1 and 2: `gpurun` answers rc 86.)"""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from helpers import build_product, make_renderer, renderer_kwargs
from oracle import synth
net, conf, scene, raw = build_product("dtu_ns3", device="cuda:0", precision="fp16")
case = synth.CASES["dtu_ns3"]
rays = synth.target_rays(case, 48, 3, 1).to("cuda:0")
renderer = make_renderer(conf, {})
par = renderer.bind_parallel(net, [0], simple_output=True).eval()
with torch.no_grad():
    rgb, depth = par(rays)
torch.cuda.synchronize()
print("ok", float(rgb.mean()), float(depth.mean()))
