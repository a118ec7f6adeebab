"""
TEST INFRASTRUCTURE -- error of the bf16 tensor-core path against the fp32 oracle ON THE SAME GPU
as a function of the density scale, at BASELINE.json's full-size configurations (bench workloads).

The synthetic density head (bench.rerandomise / oracle/synth.py) emits sigma of O(1); a trained
pixelNeRF emits sigma in the tens.  `gain` multiplies the sigma row of lin_out (weight and bias) of both
MLPs, so sigma scales by exactly that factor while everything upstream of the head is unchanged.

    python tests/sigma_sweep.py [--workloads c2,c3,c4] [--gains 1,3,6,12] [--rays 8192] [--precision bf16]

prints one JSON line per (workload, gain): sigma statistics, max-abs rgb / depth error of the coarse
pass (identical sample positions: pure arithmetic error) and of the fine pass, the share of rays whose
importance samples landed in a different coarse bin, and the PSNR of both renderers against a common
pseudo ground truth (the oracle's render with other random draws) -- north_star: |dPSNR| <= 0.05 dB.
Used by tests/test_gpu_fullsize.py; the committed table is profiles/r02_sigma_sweep.jsonl.
"""
import argparse
import json
import math
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def make_tape(n, seed, dev, kc=64, kf=32, kd=16):
    g = torch.Generator().manual_seed(seed)
    return dict(coarse=torch.rand(n, kc, generator=g).to(dev), u=torch.rand(n, kf - kd, generator=g).to(dev),
                jitter=torch.rand(n, kf - kd, generator=g).to(dev), normal=torch.randn(n, kd, generator=g).to(dev))


class Replay:
    """Feeds a pre-drawn tape to the oracle's render()."""

    def __init__(self, tape):
        self.t = tape

    def draw_coarse(self):
        return self.t["coarse"]

    def draw_fine(self):
        return self.t.get("u"), self.t.get("jitter"), self.t.get("normal")


def scale_sigma_head(net, gain, bias=None):
    """sigma row of lin_out of both MLPs times `gain` (in place; bumps the parameter versions so the
    native image is re-packed)."""
    with torch.no_grad():
        for m in (net.mlp_coarse, net.mlp_fine):
            if m is None:
                continue
            m.lin_out.weight[3] *= gain
            m.lin_out.bias[3] = (m.lin_out.bias[3] * gain) if bias is None else bias


def psnr(a, b):
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return -10.0 * math.log10(max(mse, 1e-30))


def compare(workload, gain, n_rays, precision="bf16", seed=5, device="cuda:0", frames=1):
    """Renders `n_rays` rays of bench workload `workload` with the product (precision) and with the fp32
    oracle on the same device and the same random draws; returns a dict of error statistics."""
    import bench
    from oracle import pixelnerf_oracle as po

    wl = bench.WORKLOADS[workload]
    dev = torch.device(device)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    net, renderer, conf, cam = bench.build_scene(wl, dev, precision)
    scale_sigma_head(net, gain)
    rays_all = bench.orbit_rays(wl, cam, frames, dev)
    pick = torch.randperm(rays_all.shape[0], generator=torch.Generator().manual_seed(3))[:n_rays].to(dev)
    rays = rays_all[pick].contiguous()
    n = rays.shape[0]
    r = conf["renderer"]
    kw = dict(n_coarse=r.get_int("n_coarse", 128), n_fine=r.get_int("n_fine", 0), n_fine_depth=r.get_int("n_fine_depth", 0),
              depth_std=r.get_float("depth_std", 0.01), white_bkgd=bool(r.get_float("white_bkgd", False)), lindisp=False)
    tape = make_tape(n, seed, dev, kw["n_coarse"], kw["n_fine"], kw["n_fine_depth"])
    gscene = bench.oracle_scene(net, cam, conf, device=dev)
    with torch.no_grad():
        renderer.rng_tape = dict(tape)
        ours = renderer(net, rays[None], want_weights=True, taps=True)
        from pixel_nerf_multiscale_b200 import _native as N

        N.check(N.lib().pnr_tc_check(N.stream_ptr(dev)), "pnr_tc_check")
        ref = po.render(gscene, rays[None], tape=Replay(tape), eval_batch_size=200000, **kw)
        other = po.render(gscene, rays[None], tape=Replay(make_tape(n, seed + 1000, dev, kw["n_coarse"], kw["n_fine"],
                                                                    kw["n_fine_depth"])), eval_batch_size=200000, **kw)
    torch.cuda.synchronize(dev)
    sig = ref["fine"]["out"][..., 3]
    res = {"workload": workload, "gain": gain, "rays": n, "precision": precision,
           "sigma_max": sig.max().item(), "sigma_mean": sig.mean().item(),
           "sigma_p99": sig.flatten().float().kthvalue(int(sig.numel() * 0.99))[0].item(),
           "opacity_mean": ref["fine"]["weights"].sum(-1).mean().item()}
    for lvl in ("coarse", "fine"):
        e_rgb = (ours[lvl].rgb[0] - ref[lvl]["rgb"]).abs()
        e_d = (ours[lvl].depth[0] - ref[lvl]["depth"]).abs()
        res[lvl + "_rgb_max"] = e_rgb.max().item()
        res[lvl + "_depth_max"] = e_d.max().item()
        res[lvl + "_rgb_p999"] = e_rgb.flatten().kthvalue(max(1, int(e_rgb.numel() * 0.999)))[0].item()
        res[lvl + "_depth_p999"] = e_d.flatten().kthvalue(max(1, int(e_d.numel() * 0.999)))[0].item()
    res["coarse_z_bit_equal"] = bool(torch.equal(ours.coarse.z[0], ref["coarse"]["z"]))
    dz = (ours.fine.z[0] - ref["fine"]["z"]).abs()
    step = (wl["z_far"] - wl["z_near"]) / kw["n_coarse"]
    res["bin_flip_rays"] = (dz.max(dim=-1)[0] > 0.5 * step).float().mean().item()
    res["psnr_ours_vs_ref"] = psnr(ours.fine.rgb[0], ref["fine"]["rgb"])
    res["psnr_ours_vs_gt"] = psnr(ours.fine.rgb[0], other["fine"]["rgb"])
    res["psnr_ref_vs_gt"] = psnr(ref["fine"]["rgb"], other["fine"]["rgb"])
    res["dpsnr"] = res["psnr_ours_vs_gt"] - res["psnr_ref_vs_gt"]
    del gscene, net
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c2,c3,c4")
    ap.add_argument("--gains", default="1,3,6,12")
    ap.add_argument("--rays", type=int, default=8192)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    lines = []
    for w in a.workloads.split(","):
        for g in a.gains.split(","):
            res = compare(w, float(g), a.rays, a.precision)
            print(json.dumps(res), flush=True)
            lines.append(res)
    if a.out:
        with open(a.out, "w") as f:
            for res in lines:
                f.write(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
