"""
Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference, loaded by
oracle/ref_loader.py with model.models := models.py.backup2) on the seeded synthetic cases of
oracle/synth.py.  Run in the build container only:

    python tests/golden/make_golden.py

Only OUTPUTS of the reference are stored (inputs are regenerated from seeds by oracle/synth.py),
so the fixtures stay small.  Stages captured per case:
  zx        rows fed to ResnetFC (models.py.backup2:243), via a forward-pre-hook on mlp_coarse
  net_out   PixelNeRFNet.forward output for coarse and fine MLP
  render    NeRFRenderer.forward through bind_parallel(...): rgb/depth/weights (coarse+fine),
            plus the z samples the renderer fed to composite (captured by wrapping the bound
            methods; the reference source is not modified)
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_loader, synth  # noqa: E402
from pixel_nerf_multiscale_b200.util.conf import ConfigFactory  # noqa: E402

RENDER_SEED = 123
N_POINTS = 48


def ref_net(case_name, conf, device="cpu"):
    make_model, NeRFRenderer, util = ref_loader.load()
    case = synth.CASES[case_name]
    conf["model"]["encoder"].put("pretrained", False)
    conf["model"]["encoder"].put("use_multi_scale", bool(case["multi_scale"]))
    torch.manual_seed(0)
    net = make_model(conf["model"]).to(device).eval()
    scene, raw = synth.build_case(case_name, conf["model"], device=device)
    assert net.d_in == raw["d_in"], (net.d_in, raw["d_in"])
    # the fork leaves d_latent as the per-level list in multi-scale mode (models.py.backup2:48)
    assert net.latent_size == raw["d_latent"], (net.latent_size, raw["d_latent"])
    net.mlp_coarse.load_state_dict(raw["mlp_coarse"], strict=True)
    net.mlp_fine.load_state_dict(raw["mlp_fine"], strict=True)
    sb, ns, H, W = case["sb"], case["ns"], case["H"], case["W"]
    images = torch.zeros(sb, ns, 3, H, W, device=device)
    with torch.no_grad():
        net.encode(images, raw["poses"], raw["focal"], c=raw["c"])
    # inject the synthetic feature maps (plain attributes in the fork, encoder.py:106-107)
    net.encoder.latent = raw["latents"][-1]
    net.encoder.latents = list(raw["latents"])
    return net, NeRFRenderer, scene, raw


def sample_points(case, sb, n, seed):
    rays = synth.target_rays(case, n, seed, sb)  # (sb,n,8)
    g = torch.Generator().manual_seed(seed + 5)
    t = torch.rand(sb, n, 1, generator=g)
    z = rays[..., 6:7] * (1 - t) + rays[..., 7:8] * t
    return rays[..., :3] + z * rays[..., 3:6], rays[..., 3:6].contiguous()


def run_case(case_name, variants):
    case = synth.CASES[case_name]
    conf = ConfigFactory.parse_file(os.path.join(REPO, case["conf"]))
    net, NeRFRenderer, scene, raw = ref_net(case_name, conf)
    out = {"case": case_name}
    sb = case["sb"]
    xyz, vd = sample_points(case, sb, N_POINTS, 7)
    grabbed = {}
    h = net.mlp_coarse.register_forward_pre_hook(lambda m, a: grabbed.__setitem__("zx", a[0].detach().clone()))
    with torch.no_grad():
        out["net_coarse"] = net(xyz, coarse=True, viewdirs=vd).clone()
        out["zx"] = grabbed["zx"]
        h.remove()
        out["net_fine"] = net(xyz, coarse=False, viewdirs=vd).clone()
    rays = synth.target_rays(case, case["rays"], 3, sb)
    for vname, kw in variants.items():
        rconf = ConfigFactory.from_dict(conf["renderer"].to_dict())
        for k in ("n_coarse", "n_fine", "n_fine_depth", "depth_std"):
            if k in kw:
                rconf.put(k, kw[k])
        renderer = NeRFRenderer.from_conf(rconf, lindisp=kw.get("lindisp", False), eval_batch_size=1500)
        assert bool(renderer.white_bkgd) == bool(case["white_bkgd"])
        zs = []
        orig = renderer.composite

        def spy(model, r, z, coarse=True, sb=0, _orig=orig, _zs=zs):
            _zs.append(z.detach().clone())
            return _orig(model, r, z, coarse=coarse, sb=sb)

        renderer.composite = spy
        par = renderer.bind_parallel(net, None, simple_output=False).eval()
        torch.manual_seed(RENDER_SEED)
        with torch.no_grad():
            res = par(rays, want_weights=True)
        g = {"z_coarse": zs[0]}
        for lvl in ("coarse", "fine"):
            if lvl in res:
                for k in ("rgb", "depth", "weights"):
                    g["%s_%s" % (lvl, k)] = res[lvl][k].detach().clone()
        if len(zs) > 1:
            g["z_fine"] = zs[1]
        g["kw"] = dict(kw)
        out["render_" + vname] = g
    path = os.path.join(HERE, case_name + ".pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


VARIANTS = {
    "ss_ns1": {"default": {}, "lindisp": {"lindisp": True}, "coarse_only": {"n_fine": 0, "n_fine_depth": 0},
               "no_depth": {"n_coarse": 32, "n_fine": 16, "n_fine_depth": 0},
               "video": {"n_coarse": 64, "n_fine": 128, "n_fine_depth": 16},
               # a sample count that is not a power of two: torch.linspace's step is then inexact (pins its fused form)
               "kc48": {"n_coarse": 48, "n_fine": 24, "n_fine_depth": 8}},
    "ms_ns2": {"default": {}},
    "dtu_ns3": {"default": {}},
    "ms_ns3_sb2": {"default": {}},
    "sv3_ns1": {"default": {}},
    "dtu_ns3_s6": {"default": {}},
    "ms_ns2_s6": {"default": {}},
}

if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, var in VARIANTS.items():
        if not only or name in only:
            run_case(name, var)
