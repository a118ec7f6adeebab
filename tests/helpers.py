"""Shared helpers for the test-suite (test infrastructure; may import oracle/)."""
import os

import torch

from oracle import pixelnerf_oracle as po
from oracle import synth
from pixel_nerf_multiscale_b200.util.conf import ConfigFactory

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
RENDER_SEED = 123
N_POINTS = 48


def load_conf(case_name):
    case = synth.CASES[case_name]
    conf = ConfigFactory.parse_file(os.path.join(REPO, case["conf"]))
    conf["model"]["encoder"].put("pretrained", False)
    conf["model"]["encoder"].put("use_multi_scale", bool(case["multi_scale"]))
    return conf


def load_golden(case_name):
    return torch.load(os.path.join(GOLDEN_DIR, case_name + ".pt"), map_location="cpu")


def sample_points(case, sb, n, seed):
    """Same points as tests/golden/make_golden.py:sample_points."""
    rays = synth.target_rays(case, n, seed, sb)
    g = torch.Generator().manual_seed(seed + 5)
    t = torch.rand(sb, n, 1, generator=g)
    z = rays[..., 6:7] * (1 - t) + rays[..., 7:8] * t
    return rays[..., :3] + z * rays[..., 3:6], rays[..., 3:6].contiguous()


def renderer_kwargs(conf, kw):
    r = conf["renderer"]
    out = dict(n_coarse=r.get_int("n_coarse", 128), n_fine=r.get_int("n_fine", 0),
               n_fine_depth=r.get_int("n_fine_depth", 0), depth_std=r.get_float("depth_std", 0.01),
               white_bkgd=bool(r.get_float("white_bkgd", False)), lindisp=False)
    out.update(kw)
    return out


def maxabs(a, b):
    return (a.double() - b.double()).abs().max().item()


def build_product(case_name, device="cuda", precision="fp32"):
    """This package's PixelNeRFNet + NeRFRenderer for a synthetic case, with the oracle Scene that
    holds the same weights / feature maps / cameras."""
    import pixel_nerf_multiscale_b200 as pk

    conf = load_conf(case_name)
    case = synth.CASES[case_name]
    scene, raw = synth.build_case(case_name, conf["model"], device=device)
    torch.manual_seed(0)
    net = pk.make_model(conf["model"]).to(device).eval()
    net.precision = precision
    assert net.d_in == raw["d_in"] and net.latent_size == raw["d_latent"]
    net.mlp_coarse.load_state_dict(raw["mlp_coarse"], strict=True)
    net.mlp_fine.load_state_dict(raw["mlp_fine"], strict=True)
    sb, ns, H, W = case["sb"], case["ns"], case["H"], case["W"]
    images = torch.zeros(sb, ns, 3, H, W, device=device)
    with torch.no_grad():
        net.encode(images, raw["poses"].to(device), raw["focal"].to(device),
                   c=None if raw["c"] is None else raw["c"].to(device))
    net.encoder.latent = raw["latents"][-1]
    net.encoder.latents = list(raw["latents"])
    net.invalidate_scene()
    return net, conf, scene, raw


def make_renderer(conf, kw, eval_batch_size=50000):
    import pixel_nerf_multiscale_b200 as pk
    from pixel_nerf_multiscale_b200.util.conf import ConfigFactory as CF

    rconf = CF.from_dict(conf["renderer"].to_dict())
    for k in ("n_coarse", "n_fine", "n_fine_depth", "depth_std"):
        if k in kw:
            rconf.put(k, kw[k])
    return pk.NeRFRenderer.from_conf(rconf, lindisp=kw.get("lindisp", False), eval_batch_size=eval_batch_size)
