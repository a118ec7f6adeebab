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


# ---- synthetic on-disk datasets (tests of the data adapters and of the reference's drivers) -------------------
def write_srn_fixture(root, name="cars", stage="test", n_obj=2, n_views=4, size=32, focal=40.0, radius=1.3, seed=0):
    """SRN-format folder ``<root>/<name>_<stage>/obj_k/{intrinsics.txt, rgb/*.png, pose/*.txt}`` with
    seeded random images and spherical poses.  Returns the list of expected items."""
    import cv2
    import numpy as np

    import pixel_nerf_multiscale_b200 as pk

    g = np.random.RandomState(seed)
    flip = np.diag([1.0, -1.0, -1.0, 1.0]).astype(np.float32)
    base = os.path.join(root, "%s_%s" % (name, stage))
    expect = []
    for k in range(n_obj):
        d = os.path.join(base, "obj_%02d" % k)
        os.makedirs(os.path.join(d, "rgb"), exist_ok=True)
        os.makedirs(os.path.join(d, "pose"), exist_ok=True)
        with open(os.path.join(d, "intrinsics.txt"), "w") as f:
            f.write("%f %f %f 0.\n0. 0. 0.\n1.\n%d %d\n" % (focal, size / 2.0, size / 2.0, size, size))
        imgs, poses = [], []
        for v in range(n_views):
            img = g.randint(0, 255, size=(size, size, 3)).astype(np.uint8)  # never pure white: all foreground
            cv2.imwrite(os.path.join(d, "rgb", "%06d.png" % v), img[..., ::-1])
            pose = pk.util.pose_spherical(40.0 * v + 17.0 * k, -15.0, radius).numpy()
            np.savetxt(os.path.join(d, "pose", "%06d.txt" % v), (pose @ flip).reshape(1, 16))
            imgs.append(img)
            poses.append(pose)
        expect.append(dict(path=d, images=np.stack(imgs), poses=np.stack(poses), focal=focal, c=(size / 2.0, size / 2.0)))
    return expect
