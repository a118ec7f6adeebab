"""
TEST INFRASTRUCTURE ONLY -- deterministic synthetic inputs for the oracle, the golden
generator, the GPU parity tests, smoke() and bench.py (there is no dataset and no checkpoint;
SURVEY.md section 8d).  Everything is produced from explicit seeded CPU generators so that the build
container (where the goldens are made by running the reference) and the GPU box regenerate
bit-identical inputs.

Weights are NOT the reference's default init: at default init every ResnetBlockFC is the
identity (fc_1.weight is zero-initialised, src/model/resnetfc.py:39) and sigma is ~0, so a
parity test would exercise neither the block GEMMs nor compositing.  All
weights and biases are therefore re-randomised here.
"""
import math

import torch

from . import pixelnerf_oracle as po

# name -> case description.  'levels' are the (C,H,W) of the injected feature maps.
CASES = {
    # tiny cases whose reference outputs are committed under tests/golden/
    "ss_ns1": dict(ns=1, sb=1, H=32, W=32, focal=40.0, c=None, levels=[(256, 9, 11)], z_near=1.2, z_far=4.0,
                   radius=2.6, conf="conf/exp/sn64.conf", multi_scale=False, rays=96, white_bkgd=True),
    "ms_ns2": dict(ns=2, sb=1, H=32, W=32, focal=40.0, c=None,
                   levels=[(64, 24, 40), (64, 24, 40), (128, 12, 20), (256, 6, 10)], z_near=1.2, z_far=4.0,
                   radius=2.6, conf="conf/exp/sn64_multiscale.conf", multi_scale=True, rays=96, white_bkgd=True),
    "dtu_ns3": dict(ns=3, sb=1, H=30, W=40, focal=(72.3, 72.3), c=(20.0, 15.0), levels=[(256, 19, 25)],
                    z_near=0.1, z_far=5.0, radius=2.2, conf="conf/exp/dtu.conf", multi_scale=False, rays=96,
                    white_bkgd=False),
    # the single-view base schema: 3 blocks, never pooled (combine_layer defaults to 1000), lin_z in every block
    "sv3_ns1": dict(ns=1, sb=1, H=32, W=32, focal=40.0, c=None, levels=[(256, 9, 11)], z_near=1.2, z_far=4.0,
                    radius=2.6, conf="conf/default.conf", multi_scale=False, rays=64, white_bkgd=True),
    # the density head of round 1's FIRST fixtures (sigma row x6, bias 1.5: sigma up to ~15-25, closer to a trained
    # pixelNeRF than the O(1) head above), restored as extra cases after VERDICT r1 ("un-soften")
    "dtu_ns3_s6": dict(ns=3, sb=1, H=30, W=40, focal=(72.3, 72.3), c=(20.0, 15.0), levels=[(256, 19, 25)],
                       z_near=0.1, z_far=5.0, radius=2.2, conf="conf/exp/dtu.conf", multi_scale=False, rays=96,
                       white_bkgd=False, sigma_gain=6.0, sigma_bias=1.5),
    "ms_ns2_s6": dict(ns=2, sb=1, H=32, W=32, focal=40.0, c=None,
                      levels=[(64, 24, 40), (64, 24, 40), (128, 12, 20), (256, 6, 10)], z_near=1.2, z_far=4.0,
                      radius=2.6, conf="conf/exp/sn64_multiscale.conf", multi_scale=True, rays=96, white_bkgd=True,
                      sigma_gain=6.0, sigma_bias=1.5),
    "ms_ns3_sb2": dict(ns=3, sb=2, H=30, W=40, focal=(72.3, 72.3), c=(20.0, 15.0),
                       levels=[(64, 15, 20), (64, 15, 20), (128, 8, 10), (256, 4, 5)], z_near=0.5, z_far=4.5,
                       radius=2.2, conf="conf/exp/dtu.conf", multi_scale=True, rays=64, white_bkgd=False),
}


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def mlp_state(seed, d_in, d_latent, d_hidden=512, n_blocks=5, combine_layer=3, d_out=4, sigma_gain=1.0, sigma_bias=1.0):
    """State-dict slice of one ResnetFC (key names as in src/model/resnetfc.py:127-165)."""
    g = _gen(seed)

    def lin(n_out, n_in, gain=1.0, bias_std=0.05):
        w = torch.randn(n_out, n_in, generator=g) * (gain * math.sqrt(2.0 / n_in))
        b = torch.randn(n_out, generator=g) * bias_std
        return w, b

    sd = {}
    sd["lin_in.weight"], sd["lin_in.bias"] = lin(d_hidden, d_in)
    sd["lin_out.weight"], sd["lin_out.bias"] = lin(d_out, d_hidden, gain=0.25)
    # density head: positive offset so that sigma > 0 on most samples and compositing sees
    # pixel opacities between ~0.5 and 1.  The default cases keep the sigma row as drawn
    # (sigma of O(1)); the *_s6 cases scale it by 6 with bias 1.5 (sigma in the tens)
    sd["lin_out.weight"][3] *= sigma_gain
    sd["lin_out.bias"][3] = sigma_bias
    for b in range(n_blocks):
        sd["blocks.%d.fc_0.weight" % b], sd["blocks.%d.fc_0.bias" % b] = lin(d_hidden, d_hidden)
        sd["blocks.%d.fc_1.weight" % b], sd["blocks.%d.fc_1.bias" % b] = lin(d_hidden, d_hidden, gain=0.5)
    if d_latent > 0:
        for b in range(min(combine_layer, n_blocks)):
            sd["lin_z.%d.weight" % b], sd["lin_z.%d.bias" % b] = lin(d_hidden, d_latent, gain=0.7)
    return sd


def feature_levels(seed, n_views, levels):
    """Synthetic post-ReLU-like feature maps, one (V,C,H,W) tensor per level."""
    g = _gen(seed)
    out = []
    for (C, H, W) in levels:
        base = torch.randn(n_views, C, H, W, generator=g)
        # smooth a little along x/y so bilinear weights matter but neighbours differ
        sm = 0.5 * base + 0.25 * torch.roll(base, 1, dims=-1) + 0.25 * torch.roll(base, 1, dims=-2)
        out.append(torch.relu(sm).contiguous())
    return out


def source_poses(n_views, radius, sb=1):
    poses = []
    for o in range(sb):
        for i in range(n_views):
            poses.append(po.pose_spherical(30.0 * i + 47.0 * o, -20.0 - 5.0 * o, radius))
    return torch.stack(poses).reshape(sb, n_views, 4, 4)


def target_rays(case, n_rays, seed, sb=1):
    """n_rays rays per object, picked from a target view at (-63deg, -10deg)."""
    g = _gen(seed)
    H, W = case["H"], case["W"]
    pose = po.pose_spherical(-63.0, -10.0, case["radius"]).unsqueeze(0)
    focal, c = intrinsics(case)
    rays = po.gen_rays(pose, W, H, focal, case["z_near"], case["z_far"], c).reshape(-1, 8)
    pick = torch.randint(0, rays.shape[0], (sb, n_rays), generator=g)
    return rays[pick]  # (sb, n_rays, 8)


def model_hparams(conf_model):
    """Static hyper-parameters the oracle's Scene needs, read from a conf 'model' subtree."""
    code = conf_model.get("code", {})
    mc = conf_model["mlp_coarse"]
    return dict(
        n_blocks=mc.get_int("n_blocks", 5), combine_layer=mc.get_int("combine_layer", 1000),
        combine_type=mc.get_string("combine_type", "average"),
        use_viewdirs=conf_model.get_bool("use_viewdirs", False), use_code=conf_model.get_bool("use_code", False),
        use_code_viewdirs=conf_model.get_bool("use_code_viewdirs", True),
        normalize_z=conf_model.get_bool("normalize_z", True), use_xyz=conf_model.get_bool("use_xyz", False),
        num_freqs=int(code.get("num_freqs", 6)), freq_factor=float(code.get("freq_factor", math.pi)),
        include_input=bool(code.get("include_input", True)),
    )


def d_in_of(hp):
    d = 3 if hp["use_xyz"] else 1
    if hp["use_viewdirs"] and hp["use_code_viewdirs"]:
        d += 3
    if hp["use_code"]:
        d = hp["num_freqs"] * 2 * d + (d if hp["include_input"] else 0)
    if hp["use_viewdirs"] and not hp["use_code_viewdirs"]:
        d += 3
    return d


def intrinsics(case):
    """focal / c in the shapes the reference's callers pass to encode(): a 0-dim tensor for a
    single focal length, (1,2) for (fx,fy) / (cx,cy)  (eval/gen_video.py:93-101,204-209)."""
    f = torch.tensor(case["focal"], dtype=torch.float32)
    if f.dim() == 1:
        f = f[None]
    c = None if case["c"] is None else torch.tensor(case["c"], dtype=torch.float32)[None]
    return f, c


def build_case(name, conf_model, device="cpu", seed=0):
    """Returns (oracle Scene, raw) where raw holds the tensors a PixelNeRFNet needs."""
    case = CASES[name] if isinstance(name, str) else name
    hp = model_hparams(conf_model)
    ns, sb = case["ns"], case["sb"]
    d_latent = sum(l[0] for l in case["levels"])
    d_in = d_in_of(hp)
    lat = [t.to(device) for t in feature_levels(seed + 11, sb * ns, case["levels"])]
    poses = source_poses(ns, case["radius"], sb).to(device)
    focal_in, c_in = intrinsics(case)
    w2c, focal, c = po.encode_cameras(poses.reshape(-1, 4, 4), focal_in, c_in, case["W"], case["H"])
    head = dict(sigma_gain=case.get("sigma_gain", 1.0), sigma_bias=case.get("sigma_bias", 1.0))
    sd_c = {k: v.to(device) for k, v in mlp_state(seed + 1, d_in, d_latent, n_blocks=hp["n_blocks"],
                                                  combine_layer=hp["combine_layer"], **head).items()}
    sd_f = {k: v.to(device) for k, v in mlp_state(seed + 2, d_in, d_latent, n_blocks=hp["n_blocks"],
                                                  combine_layer=hp["combine_layer"], **head).items()}
    scene = po.Scene(lat, w2c.to(device), focal.to(device), c.to(device), ns, sd_c, sd_f, d_latent=d_latent, **hp)
    raw = dict(latents=lat, poses=poses, focal=focal_in, c=c_in,
               mlp_coarse=sd_c, mlp_fine=sd_f, d_in=d_in, d_latent=d_latent, hp=hp, case=case)
    return scene, raw
