"""
Randomised sweep over the conf schema the path reads: feature / code switches (use_xyz, normalize_z,
use_code, num_freqs, include_input, freq_factor, use_viewdirs, use_code_viewdirs), MLP shape (n_blocks,
combine_layer incl. "never pooled"), encoder width (num_layers 3/4/5, single / multi-scale: d_latent 128 .. 1024),
views per object 1..16, objects per call 1..2, sampling (odd n_coarse, n_fine 0 / <= 32 / > 32, depth samples,
lindisp, white background) and tiny ray counts (1 ray: fewer tiles than cluster pairs).  Every configuration is
rendered by the fp32 validation path (bar 1e-4) and by the tensor-core path with f16 operands (bar 1e-2 at
identical sample positions) and compared with the oracle on the same device and the same random draws.
The seeds are fixed: the sweep is deterministic.
"""
import os
import random

import pytest
import torch

from helpers import load_conf, maxabs
from oracle import pixelnerf_oracle as po
from oracle import synth
from sigma_sweep import Replay

pytestmark = pytest.mark.gpu
N_CONFIGS = int(os.environ.get("PNR_FUZZ_CONFIGS", "48"))   # a longer one-off sweep: PNR_FUZZ_CONFIGS=300


def _draw(i):
    r = random.Random(1000 + i)
    use_code = r.random() < 0.8
    use_viewdirs = r.random() < 0.7
    m = dict(use_xyz=r.random() < 0.75, normalize_z=r.random() < 0.6, use_code=use_code, use_viewdirs=use_viewdirs,
             use_code_viewdirs=use_viewdirs and r.random() < 0.5,
             code=dict(num_freqs=r.choice([2, 4, 6, 10]), include_input=r.random() < 0.7, freq_factor=r.choice([1.5, 3.14159265])))
    n_blocks = r.choice([1, 2, 3, 5, 6, 8])
    ns = r.choice([1, 1, 2, 3, 3, 4, 5, 6, 8, 11, 16])
    combine = r.randint(1, n_blocks) if n_blocks > 1 else 1
    if ns == 1 and r.random() < 0.4:
        combine = 1000                       # the single-view base schema: never pooled
    if combine >= n_blocks and ns > 1:
        combine = n_blocks - 1 if n_blocks > 1 else 1
        n_blocks = max(n_blocks, combine + 1)   # multi-view rows need a pooling point inside the network
    num_layers = r.choice([3, 4, 4, 5])
    multi = r.random() < 0.5
    chans = [64, 64, 128, 256, 512][:num_layers]
    sizes = [(r.randint(5, 14), r.randint(5, 14)) for _ in chans]
    levels = [(c, h, w) for c, (h, w) in zip(chans, sizes)] if multi else [(chans[-1],) + sizes[-1]]
    sb = r.choice([1, 1, 2])
    kc = r.choice([17, 32, 48, 64, 100])
    kf = r.choice([0, 8, 32, 40])
    kd = 0 if kf == 0 else r.choice([0, kf // 2, kf])
    case = dict(ns=ns, sb=sb, H=24, W=32, focal=38.0, c=None if r.random() < 0.5 else (15.0, 13.0), levels=levels,
                z_near=1.0, z_far=3.6, radius=2.4, multi_scale=multi, rays=r.choice([1, 3, 40, 97, 300]), white_bkgd=r.random() < 0.5,
                sigma_gain=r.choice([1.0, 4.0]), sigma_bias=1.2)
    rend = dict(n_coarse=kc, n_fine=kf, n_fine_depth=kd, lindisp=r.random() < 0.3)
    return m, dict(n_blocks=n_blocks, combine_layer=combine), dict(num_layers=num_layers, use_multi_scale=multi), case, rend


def _build(i, precision):
    import pixel_nerf_multiscale_b200 as pk
    from pixel_nerf_multiscale_b200.util.conf import ConfigFactory

    m, mlp, enc, case, rend = _draw(i)
    conf = load_conf("dtu_ns3")
    model = conf["model"]
    for k, v in m.items():
        if k == "code":
            for kk, vv in v.items():
                model["code"].put(kk, vv)
        else:
            model.put(k, v)
    for which in ("mlp_coarse", "mlp_fine"):
        for k, v in mlp.items():
            model[which].put(k, v)
    for k, v in enc.items():
        model["encoder"].put(k, v)
    scene, raw = synth.build_case(case, model, device="cuda", seed=i)
    torch.manual_seed(0)
    net = pk.make_model(model).to("cuda").eval()
    net.precision = precision
    assert net.d_in == raw["d_in"] and net.latent_size == raw["d_latent"], (net.d_in, raw["d_in"], net.latent_size, raw["d_latent"])
    net.mlp_coarse.load_state_dict(raw["mlp_coarse"], strict=True)
    net.mlp_fine.load_state_dict(raw["mlp_fine"], strict=True)
    with torch.no_grad():
        net.encode(torch.zeros(case["sb"], case["ns"], 3, case["H"], case["W"], device="cuda"), raw["poses"].cuda(),
                   raw["focal"].cuda(), c=None if raw["c"] is None else raw["c"].cuda())
    net.encoder.latent = raw["latents"][-1]
    net.encoder.latents = list(raw["latents"])
    net.invalidate_scene()
    rconf = ConfigFactory.from_dict(conf["renderer"].to_dict())
    for k in ("n_coarse", "n_fine", "n_fine_depth"):
        rconf.put(k, rend[k])
    rconf.put("white_bkgd", case["white_bkgd"])
    renderer = pk.NeRFRenderer.from_conf(rconf, lindisp=rend["lindisp"])
    kw = dict(n_coarse=rend["n_coarse"], n_fine=rend["n_fine"], n_fine_depth=rend["n_fine_depth"], depth_std=0.01,
              white_bkgd=bool(case["white_bkgd"]), lindisp=rend["lindisp"])
    rays = synth.target_rays(case, case["rays"], 3 + i, case["sb"]).cuda()
    return net, renderer, scene, case, kw, rays


def _tape(n, kw, seed):
    g = torch.Generator().manual_seed(seed)
    t = {"coarse": torch.rand(n, kw["n_coarse"], generator=g).cuda()}
    n_imp = kw["n_fine"] - kw["n_fine_depth"]
    if n_imp > 0:
        t["u"], t["jitter"] = torch.rand(n, n_imp, generator=g).cuda(), torch.rand(n, n_imp, generator=g).cuda()
    if kw["n_fine_depth"] > 0:
        t["normal"] = torch.randn(n, kw["n_fine_depth"], generator=g).cuda()
    return t


@pytest.mark.parametrize("i", range(N_CONFIGS))
def test_random_config(i):
    from pixel_nerf_multiscale_b200 import _native as N

    torch.backends.cuda.matmul.allow_tf32 = False
    for precision, tol in (("fp32", 1e-4), ("fp16", 1e-2)):
        net, renderer, scene, case, kw, rays = _build(i, precision)
        n = rays.shape[0] * rays.shape[1]
        tape = _tape(n, kw, 50 + i)
        with torch.no_grad():
            renderer.rng_tape = dict(tape)
            res = renderer(net, rays, want_weights=True, taps=True)
            N.check(N.lib().pnr_tc_check(N.stream_ptr(rays.device)), "pnr_tc_check")
            ref = po.render(scene, rays, tape=Replay(tape), **kw)
        flat = rays.reshape(-1, 8)
        assert torch.equal(res.coarse.z.reshape(n, -1), ref["coarse"]["z"]), "coarse sample positions must be bit-equal"
        e = [maxabs(res.coarse.rgb.reshape(-1, 3), ref["coarse"]["rgb"]), maxabs(res.coarse.depth.reshape(-1), ref["coarse"]["depth"])]
        if kw["n_fine"] > 0:
            K = res.fine.z.shape[-1]
            with torch.no_grad():  # the oracle on the product's own fine samples: pure arithmetic error
                _w, rgb_s, d_s, _o = po.composite(scene, flat, res.fine.z.reshape(n, K).contiguous(), False, case["sb"], kw["white_bkgd"])
            e += [maxabs(res.fine.rgb.reshape(-1, 3), rgb_s), maxabs(res.fine.depth.reshape(-1), d_s)]
            zf = res.fine.z.reshape(n, K)
            assert (zf[:, 1:] >= zf[:, :-1]).all()
            if precision == "fp32":  # sample placement itself: equal up to rare last-bit bin flips
                same = ((zf - ref["fine"]["z"]).abs().max(dim=-1)[0] < 1e-5).float().mean().item()
                assert same > 0.9, same
        print("config %d [%s] ns=%d sb=%d rays=%d L=%d blocks=%d/%s Kc=%d Kf=%d/%d: max err %s" % (
            i, precision, case["ns"], case["sb"], case["rays"], net.latent_size, net.mlp_coarse.n_blocks,
            net.mlp_coarse.combine_layer, kw["n_coarse"], kw["n_fine"], kw["n_fine_depth"], ["%.1e" % x for x in e]))
        assert all(torch.isfinite(t).all() for t in (res.coarse.rgb, res.coarse.depth))
        assert max(e) < tol, (precision, e)
