"""bind_parallel(net, [0, 1, ...]) -- single-process multi-device renderer (needs >= 2 GPUs)."""
import pytest
import torch

from oracle import synth
from helpers import build_product, make_renderer, maxabs

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_multi_device_matches_single_device():
    net, conf, scene, raw = build_product("dtu_ns3", precision="fp32")
    case = synth.CASES["dtu_ns3"]
    rays = synth.target_rays(case, 90, 3, 1).cuda()
    renderer = make_renderer(conf, {})
    B = rays.shape[1]
    g = torch.Generator().manual_seed(4)
    tape = dict(coarse=torch.rand(B, 64, generator=g), u=torch.rand(B, 16, generator=g), jitter=torch.rand(B, 16, generator=g),
                normal=torch.randn(B, 16, generator=g))
    single = renderer.bind_parallel(net, [0], simple_output=True).eval()
    renderer.rng_tape = {k: v.cuda() for k, v in tape.items()}
    rgb1, d1 = single(rays)
    multi = renderer.bind_parallel(net, list(range(torch.cuda.device_count())), simple_output=True).eval()
    # a pre-drawn tape is split per shard: the sharded render equals the single-device one exactly (rays are independent)
    renderer.rng_tape = {k: v.cuda() for k, v in tape.items()}
    rgbm, dm = multi(rays)
    assert rgbm.shape == rgb1.shape and dm.shape == d1.shape and rgbm.device == rgb1.device
    assert torch.equal(rgbm, rgb1) and torch.equal(dm, d1)
    # without a tape every device draws its own randoms: same scene, different jitter
    rgbm, dm = multi(rays)
    assert torch.isfinite(rgbm).all() and torch.isfinite(dm).all()
    assert maxabs(rgbm, rgb1) < 0.2
    full = renderer.bind_parallel(net, [0, 1], simple_output=False).eval()(rays, want_weights=True)
    assert full["fine"]["weights"].shape == (1, B, 96)
    # weights update is picked up by the replicas
    with torch.no_grad():
        net.mlp_fine.lin_out.bias[:3] += 5.0
    rgb2, _ = multi(rays)
    assert (rgb2.mean() - rgbm.mean()).abs() > 0.05
