"""
Pins the oracle (oracle/pixelnerf_oracle.py) against outputs of the unmodified reference
(tests/golden/*.pt, produced by tests/golden/make_golden.py).  CPU only.
Tolerances: fp32 restatement vs fp32 reference on the same CPU build -> 2e-5 abs on O(1)
quantities (summation order inside addmm / cumprod may differ), indices and sample
positions essentially exact.
"""
import math

import pytest
import torch

from oracle import pixelnerf_oracle as po
from oracle import synth
from helpers import (N_POINTS, RENDER_SEED, load_conf, load_golden, maxabs, renderer_kwargs, sample_points)

CASES = ["ss_ns1", "ms_ns2", "dtu_ns3", "ms_ns3_sb2", "sv3_ns1", "dtu_ns3_s6", "ms_ns2_s6"]


@pytest.mark.parametrize("name", CASES)
def test_mlp_input_and_net_forward(name):
    gold = load_golden(name)
    conf = load_conf(name)
    scene, raw = synth.build_case(name, conf["model"])
    case = synth.CASES[name]
    xyz, vd = sample_points(case, case["sb"], N_POINTS, 7)
    zx, _ = po.mlp_input(scene, xyz, vd)
    assert zx.shape == gold["zx"].shape
    # gathered latent + code + viewdirs: pure elementwise/bilinear arithmetic
    assert maxabs(zx, gold["zx"]) < 2e-5
    for coarse, key in ((True, "net_coarse"), (False, "net_fine")):
        out = po.net_forward(scene, xyz, coarse=coarse, viewdirs=vd)
        assert out.shape == gold[key].shape
        assert maxabs(out[..., :3], gold[key][..., :3]) < 2e-5
        # sigma is unbounded (O(10)): relative tolerance
        assert torch.allclose(out[..., 3], gold[key][..., 3], rtol=2e-5, atol=2e-5)


def _render_variants(name):
    gold = load_golden(name)
    return [k[len("render_"):] for k in gold if k.startswith("render_")]


@pytest.mark.parametrize("name,variant", [(n, v) for n in CASES for v in _render_variants(n)])
def test_render(name, variant):
    gold = load_golden(name)["render_" + variant]
    conf = load_conf(name)
    scene, raw = synth.build_case(name, conf["model"])
    case = synth.CASES[name]
    rays = synth.target_rays(case, case["rays"], 3, case["sb"])
    kw = renderer_kwargs(conf, gold["kw"])
    torch.manual_seed(RENDER_SEED)
    res = po.render(scene, rays, eval_batch_size=1500, **kw)
    # RNG tape equals the reference's own draws -> coarse sample positions are bit-equal
    assert torch.equal(res["coarse"]["z"], gold["z_coarse"])
    sb = case["sb"]
    for lvl in ("coarse", "fine"):
        if lvl + "_rgb" not in gold:
            assert lvl not in res
            continue
        assert maxabs(res[lvl]["rgb"].reshape(sb, -1, 3), gold[lvl + "_rgb"]) < 2e-5
        assert maxabs(res[lvl]["depth"].reshape(sb, -1), gold[lvl + "_depth"]) < 5e-5
        assert maxabs(res[lvl]["weights"].reshape(sb, -1, res[lvl]["weights"].shape[-1]), gold[lvl + "_weights"]) < 2e-5
    if "z_fine" in gold:
        # importance samples depend on the coarse weights through a bin lookup: identical
        # except where a 1e-7 cdf difference flips a bin (none at these sizes)
        assert maxabs(res["fine"]["z"], gold["z_fine"]) < 1e-5


def test_fine_indices_definition():
    """searchsorted(right=True)-1 with clamp_min only (nerf.py:138-139): ties and overflow."""
    cdf = torch.tensor([[0.0, 0.25, 0.25, 0.75, 0.999]])
    u = torch.tensor([[0.0, 0.2499, 0.25, 0.5, 0.75, 0.9995]])
    assert po.fine_indices(cdf, u).tolist() == [[0.0, 0.0, 2.0, 2.0, 3.0, 4.0]]


def test_positional_encoding_layout():
    x = torch.tensor([[0.1, -0.2, 0.3]])
    e = po.positional_encoding(x, 2, 1.5, True)
    assert e.shape == (1, 15)
    assert torch.allclose(e[0, :3], x[0])
    assert torch.allclose(e[0, 3:6], torch.sin(1.5 * x[0]), atol=1e-6)
    assert torch.allclose(e[0, 6:9], torch.cos(1.5 * x[0]), atol=1e-6)
    assert torch.allclose(e[0, 9:12], torch.sin(3.0 * x[0]), atol=1e-6)


def test_frame_metrics_restatement():
    """SSIM/PSNR restatement (skimage is absent): brute-force window loop on a small image + invariants."""
    import numpy as np

    rng = np.random.default_rng(0)
    H, W, C, win = 12, 11, 3, 7
    gt = rng.random((H, W, C)).astype(np.float32)
    img = (gt + 0.1 * rng.standard_normal((H, W, C))).astype(np.float32)  # leaves [0,1]: exercises the clamp
    psnr, ssim = po.frame_metrics(img, gt)
    a = np.clip(img, 0, 1).astype(np.float64)
    b = gt.astype(np.float64)
    acc = []
    for c in range(C):
        for y in range(H - win + 1):
            for x in range(W - win + 1):
                p, q = a[y:y + win, x:x + win, c].ravel(), b[y:y + win, x:x + win, c].ravel()
                ux, uy = p.mean(), q.mean()
                vx, vy = p.var(ddof=1), q.var(ddof=1)
                vxy = ((p - ux) * (q - uy)).sum() / (win * win - 1)
                acc.append(((2 * ux * uy + 1e-4) * (2 * vxy + 9e-4)) / ((ux * ux + uy * uy + 1e-4) * (vx + vy + 9e-4)))
    assert abs(ssim - float(np.mean(acc))) < 1e-12
    assert abs(psnr - 10 * math.log10(1.0 / np.mean((a - b) ** 2))) < 1e-12
    assert abs(po.frame_metrics(gt, gt + 0.0)[1] - 1.0) < 1e-12       # identical images
    const = np.full((9, 9, 1), 0.5, np.float32)
    assert abs(po.frame_metrics(const, const - 0.1)[0] - 20.0) < 1e-5  # mse = 0.01
