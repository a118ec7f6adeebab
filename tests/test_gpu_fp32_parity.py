"""
GPU parity, fp32 validation mode: the CUDA path (through the C ABI) against the oracle on the
same device and against the reference-generated goldens.  Tolerance (BASELINE.json north_star):
per-pixel RGB/depth max-abs <= 1e-4 in fp32 mode.
"""
import math

import pytest
import torch

from oracle import pixelnerf_oracle as po
from oracle import synth
from helpers import (N_POINTS, RENDER_SEED, build_product, load_golden, make_renderer, maxabs, renderer_kwargs,
                     sample_points)

pytestmark = pytest.mark.gpu
CASES = ["ss_ns1", "ms_ns2", "dtu_ns3", "ms_ns3_sb2", "sv3_ns1", "dtu_ns3_s6", "ms_ns2_s6"]
TOL = 1e-4


@pytest.mark.parametrize("name", CASES)
def test_point_features_and_net_forward(name):
    from pixel_nerf_multiscale_b200 import _native as N

    gold = load_golden(name)
    net, conf, scene, raw = build_product(name, precision="fp32")
    case = synth.CASES[name]
    xyz, vd = sample_points(case, case["sb"], N_POINTS, 7)
    xyz, vd = xyz.cuda(), vd.cuda()
    # kernel (a) alone: rows of the MLP input
    sc, keep = net.native_scene()
    rows = case["sb"] * case["ns"] * N_POINTS
    zx = torch.empty(rows, raw["d_latent"] + raw["d_in"], device="cuda")
    N.check(N.lib().pnr_point_features_f32(sc, N.ptr(xyz.contiguous()), N.ptr(vd.contiguous()), case["sb"], N_POINTS,
                                           N.ptr(zx), N.stream_ptr(zx.device)), "features")
    assert maxabs(zx.cpu(), gold["zx"]) < 2e-5
    # kernels (a)+(b): PixelNeRFNet.forward, coarse and fine MLP
    for coarse, key in ((True, "net_coarse"), (False, "net_fine")):
        out = net(xyz, coarse=coarse, viewdirs=vd)
        ref = po.net_forward(scene, xyz, coarse=coarse, viewdirs=vd)
        assert out.shape == gold[key].shape
        for other in (ref.cpu(), gold[key]):
            assert maxabs(out[..., :3].cpu(), other[..., :3]) < TOL
            assert torch.allclose(out[..., 3].cpu(), other[..., 3], rtol=1e-4, atol=1e-4)


def test_mlp_rows_api():
    """ResnetFC.forward on caller-provided rows (pnr_mlp_forward) vs oracle."""
    net, conf, scene, raw = build_product("ms_ns2", precision="fp32")
    torch.manual_seed(5)
    ns, p = 2, 37
    zx = torch.randn(ns * p, raw["d_latent"] + raw["d_in"], device="cuda")
    from pixel_nerf_multiscale_b200 import _native as N

    out = net.mlp_coarse(zx, combine_inner_dims=(ns, p), precision=N.FP32)
    ref = po.resnetfc_forward(scene.mlp_coarse, zx, raw["d_latent"], 5, 3, (ns, p))
    assert out.reshape(-1, 4).shape == ref.reshape(-1, 4).shape
    assert torch.allclose(out.reshape(-1, 4), ref.reshape(-1, 4), rtol=1e-4, atol=1e-4)


def _variants(name):
    return [k[len("render_"):] for k in load_golden(name) if k.startswith("render_")]


@pytest.mark.parametrize("name,variant", [(n, v) for n in CASES for v in _variants(n)])
def test_render_vs_golden_and_oracle(name, variant):
    gold = load_golden(name)["render_" + variant]
    net, conf, scene, raw = build_product(name, precision="fp32")
    case = synth.CASES[name]
    sb = case["sb"]
    rays = synth.target_rays(case, case["rays"], 3, sb).cuda()
    kw = renderer_kwargs(conf, gold["kw"])
    renderer = make_renderer(conf, gold["kw"])
    # replay the reference's own random draws: CPU generator, same seed, same order
    torch.manual_seed(RENDER_SEED)
    tape = po.RngTape(rays.shape[0] * rays.shape[1], kw["n_coarse"], kw["n_fine"], kw["n_fine_depth"], "cpu")
    tape.draw_coarse()
    if kw["n_fine"] > 0:
        tape.draw_fine()
    cu = lambda t: None if t is None else t.cuda()
    dtape = {k: v for k, v in dict(coarse=cu(tape.coarse), u=cu(tape.u), jitter=cu(tape.jit), normal=cu(tape.nrm)).items()
             if v is not None}
    renderer.rng_tape = dict(dtape)
    res = renderer(net, rays, want_weights=True, taps=True)
    assert torch.equal(res.coarse.z.reshape(-1, kw["n_coarse"]).cpu(), gold["z_coarse"])
    for lvl in ("coarse", "fine"):
        if lvl + "_rgb" not in gold:
            assert lvl not in res
            continue
        assert maxabs(res[lvl].rgb.cpu(), gold[lvl + "_rgb"]) < TOL
        assert maxabs(res[lvl].depth.cpu(), gold[lvl + "_depth"]) < TOL
        assert maxabs(res[lvl].weights.cpu(), gold[lvl + "_weights"]) < TOL
    if "z_fine" in gold:
        assert maxabs(res.fine.z.reshape(gold["z_fine"].shape).cpu(), gold["z_fine"]) < 1e-5
    # same thing through the generic (model-callable) path: per-ray kernels + model() calls
    renderer.rng_tape = dict(dtape)
    class Opaque:  # not a PixelNeRFNet instance -> generic path
        use_viewdirs = net.use_viewdirs

        def __call__(self, x, coarse=True, viewdirs=None):
            return net(x, coarse=coarse, viewdirs=viewdirs)

    renderer.eval_batch_size = 1500
    res2 = renderer(Opaque(), rays, want_weights=True)
    last = "fine" if "fine" in res else "coarse"
    assert maxabs(res2[last].rgb, res[last].rgb) < 1e-5
    assert maxabs(res2[last].depth, res[last].depth) < 1e-5


def test_bind_parallel_wrapper_outputs():
    net, conf, scene, raw = build_product("ss_ns1", precision="fp32")
    case = synth.CASES["ss_ns1"]
    rays = synth.target_rays(case, 50, 3, 1).cuda()
    renderer = make_renderer(conf, {})
    par = renderer.bind_parallel(net, [0], simple_output=True).eval()
    torch.manual_seed(1)
    rgb, depth = par(rays)
    assert rgb.shape == (1, 50, 3) and depth.shape == (1, 50)
    full = renderer.bind_parallel(net, None, simple_output=False).eval()
    torch.manual_seed(1)
    d = full(rays, want_weights=True)
    assert set(d.keys()) == {"coarse", "fine"} and set(d["fine"].keys()) == {"rgb", "depth", "weights"}
    assert torch.equal(d["fine"]["rgb"], rgb)
    # empty batch guard (nerf.py:23-27)
    e_rgb, e_depth = par(rays[:0])
    assert e_rgb.shape == (0, 3) and e_depth.shape == (0,)


def test_fine_indices_bit_exact_given_cdf():
    """Importance-sample indices must be bit-exact given identical CDFs (north_star)."""
    from pixel_nerf_multiscale_b200 import _native as N

    torch.manual_seed(3)
    B, Kc, Kf = 4096, 64, 16
    w = torch.rand(B, Kc) ** 4
    w[::7] = 0.0  # empty rays -> uniform cdf
    cdf = po.fine_cdf(w)
    u = torch.rand(B, Kf)
    u[0, :4] = torch.tensor([0.0, 1.0 - 1e-7, cdf[0, 5].item(), cdf[0, 64].item()])  # ties / top edge
    ref = po.fine_indices(cdf, u)
    inds = torch.empty(B, Kf, device="cuda")
    cd, ud = cdf.cuda().contiguous(), u.cuda().contiguous()
    N.check(N.lib().pnr_fine_indices(N.ptr(cd), N.ptr(ud), B, Kc, Kf, N.ptr(inds), N.stream_ptr(inds.device)), "inds")
    assert torch.equal(inds.cpu(), ref)


def _sample_fine_sorted(rays, z_c, w, depth, u, jit, nrm, Kc, n_fine, n_dep, depth_std=0.01):
    from pixel_nerf_multiscale_b200 import _native as N

    B = rays.shape[0]
    out = torch.empty(B, Kc + n_fine, device="cuda")
    c = lambda t: None if t is None else t.cuda().contiguous()
    keep = [c(t) for t in (rays, z_c, w, depth, u, jit, nrm)]
    N.check(N.lib().pnr_sample_fine_sorted(*[N.ptr(t) for t in keep], B, Kc, n_fine, n_dep, depth_std, 0, N.ptr(out),
                                           N.stream_ptr(out.device)), "sample_fine_sorted")
    return out.cpu()


@pytest.mark.parametrize("Kc,n_dep,unsorted", [(64, 32, False), (64, 16, False), (64, 48, False), (120, 8, False),
                                               (64, 32, True), (33, 1, False)])
def test_sample_fine_sort_and_merge_bit_exact(Kc, n_dep, unsorted):
    """cat + sort of nerf.py:286-295 with depth-guided samples only (their arithmetic is elementwise, so the
    sorted result must be BIT-equal to torch.sort): the merge-by-rank fast path (<= 32 new samples, with and
    without padding lanes), the general bitonic path (> 32) and the fallback for a coarse list that is not
    ascending, incl. ties between coarse and new samples."""
    torch.manual_seed(11 + Kc + n_dep)
    B = 2048
    near, far = 0.5, 4.5
    rays = torch.zeros(B, 8)
    rays[:, 6], rays[:, 7] = near, far
    z_c = po.sample_coarse(rays, Kc, False, torch.rand(B, Kc))
    if unsorted:
        z_c[::3, [5, 6]] = z_c[::3, [6, 5]]          # an inversion in every third ray
    depth = near + (far - near) * torch.rand(B)
    nrm = torch.randn(B, n_dep)
    depth[::5] = z_c[::5, 7]                          # exact ties between a coarse and a new sample
    nrm[::5, 0] = 0.0
    got = _sample_fine_sorted(rays, z_c, None, depth, None, None, nrm, Kc, n_dep, n_dep)
    ref, _ = torch.sort(torch.cat([z_c, po.sample_fine_depth(rays, depth, 0.01, nrm)], dim=-1), dim=-1)
    assert torch.equal(got, ref)


def test_sample_fine_importance_vs_oracle():
    """Full sample_fine + sample_fine_depth + sort against the oracle on the CPU (sequential cumsum, like the
    reference on its CPU path).  The sum of the pdf is accumulated in another order than torch.sum, so a cdf
    entry may differ in its last bit and move a sample sitting exactly on a bin edge: all but a vanishing share
    of rays must be bit-equal."""
    torch.manual_seed(5)
    B, Kc, n_fine, n_dep = 4096, 64, 32, 16
    rays = torch.zeros(B, 8)
    rays[:, 6], rays[:, 7] = 0.8, 1.8
    z_c = po.sample_coarse(rays, Kc, False, torch.rand(B, Kc))
    w = torch.rand(B, Kc) ** 6
    w[::9] = 0.0
    depth = 0.8 + torch.rand(B)
    u, jit, nrm = torch.rand(B, n_fine - n_dep), torch.rand(B, n_fine - n_dep), torch.randn(B, n_dep)
    got = _sample_fine_sorted(rays, z_c, w, depth, u, jit, nrm, Kc, n_fine, n_dep)
    ref, _ = torch.sort(torch.cat([z_c, po.sample_fine(rays, w, Kc, False, u, jit), po.sample_fine_depth(rays, depth, 0.01, nrm)], -1), -1)
    same = (got == ref).all(dim=-1)
    print("rays bit-equal to the oracle: %d / %d" % (int(same.sum()), B))
    assert same.float().mean().item() > 0.995
    assert (got[:, 1:] >= got[:, :-1]).all()


def test_errors_are_loud():
    net, conf, scene, raw = build_product("ss_ns1", precision="fp32")
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 4, 3), coarse=True, viewdirs=torch.zeros(1, 4, 3))  # CPU tensors: no fallback
    with pytest.raises(AssertionError):
        net.mlp_coarse(torch.zeros(8, 17, device="cuda"))  # wrong row width (resnetfc.py:190)


def test_gen_rays_on_device():
    """device-side ray generation vs the oracle's gen_rays (itself bit-equal to the
    reference's util.gen_rays on CPU)."""
    import pixel_nerf_multiscale_b200 as pk

    poses = torch.stack([po.pose_spherical(33.0, -12.0, 2.1), po.pose_spherical(-63.0, -10.0, 1.3),
                         po.pose_spherical(170.0, 25.0, 3.0)])
    for focal, c, W, H in ((torch.tensor([[72.3, 70.0]]), torch.tensor([[20.0, 15.0]]), 40, 30), (torch.tensor(119.4), None, 64, 64)):
        ref = po.gen_rays(poses, W, H, focal, 0.1, 5.0, c)
        got = pk.util.gen_rays(poses.cuda(), W, H, focal, 0.1, 5.0, c)
        assert got.shape == ref.shape and got.is_cuda
        assert maxabs(got.cpu(), ref) < 1e-6
        assert torch.equal(got[..., :3].cpu(), ref[..., :3]) and torch.equal(got[..., 6:].cpu(), ref[..., 6:])


def test_finalize_frames_on_device():
    """clamp / uint8 / PSNR on the device vs the numpy/torch formulation of the drivers."""
    import numpy as np
    import pixel_nerf_multiscale_b200 as pk

    g = torch.Generator().manual_seed(9)
    rgb = torch.rand(3, 37, 41, 3, generator=g) * 1.4 - 0.2
    gt = torch.rand(3, 37, 41, 3, generator=g)
    u8, psnr = pk.util.finalize_frames(rgb.cuda(), gt.cuda())
    ref_u8 = (rgb.clamp(0, 1).numpy() * 255).astype(np.uint8)
    assert np.array_equal(u8.cpu().numpy(), ref_u8)
    ref_psnr = pk.util.psnr(rgb.clamp(0, 1).double(), gt.double())
    assert abs(psnr.item() - ref_psnr) < 1e-6
    u8b, none = pk.util.finalize_frames(rgb.cuda())
    assert none is None and torch.equal(u8b, u8)


def test_frame_metrics_on_device():
    """per-view PSNR / SSIM of the eval driver (eval/eval.py:314-343) on the device vs the
    fp64 restatement of skimage's compare_psnr / compare_ssim; fp64 window sums -> 1e-9."""
    import pixel_nerf_multiscale_b200 as pk

    g = torch.Generator().manual_seed(11)
    NV, H, W = 3, 37, 52
    gt = torch.rand(NV, H, W, 3, generator=g)
    rgb = gt + 0.08 * torch.randn(NV, H, W, 3, generator=g)  # partly outside [0,1]
    rgb[2] = gt[2]                                           # identical view: ssim 1, psnr inf
    psnr, ssim = pk.util.frame_metrics(rgb.cuda(), gt.cuda())
    assert psnr.shape == (NV,) and ssim.shape == (NV,) and psnr.dtype == torch.float64
    for v in range(2):
        rp, rs = po.frame_metrics(rgb[v].numpy(), gt[v].numpy())
        assert abs(psnr[v].item() - rp) < 1e-9 and abs(ssim[v].item() - rs) < 1e-9, (v, psnr[v].item(), rp, ssim[v].item(), rs)
    assert math.isinf(psnr[2].item()) and abs(ssim[2].item() - 1.0) < 1e-12
    # single image form + loud argument errors
    p1, s1 = pk.util.frame_metrics(rgb[0].cuda(), gt[0].cuda())
    assert torch.equal(p1, psnr[:1]) and abs(s1.item() - ssim[0].item()) < 1e-12  # atomics order may differ in the last bit
    with pytest.raises(AssertionError):
        pk.util.frame_metrics(torch.rand(1, 5, 5, 3).cuda(), torch.rand(1, 5, 5, 3).cuda())  # 7x7 window does not fit


def test_render_views_generates_rays_on_device():
    """the frame loop with per-rank on-device ray generation equals rendering the same
    rays passed in explicitly (same torch RNG stream, one batch per frame)."""
    import pixel_nerf_multiscale_b200 as pk
    from pixel_nerf_multiscale_b200.parallel import render_views

    net, conf, scene, raw = build_product("ss_ns1", precision="fp32")
    renderer = make_renderer(conf, {})
    par = renderer.bind_parallel(net, [0], simple_output=True).eval()
    W, H = 12, 10
    poses = torch.stack([pk.util.pose_spherical(25.0 * i, -15.0, 2.6) for i in range(2)]).cuda()
    torch.manual_seed(5)
    rgb, depth = render_views(par, poses, W, H, 60.0, 1.2, 4.0, ray_batch_size=W * H)
    assert rgb.shape == (2, H, W, 3) and depth.shape == (2, H, W)
    rays = pk.util.gen_rays(poses, W, H, 60.0, 1.2, 4.0).reshape(2, W * H, 8)
    torch.manual_seed(5)
    for f in range(2):
        r, d = par(rays[f][None])
        assert torch.equal(r[0], rgb[f].reshape(-1, 3)) and torch.equal(d[0], depth[f].reshape(-1))
    assert torch.isfinite(rgb).all() and float(depth.min()) >= 0.0
