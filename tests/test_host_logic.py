"""Host-side mirror of the reference interface: conf schema, factories, renderer bookkeeping,
loud failure without CUDA.  CPU only."""
import math
import os
import types

import pytest
import torch

import pixel_nerf_multiscale_b200 as pk
from helpers import REPO
from pixel_nerf_multiscale_b200.parallel import shard_bounds
from pixel_nerf_multiscale_b200.render.nerf import RenderOutput
from pixel_nerf_multiscale_b200.util.conf import ConfigFactory


def _conf(name):
    c = ConfigFactory.parse_file(os.path.join(REPO, "conf", "exp", name))
    c["model"]["encoder"].put("pretrained", False)
    return c


def test_conf_inheritance_and_accessors():
    c = _conf("dtu.conf")
    assert c.get_int("model.mlp_coarse.n_blocks") == 5 and c["model"]["mlp_coarse"].get_int("combine_layer", 1000) == 3
    assert c.get_bool("renderer.white_bkgd") is False and c.get_float("model.code.freq_factor") == 1.5
    assert c.get_list("renderer.sched", None) == [] and c.get_string("data.format") == "dvr_dtu"
    assert "model.encoder.backbone" in c and "model.nope" not in c
    with pytest.raises(KeyError):
        c["model.nope"]
    s = ConfigFactory.parse_string('a { b = 1, c = [1, [2, 3]] }\na.d : "x y"\ne = 5e-4 // trailing\n')
    assert s["a.c"] == [1, [2, 3]] and s["a.d"] == "x y" and s.get_float("e") == 5e-4


@pytest.mark.parametrize("name,d_in,latent", [("sn64.conf", 42, 256), ("dtu.conf", 42, 256), ("sn64_multiscale.conf", 78, 512)])
def test_model_schema(name, d_in, latent):
    net = pk.make_model(_conf(name)["model"])
    assert net.d_in == d_in and net.latent_size == latent and net.d_out == 4
    assert net.use_viewdirs and net.use_xyz and net.normalize_z
    assert net.mlp_coarse.n_blocks == 5 and net.mlp_coarse.combine_layer == 3 and len(net.mlp_coarse.lin_z) == 3
    keys = net.state_dict().keys()
    for k in ("mlp_fine.blocks.4.fc_1.weight", "mlp_coarse.lin_z.2.bias", "encoder.model.conv1.weight",
              "encoder.layers.0.0.weight", "code._freqs", "code._phases"):
        assert k in keys
    # callers do `net.mlp_fine = None` (eval/eval.py:141)
    net.mlp_fine = None
    assert net.mlp_fine is None


def test_unsupported_configs_fail_loudly():
    c = _conf("dtu.conf")
    c["model"]["mlp_coarse"].put("beta", 100.0)
    with pytest.raises(NotImplementedError):
        pk.make_model(c["model"])
    c = _conf("dtu.conf")
    c["model"]["mlp_coarse"].put("combine_type", "max")
    with pytest.raises(NotImplementedError):
        pk.make_model(c["model"])
    c = _conf("dtu.conf")
    c["model"].put("use_global_encoder", True)
    with pytest.raises(NotImplementedError):
        pk.make_model(c["model"])
    # sampling options of SpatialEncoder.index the native gather does not implement (encoder.py:182-188)
    for key, val in (("index_interp", "nearest"), ("index_padding", "zeros")):
        c = _conf("dtu.conf")
        c["model"]["encoder"].put(key, val)
        net = pk.make_model(c["model"]).eval()
        with torch.no_grad():
            net.encode(torch.zeros(1, 1, 3, 32, 32), torch.eye(4)[None, None], torch.tensor(30.0))
        with pytest.raises(NotImplementedError):
            net.native_scene()


def test_encode_bookkeeping_matches_oracle():
    from oracle import pixelnerf_oracle as po

    net = pk.make_model(_conf("dtu.conf")["model"]).eval()
    poses = torch.stack([pk.util.pose_spherical(30.0 * i, -20.0, 2.2) for i in range(3)])[None]
    focal = torch.tensor([[72.3, 70.1]])
    c = torch.tensor([[20.0, 15.0]])
    with torch.no_grad():
        net.encode(torch.zeros(1, 3, 3, 32, 48), poses, focal, c=c)
    w2c, f, cc = po.encode_cameras(poses[0], focal, c, 48, 32)
    assert torch.allclose(net.poses, w2c) and torch.allclose(net.focal, f) and torch.allclose(net.c, cc)
    assert net.image_shape.tolist() == [48.0, 32.0] and net.num_views_per_obj == 3 and net.num_objs == 1
    assert net.encoder.latent.shape == (3, 256, 2, 3)
    with torch.no_grad():
        net.encode(torch.zeros(3, 3, 32, 48), poses[0], torch.tensor(50.0))
    assert net.num_views_per_obj == 1 and net.focal.tolist() == [[50.0, -50.0]] and net.c.tolist() == [[24.0, 16.0]]


def test_no_cpu_path():
    net = pk.make_model(_conf("sn64.conf")["model"]).eval()
    with torch.no_grad():
        net.encode(torch.zeros(1, 1, 3, 32, 32), torch.eye(4)[None, None], torch.tensor(40.0))
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 5, 3), coarse=True, viewdirs=torch.zeros(1, 5, 3))
    r = pk.NeRFRenderer(n_coarse=8, n_fine=4)
    with pytest.raises(RuntimeError):
        r(net, torch.zeros(1, 5, 8))
    with pytest.raises(RuntimeError):
        net.mlp_coarse(torch.zeros(4, 256 + 42))


def test_renderer_from_conf_sched_and_wrapper():
    conf = _conf("sn64.conf")
    r = pk.NeRFRenderer.from_conf(conf["renderer"], lindisp=False, eval_batch_size=50000)
    assert (r.n_coarse, r.n_fine, r.n_fine_depth, r.eval_batch_size) == (64, 32, 16, 50000)
    assert bool(r.white_bkgd) and r.using_fine and r.sched is None
    assert set(r.state_dict().keys()) == {"iter_idx", "last_sched"}
    r2 = pk.NeRFRenderer(n_coarse=16, n_fine=8, sched=[[2, 4], [32, 64], [16, 32]])
    r2.sched_step(1)
    assert r2.n_coarse == 16
    r2.sched_step(1)
    assert (r2.n_coarse, r2.n_fine, int(r2.last_sched)) == (32, 16, 1)
    r2.sched_step(5)
    assert (r2.n_coarse, r2.n_fine, int(r2.last_sched)) == (64, 32, 2)
    net = pk.make_model(conf["model"])
    par = r.bind_parallel(net, [0], simple_output=True).eval()
    rgb, depth = par(torch.zeros(0, 8))  # empty-batch guard needs no GPU (nerf.py:23-27)
    assert rgb.shape == (0, 3) and depth.shape == (0,)
    o = RenderOutput(coarse=RenderOutput(rgb=1))
    assert o.coarse.rgb == 1 and o.toDict() == {"coarse": {"rgb": 1}}


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 50000, 120000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_util_rays():
    poses = torch.stack([pk.util.pose_spherical(10.0, -10.0, 2.0)])
    rays = pk.util.gen_rays(poses, 8, 6, torch.tensor(20.0), 0.5, 3.0)
    assert rays.shape == (1, 6, 8, 8)
    assert torch.allclose(rays[..., 3:6].norm(dim=-1), torch.ones(1, 6, 8), atol=1e-6)
    assert torch.allclose(rays[..., :3], poses[0, :3, 3].expand(1, 6, 8, 3)) and rays[..., 6].eq(0.5).all()
    assert math.isclose(pk.util.psnr(torch.zeros(4), torch.full((4,), 0.1)), 20.0, rel_tol=1e-5)


def test_checkpoint_compatibility(tmp_path):
    """trainer dicts, DataParallel prefixes and upstream pixelNeRF files load strictly."""
    from pixel_nerf_multiscale_b200.model.checkpoint import normalize_state_dict

    conf = pk.util.conf.ConfigFactory.parse_file(os.path.join(REPO, "conf/exp/sn64.conf"))
    conf.put("model.encoder.pretrained", False)
    torch.manual_seed(0)
    src = pk.make_model(conf["model"])
    sd = src.state_dict()
    assert any(k.startswith("encoder.layers.") for k in sd)  # the fork's aliases of encoder.model.layerN
    upstream = {k: v for k, v in sd.items() if not k.startswith("encoder.layers.")}
    upstream["poses"] = torch.zeros(1, 3, 4)       # transient buffers some upstream versions saved
    upstream["encoder.latent"] = torch.zeros(1, 1, 1, 1)
    variants = {
        "own": sd,
        "upstream": upstream,
        "trainer": {"epoch": 3, "iter": 10, "net_state_dict": sd, "optimizer_state_dict": {}},
        "dataparallel": {"module." + k: v for k, v in upstream.items()},
    }
    for name, obj in variants.items():
        torch.manual_seed(1)
        dst = pk.make_model(conf["model"])
        res = dst.load_state_dict(obj, strict=True)
        assert not res.missing_keys and not res.unexpected_keys, name
        for k, v in dst.state_dict().items():
            assert torch.equal(v, sd[k]), (name, k)
    # load_weights reads the same formats from checkpoints/<name>/pixel_nerf_latest
    os.makedirs(tmp_path / "exp")
    torch.save(variants["trainer"], tmp_path / "exp" / "pixel_nerf_latest")
    args = types.SimpleNamespace(checkpoints_path=str(tmp_path), name="exp", resume=True)
    dst = pk.make_model(conf["model"]).load_weights(args)
    assert torch.equal(dst.mlp_coarse.lin_out.weight, src.mlp_coarse.lin_out.weight)
    # latent-width mismatch names the conf switch
    bad = dict(sd)
    bad["mlp_coarse.lin_z.0.weight"] = torch.zeros(512, 512)
    with pytest.raises(RuntimeError, match="use_multi_scale"):
        normalize_state_dict(bad, sd)


def test_output_tail_helpers():
    """Frame assembly / depth normalisation (eval.py:278-292) and the small driver-side helpers."""
    import numpy as np

    import pixel_nerf_multiscale_b200 as pk

    g = torch.Generator().manual_seed(0)
    rgb, depth = torch.rand(2 * 3 * 4, 3, generator=g) * 1.4 - 0.2, torch.rand(2 * 3 * 4, generator=g) * 3 + 1.2
    frames, dn = pk.util.assemble_frames(rgb, depth, 2, 3, 4, 1.2, 4.0)
    ref_rgb = np.clip(rgb.reshape(2, 3, 4, 3).numpy(), 0.0, 1.0)
    ref_d = ((depth - 1.2) / (4.0 - 1.2)).reshape(2, 3, 4).numpy()
    assert np.array_equal(frames.numpy(), ref_rgb) and np.allclose(dn.numpy(), ref_d, atol=1e-7)
    q = torch.tensor([[0.9698, 0.2121, 0.1203, -0.0039], [0.7020, 0.1578, 0.4525, 0.5268]])
    R = pk.util.quat_to_rot(q)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(2, 3, 3), atol=1e-6)
    assert torch.allclose(pk.util.quat_to_rot(pk.util.rot_to_quat(R)), R, atol=1e-5)
    assert torch.equal(pk.util.coord_from_blender() @ pk.util.coord_to_blender(), torch.eye(4))
    assert pk.util.get_cuda(0).type in ("cuda", "cpu")


def test_bench_step_definitions():
    """bench.py: every workload's steps cover whole frames (or whole 50 000-ray batches of the orbit), and both arms
    print the same config."""
    import bench

    for name, wl in bench.WORKLOADS.items():
        rng = bench.step_ranges(wl)
        per = wl["W"] * wl["H"]
        assert rng and all(hi > lo for lo, hi in rng) and rng[-1][1] <= wl["video_frames"] * per
        if wl["step"] == "batch":
            assert all(hi - lo == bench.RAY_BATCH for lo, hi in rng)
        else:
            assert all(hi - lo == per and lo % per == 0 for lo, hi in rng)
        cfg = bench.workload_config(name)
        assert cfg["workload"].startswith(name + ":") and "l2" in cfg and cfg == bench.workload_config(name)
    assert bench.workload_config("c3")["workload"].startswith("c3: DTU 300x400")
