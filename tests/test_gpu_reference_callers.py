"""
The reference's OWN drivers, unmodified, on this repo's drop-in packages:

    eval/gen_video.py   (video of one object)            eval/eval.py   (PSNR/SSIM sweep over a split)

are taken byte-identical from the staged reference tree (git-ignored baseline/_ref, written by
oracle/stage_reference.py in the build container; their sha256 is checked against its manifest), placed in
a scratch tree whose ``src`` is a symlink to ``dropin/src`` -- exactly what a user does to switch -- and run
as subprocesses on a synthetic SRN-format dataset on disk (tests/helpers.write_srn_fixture, read by
pixel_nerf_multiscale_b200.data through the ``data`` shim) with a seeded checkpoint.  Third-party modules the
scripts import but the image lacks (imageio, ipdb, skimage) are stand-ins that record what was written /
restate the metric (test infrastructure; the imageio stand-in also seeds torch so that the renderer's
random draws can be replayed).  The frames they produce are compared with the fp32 oracle fed the same
draws.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import REPO, write_srn_fixture

pytestmark = pytest.mark.gpu
REF = os.path.join(REPO, "baseline", "_ref")
SEED = 4321

STUB_IMAGEIO = '''
import numpy as np, torch
torch.manual_seed(%d)          # imported before any random draw of the driver: makes the render replayable
def mimwrite(path, frames, **kw): np.save(path + ".npy", np.asarray(frames))
def imwrite(path, img, **kw): np.save(path + ".npy", np.asarray(img))
def imread(path): raise NotImplementedError
''' % SEED
STUB_IPDB = "def set_trace(*a, **k):\n    pass\n"
STUB_SKIMAGE = '''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from oracle import pixelnerf_oracle as po
def _both(a, b, data_range):
    return po.frame_metrics(np.asarray(a), np.asarray(b), data_range=float(data_range))
def compare_psnr(a, b, data_range=1): return _both(a, b, data_range)[0]
def compare_ssim(a, b, multichannel=True, data_range=1): return _both(a, b, data_range)[1]
''' % REPO


def _stage(tmp, script):
    src = os.path.join(REF, "eval", script)
    if not os.path.isfile(src):
        pytest.skip("reference tree is not staged under baseline/_ref (run oracle/stage_reference.py in the build container)")
    manifest = json.load(open(os.path.join(REF, "MANIFEST.json")))["files"]
    assert hashlib.sha256(open(src, "rb").read()).hexdigest() == manifest["eval/" + script], "staged driver was modified"
    os.makedirs(os.path.join(tmp, "eval"), exist_ok=True)
    shutil.copyfile(src, os.path.join(tmp, "eval", script))
    if not os.path.exists(os.path.join(tmp, "src")):
        os.symlink(os.path.join(REPO, "dropin", "src"), os.path.join(tmp, "src"))
    stubs = os.path.join(tmp, "stubs")
    os.makedirs(os.path.join(stubs, "skimage"), exist_ok=True)
    open(os.path.join(stubs, "imageio.py"), "w").write(STUB_IMAGEIO)
    open(os.path.join(stubs, "ipdb.py"), "w").write(STUB_IPDB)
    open(os.path.join(stubs, "skimage", "__init__.py"), "w").write("from . import measure\n")
    open(os.path.join(stubs, "skimage", "measure.py"), "w").write(STUB_SKIMAGE)
    return stubs


def _conf_and_checkpoint(tmp, name):
    import bench
    import pixel_nerf_multiscale_b200 as pk
    from pixel_nerf_multiscale_b200.util.conf import ConfigFactory

    conf_path = os.path.join(tmp, "test_srn.conf")
    open(conf_path, "w").write('include required("%s")\nmodel {\n  encoder {\n    pretrained = false\n  }\n}\n'
                               % os.path.join(REPO, "conf", "exp", "srn.conf"))
    conf = ConfigFactory.parse_file(conf_path)
    torch.manual_seed(0)
    net = pk.make_model(conf["model"]).eval()
    bench.rerandomise(net.mlp_coarse, 1)
    bench.rerandomise(net.mlp_fine, 2)
    os.makedirs(os.path.join(tmp, "checkpoints", name), exist_ok=True)
    torch.save(net.state_dict(), os.path.join(tmp, "checkpoints", name, "pixel_nerf_latest"))
    return conf_path, conf, net


def _run(tmp, stubs, script, argv, precision):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([stubs, REPO, env.get("PYTHONPATH", "")])
    env["PIXELNERF_B200_PRECISION"] = precision
    r = subprocess.run([sys.executable, os.path.join(tmp, "eval", script)] + argv, cwd=tmp, env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, "%s failed:\n%s\n%s" % (script, r.stdout[-2000:], r.stderr[-4000:])
    return r.stdout


def _oracle_frames(net, conf, item, src_views, rays, ray_batch, z_near, z_far):
    """The oracle's frames for `rays` (n,8), replaying the renderer's draws (seed SEED, per ray batch, in
    the reference's order) on the GPU."""
    import bench
    from oracle import pixelnerf_oracle as po
    from sigma_sweep import Replay

    dev = torch.device("cuda:0")
    net = net.to(dev)
    net.precision = "fp32"
    with torch.no_grad():
        net.encode(item["images"][src_views].unsqueeze(0).to(dev), item["poses"][src_views].unsqueeze(0).to(dev),
                   item["focal"][None].to(dev), c=item["c"].to(dev).unsqueeze(0))
    scene = bench.oracle_scene(net, None, conf, device=dev)
    kw = bench.renderer_kwargs(conf)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(SEED)
    out = []
    with torch.no_grad():
        for batch in torch.split(rays.to(dev), ray_batch, dim=0):
            n = batch.shape[0]
            tape = {"coarse": torch.rand(n, kw["n_coarse"], device=dev)}
            tape["u"] = torch.rand(n, kw["n_fine"] - kw["n_fine_depth"], device=dev)
            tape["jitter"] = torch.rand(n, kw["n_fine"] - kw["n_fine_depth"], device=dev)
            tape["normal"] = torch.randn(n, kw["n_fine_depth"], device=dev)
            res = po.render(scene, batch[None], tape=Replay(tape), eval_batch_size=200000, **kw)
            out.append(res["fine"]["rgb"])
    return torch.cat(out).cpu()


# (precision, tolerated |frame - oracle| at the 99.9th percentile, in levels of 255).  The encoder of this run is a
# random-init ResNet34 with identity BatchNorm statistics, whose features have a standard deviation of ~30 (a trained
# encoder emits O(1)): bf16 operands (8 significand bits) then show up to ~10 levels, f16 operands stay within
# the uint8 quantisation.
@pytest.mark.parametrize("precision,tol_levels", [("fp32", 1), ("fp16", 1), ("bf16", 10)])
def test_gen_video_runs_unmodified_on_the_dropin(tmp_path, precision, tol_levels):
    import pixel_nerf_multiscale_b200 as pk
    from pixel_nerf_multiscale_b200.data import get_split_dataset

    tmp = str(tmp_path)
    stubs = _stage(tmp, "gen_video.py")
    write_srn_fixture(os.path.join(tmp, "data"), name="cars", stage="test", n_obj=1, n_views=4, size=128, focal=131.25)
    conf_path, conf, net = _conf_and_checkpoint(tmp, "caller")
    out = _run(tmp, stubs, "gen_video.py",
               ["-n", "caller", "-c", conf_path, "-F", "srn", "-D", os.path.join(tmp, "data", "cars"), "--split", "test",
                "-P", "0 2", "--num_views", "2", "--checkpoints_path", os.path.join(tmp, "checkpoints"),
                "--visual_path", os.path.join(tmp, "visuals"), "--gpu_id", "0", "-R", "20000"], precision)
    assert "Rendering 32768 rays" in out
    vid = os.path.join(tmp, "visuals", "caller", "videot0000_v000_002.mp4.npy")
    frames = np.load(vid)
    assert frames.shape == (2, 128, 128, 3) and frames.dtype == np.uint8
    # the same frames from the oracle: the driver's default 360-degree loop (gen_video.py:157-172)
    dset = get_split_dataset("srn", os.path.join(tmp, "data", "cars"), want_split="test", training=False)
    item = dset[0]
    radius = (dset.z_near + dset.z_far) * 0.5
    poses = torch.stack([pk.util.pose_spherical(float(a), -10.0, radius) for a in np.linspace(-180, 180, 3)[:-1]])
    rays = pk.util.gen_rays(poses, 128, 128, item["focal"][None], dset.z_near, dset.z_far, c=item["c"][None]).reshape(-1, 8)
    ref = _oracle_frames(net, conf, item, torch.tensor([0, 2]), rays, 20000, dset.z_near, dset.z_far)
    ref_u8 = (ref.reshape(2, 128, 128, 3).numpy() * 255).astype(np.uint8)
    diff = np.abs(frames.astype(np.int32) - ref_u8.astype(np.int32))
    print("gen_video.py (%s): |frame - oracle| in levels of 255: max %d, p99.9 %.1f, mean %.4f, share > %d: %.5f" % (
        precision, diff.max(), np.percentile(diff, 99.9), diff.mean(), tol_levels, (diff > tol_levels).mean()))
    assert np.percentile(diff, 99.9) <= tol_levels and diff.mean() < (1.5 if precision == "bf16" else 0.5)
    assert frames.std() > 5.0   # a real image, not a constant


def test_eval_py_runs_unmodified_on_the_dropin(tmp_path):
    tmp = str(tmp_path)
    stubs = _stage(tmp, "eval.py")
    expect = write_srn_fixture(os.path.join(tmp, "data"), name="cars", stage="test", n_obj=2, n_views=4, size=128, focal=131.25)
    conf_path, conf, net = _conf_and_checkpoint(tmp, "caller")
    out_dir = os.path.join(tmp, "eval_out")
    out = _run(tmp, stubs, "eval.py",
               ["-n", "caller", "-c", conf_path, "-F", "srn", "-D", os.path.join(tmp, "data", "cars"), "--split", "test",
                "-P", "0 2", "-O", out_dir, "--checkpoints_path", os.path.join(tmp, "checkpoints"),
                "--visual_path", os.path.join(tmp, "visuals"), "--gpu_id", "0", "-R", "20000"], "fp16")
    lines = [l.split() for l in open(os.path.join(out_dir, "finish.txt")).read().splitlines() if l.strip()]
    assert [l[0] for l in lines] == ["obj_00", "obj_01"] and all(int(l[3]) == 1 for l in lines)
    assert "final psnr" in out
    # the metric in finish.txt is the PSNR of the written images against the dataset's ground truth
    from oracle import pixelnerf_oracle as po

    for k, l in enumerate(lines):
        ps = []
        for v in (1, 3):   # views 0 and 2 are the sources
            img = np.load(os.path.join(out_dir, "obj_%02d" % k, "%06d.png.npy" % v))
            assert img.shape == (128, 128, 3) and img.dtype == np.uint8
            gt = expect[k]["images"][v].astype(np.float32) / 255.0
            p, _ = po.frame_metrics(img.astype(np.float32) / 255.0, gt)
            ps.append(p)
        assert abs(float(l[1]) - np.mean(ps)) < 0.05, (l, ps)   # uint8 quantisation of the written image
