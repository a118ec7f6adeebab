"""Dataset adapters behind get_split_dataset: round trips through synthetic on-disk fixtures in
the SRN, DVR-ShapeNet and DVR-DTU formats.  (The reference's own src/data is absent: there is nothing to
compare with except the contract its drivers rely on.)"""
import os

import numpy as np
import pytest
import torch

from helpers import write_srn_fixture


def test_srn_round_trip(tmp_path):
    from pixel_nerf_multiscale_b200.data import get_split_dataset

    expect = write_srn_fixture(str(tmp_path), name="cars", stage="test", n_obj=2, n_views=3, size=16, focal=21.5)
    dset = get_split_dataset("srn", os.path.join(str(tmp_path), "cars"), want_split="test", training=False, image_size=(16, 16))
    assert len(dset) == 2 and (dset.z_near, dset.z_far, dset.lindisp) == (0.8, 1.8, False)
    for item, e in zip((dset[0], dset[1]), expect):
        assert item["path"] == e["path"]
        assert item["images"].shape == (3, 3, 16, 16) and item["images"].dtype == torch.float32
        u8 = ((item["images"].permute(0, 2, 3, 1) * 0.5 + 0.5) * 255).round().numpy().astype(np.uint8)
        assert np.array_equal(u8, e["images"])
        assert np.allclose(item["poses"].numpy(), e["poses"], atol=1e-6)
        assert abs(float(item["focal"]) - 21.5) < 1e-6 and item["focal"].dim() == 0
        assert torch.allclose(item["c"], torch.tensor([8.0, 8.0]))
        assert item["masks"].shape == (3, 1, 16, 16) and item["bbox"].shape == (3, 4)
    # resizing rescales the intrinsics
    from pixel_nerf_multiscale_b200.data import SRNDataset

    half = SRNDataset(os.path.join(str(tmp_path), "cars"), stage="test", image_size=(8, 8))[0]
    assert half["images"].shape[-2:] == (8, 8) and abs(float(half["focal"]) - 10.75) < 1e-6
    with pytest.raises(AssertionError):
        get_split_dataset("srn", os.path.join(str(tmp_path), "cars"), want_split="all")  # no train split on disk


def _write_images(d, n, size, g):
    import cv2

    os.makedirs(os.path.join(d, "image"), exist_ok=True)
    for v in range(n):
        cv2.imwrite(os.path.join(d, "image", "%04d.png" % v), g.randint(0, 255, size=(size[0], size[1], 3)).astype(np.uint8))


def test_dvr_shapenet_and_dtu_round_trip(tmp_path):
    import pixel_nerf_multiscale_b200 as pk
    from pixel_nerf_multiscale_b200.data import get_split_dataset

    g = np.random.RandomState(1)
    flip = np.diag([1.0, -1.0, -1.0, 1.0])
    zup = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)
    # ---- ShapeNet renders (NMR): normalised intrinsics, world_mat_inv given
    root = os.path.join(str(tmp_path), "nmr")
    obj = os.path.join(root, "02958343", "abc")
    _write_images(obj, 3, (64, 64), g)
    with open(os.path.join(root, "02958343", "softras_test.lst"), "w") as f:
        f.write("abc\n")
    want = [pk.util.pose_spherical(30.0 * v, -20.0, 2.7).double().numpy() for v in range(3)]
    cams = {}
    for v, pose in enumerate(want):
        stored = np.linalg.inv(zup) @ pose @ np.linalg.inv(flip)   # loader applies world @ stored @ cam
        cams["world_mat_inv_%d" % v] = stored
        cams["world_mat_%d" % v] = np.linalg.inv(stored)
        cams["camera_mat_%d" % v] = np.array([[3.7, 0, 0, 0], [0, 3.7, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
    np.savez(os.path.join(obj, "cameras.npz"), **cams)
    dset = get_split_dataset("dvr", root, want_split="test", training=False)
    item = dset[0]
    assert len(dset) == 1 and item["images"].shape == (3, 3, 64, 64) and dset.sub_format == "shapenet"
    assert np.allclose(item["poses"].numpy(), np.stack(want), atol=1e-5)
    assert abs(float(item["focal"]) - 3.7 * 32.0) < 1e-4 and "c" not in item
    assert (dset.z_near, dset.z_far) == (1.2, 4.0)
    # ---- DTU in DVR's format: world_mat = K [R|t] in pixels, decomposed by the loader; scale_mat normalises
    root = os.path.join(str(tmp_path), "dtu")
    obj = os.path.join(root, "scans", "scan21")
    _write_images(obj, 2, (30, 40), g)
    with open(os.path.join(root, "scans", "new_val.lst"), "w") as f:
        f.write("scan21\n")
    K = np.array([[72.3, 0, 20.0], [0, 72.3, 15.0], [0, 0, 1.0]])
    cams, want = {}, []
    scale = np.diag([2.0, 2.0, 2.0, 1.0])
    scale[:3, 3] = [0.3, -0.2, 0.1]
    for v in range(2):
        pose = pk.util.pose_spherical(25.0 * v + 10.0, -12.0, 2.2).double().numpy()   # renderer convention, normalised scene
        want.append(pose)
        cv_pose = flip @ pose @ flip                       # loader: FLIP @ stored @ FLIP
        cv_pose[:3, 3] = cv_pose[:3, 3] * 2.0 + scale[:3, 3]   # un-normalise (scale_mat)
        w2c = np.linalg.inv(cv_pose)
        P = np.eye(4)
        P[:3] = K @ w2c[:3]
        cams["world_mat_%d" % v] = P
        cams["scale_mat_%d" % v] = scale
    np.savez(os.path.join(obj, "cameras.npz"), **cams)
    dset = get_split_dataset("dvr_dtu", root, want_split="val", training=False)
    item = dset[0]
    assert dset.sub_format == "dtu" and (dset.z_near, dset.z_far) == (0.1, 5.0)
    assert item["images"].shape == (2, 3, 30, 40)
    assert np.allclose(item["poses"].numpy(), np.stack(want), atol=1e-4)
    assert abs(float(item["focal"]) - 72.3) < 1e-3 and torch.allclose(item["c"], torch.tensor([20.0, 15.0]), atol=1e-3)


def test_multi_obj_round_trip(tmp_path):
    import json

    import cv2

    import pixel_nerf_multiscale_b200 as pk
    from pixel_nerf_multiscale_b200.data import get_split_dataset

    g = np.random.RandomState(2)
    root = os.path.join(str(tmp_path), "two_obj", "val", "scene_000")
    os.makedirs(root)
    want, imgs, frames = [], [], []
    for v in range(3):
        pose = pk.util.pose_spherical(50.0 * v, -25.0, 6.0)
        rgba = g.randint(0, 255, size=(24, 24, 4)).astype(np.uint8)
        rgba[..., 3] = np.where(g.rand(24, 24) > 0.5, 255, 0)
        cv2.imwrite(os.path.join(root, "r_%d.png" % v), rgba[..., [2, 1, 0, 3]])
        frames.append({"file_path": "./r_%d" % v, "transform_matrix": pose.tolist()})
        want.append(pose.numpy())
        a = rgba[..., 3:4].astype(np.float32) / 255.0
        imgs.append((rgba[..., :3].astype(np.float32) * a + 255.0 * (1 - a)).astype(np.uint8))
    json.dump({"camera_angle_x": 0.7, "frames": frames}, open(os.path.join(root, "transforms.json"), "w"))
    dset = get_split_dataset("multi_obj", os.path.join(str(tmp_path), "two_obj"), want_split="val", training=False)
    item = dset[0]
    assert len(dset) == 1 and (dset.z_near, dset.z_far, dset.lindisp) == (4.0, 9.0, False)
    assert np.allclose(item["poses"].numpy(), np.stack(want), atol=1e-6)
    assert abs(float(item["focal"]) - 0.5 * 24 / np.tan(0.35)) < 1e-4
    u8 = ((item["images"].permute(0, 2, 3, 1) * 0.5 + 0.5) * 255).round().numpy().astype(np.uint8)
    assert np.array_equal(u8, np.stack(imgs))
    assert item["masks"].shape == (3, 1, 24, 24) and item["bbox"].shape == (3, 4)
