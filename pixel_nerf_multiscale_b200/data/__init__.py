"""
Dataset adapters behind ``get_split_dataset`` -- the data side of the reference's drivers
(call sites eval/gen_video.py:68-90, eval/eval.py:99-101,199-240,
train/train.py).  The reference imports ``from data import get_split_dataset`` but its own ``src/data``
package is ABSENT from the tree, so these adapters restate the on-disk formats of upstream pixelNeRF
(SRN cars/chairs folders, DVR / NMR ShapeNet renders, the DVR-format DTU set) from the contract the
callers rely on (plus the two-object ``multi_obj`` scenes in NeRF-synthetic layout); they are pinned by round-trip tests on synthetic on-disk fixtures
(tests/test_data_adapters.py), not by reference goldens ("parity unpinned": nothing to compare with).

What every item is (the drivers index it as ``dset[i]`` or through a DataLoader with batch_size=1):
    path    object directory
    img_id  index of the object in the split
    images  (NV, 3, H, W) float32 in [-1, 1]
    poses   (NV, 4, 4) camera-to-world, camera looks down -z, +y up (the renderer's convention)
    focal   0-dim tensor (fx == fy) or (2,) tensor (fx, fy)  [pixels]
    c       (2,) principal point [pixels]  (absent for plain DVR ShapeNet renders: image centre)
    masks, bbox  (NV,1,H,W), (NV,4) where the format provides / implies them
and every dataset carries ``z_near``, ``z_far``, ``lindisp`` (and ``sub_format`` for DVR).
"""
import glob
import os

import numpy as np
import torch

__all__ = ["get_split_dataset", "SRNDataset", "DVRDataset", "MultiObjectDataset"]

_FLIP_YZ = torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))  # OpenCV camera (y down, z forward) <-> (y up, z back)


def _read_rgb(path, image_size=None):
    """uint8 (H,W,3) RGB; `image_size` = (H, W) resizes with area filtering."""
    import cv2

    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 2:
        img = np.repeat(img[..., None], 3, axis=2)
    if img.shape[2] == 4:  # white background under the alpha channel (SRN renders)
        a = img[..., 3:4].astype(np.float32) / 255.0
        img = (img[..., :3].astype(np.float32) * a + 255.0 * (1.0 - a)).astype(np.uint8)
    img = img[..., ::-1]  # BGR -> RGB
    if image_size is not None and tuple(img.shape[:2]) != tuple(image_size):
        img = cv2.resize(np.ascontiguousarray(img), (image_size[1], image_size[0]), interpolation=cv2.INTER_AREA)
    return np.ascontiguousarray(img)


def _to_tensor_balanced(img_u8):
    """uint8 (H,W,3) -> float (3,H,W) in [-1,1]."""
    return torch.from_numpy(img_u8).permute(2, 0, 1).float().div_(255.0).mul_(2.0).sub_(1.0)


def _bbox_of(mask):
    """(H,W) bool -> [xmin, ymin, xmax, ymax] of the foreground (whole image if empty)."""
    rows, cols = np.any(mask, axis=1), np.any(mask, axis=0)
    if not rows.any():
        return torch.tensor([0, 0, mask.shape[1] - 1, mask.shape[0] - 1], dtype=torch.float32)
    r, c = np.where(rows)[0], np.where(cols)[0]
    return torch.tensor([c[0], r[0], c[-1], r[-1]], dtype=torch.float32)


class SRNDataset(torch.utils.data.Dataset):
    """SRN ShapeNet cars / chairs: ``<path>_<stage>/<obj>/{intrinsics.txt, rgb/*.png, pose/*.txt}``.
    intrinsics.txt: first line ``focal cx cy _`` (pixels at the stored resolution), last line ``H W``;
    pose/*.txt: 16 floats, camera-to-world in OpenCV convention."""

    def __init__(self, path, stage="train", image_size=(128, 128), world_scale=1.0):
        super().__init__()
        self.base_path = path + "_" + stage
        self.dataset_name = os.path.basename(path)
        assert os.path.exists(self.base_path), "SRN split directory %s does not exist" % self.base_path
        is_chair = "chair" in self.dataset_name
        if is_chair and stage == "train":
            # upstream ships the chairs train split one level deeper
            tmp = os.path.join(self.base_path, "chairs_2.0_train")
            if os.path.exists(tmp):
                self.base_path = tmp
        self.intrins = sorted(glob.glob(os.path.join(self.base_path, "*", "intrinsics.txt")))
        self.stage = stage
        self.image_size = tuple(image_size)
        self.world_scale = world_scale
        self.z_near, self.z_far = (1.25, 2.75) if is_chair else (0.8, 1.8)
        self.lindisp = False

    def __len__(self):
        return len(self.intrins)

    def __getitem__(self, index):
        intrin_path = self.intrins[index]
        dir_path = os.path.dirname(intrin_path)
        rgb_paths = sorted(glob.glob(os.path.join(dir_path, "rgb", "*")))
        pose_paths = sorted(glob.glob(os.path.join(dir_path, "pose", "*")))
        assert len(rgb_paths) == len(pose_paths), "%s: %d images, %d poses" % (dir_path, len(rgb_paths), len(pose_paths))
        with open(intrin_path, "r") as f:
            lines = f.readlines()
        focal, cx, cy, _ = map(float, lines[0].split())
        height, width = map(int, lines[-1].split())
        images, poses, masks, bboxes = [], [], [], []
        for rgb_path, pose_path in zip(rgb_paths, pose_paths):
            img = _read_rgb(rgb_path)
            fg = (img != 255).all(axis=-1)
            pose = torch.from_numpy(np.loadtxt(pose_path, dtype=np.float32).reshape(4, 4))
            images.append(_to_tensor_balanced(img))
            masks.append(torch.from_numpy(fg.astype(np.float32))[None])
            bboxes.append(_bbox_of(fg))
            poses.append(pose @ _FLIP_YZ)
        images, poses = torch.stack(images), torch.stack(poses)
        masks, bboxes = torch.stack(masks), torch.stack(bboxes)
        if tuple(images.shape[-2:]) != self.image_size:
            scale = self.image_size[0] / images.shape[-2]
            focal, cx, cy = focal * scale, cx * scale, cy * scale
            bboxes = bboxes * scale
            images = torch.nn.functional.interpolate(images, size=self.image_size, mode="area")
            masks = torch.nn.functional.interpolate(masks, size=self.image_size, mode="area")
        if self.world_scale != 1.0:
            focal *= self.world_scale
            poses[:, :3, 3] *= self.world_scale
        return {"path": dir_path, "img_id": index, "focal": torch.tensor(focal, dtype=torch.float32),
                "c": torch.tensor([cx, cy], dtype=torch.float32), "images": images, "masks": masks,
                "bbox": bboxes, "poses": poses}


class DVRDataset(torch.utils.data.Dataset):
    """DVR-format renders: ``<path>/<category>/<list_prefix><stage>.lst`` lists objects, each with
    ``image/*`` (or ``rgb/*``), optional ``mask/*`` and ``cameras.npz`` holding per view
    ``world_mat_<i>`` (world -> image projection, 4x4), ``camera_mat_<i>`` (NDC intrinsics, ShapeNet) and
    optionally ``world_mat_inv_<i>`` / ``scale_mat_<i>``.
    sub_format "shapenet": NMR ShapeNet 64x64 renders (normalised intrinsics, focal scaled by W/2);
    sub_format "dtu": DTU scans; world_mat is K[R|t] in pixels and is decomposed; scale_mat normalises
    the scene into the unit sphere."""

    def __init__(self, path, stage="train", list_prefix="softras_", image_size=None, sub_format="shapenet",
                 scale_focal=True, max_imgs=100000, z_near=1.2, z_far=4.0, skip_step=None):
        super().__init__()
        self.base_path = path
        assert os.path.exists(self.base_path), "DVR dataset directory %s does not exist" % path
        cats = sorted(x for x in glob.glob(os.path.join(path, "*")) if os.path.isdir(x))
        list_name = {"train": "train", "val": "val", "test": "test"}[stage]
        self.all_objs = []
        for cat in cats:
            lst = os.path.join(cat, list_prefix + list_name + ".lst")
            if not os.path.exists(lst):
                continue
            with open(lst, "r") as f:
                self.all_objs += [(os.path.basename(cat), os.path.join(cat, x.strip())) for x in f.readlines() if x.strip()]
        self.stage = stage
        self.image_size = None if image_size is None else tuple(image_size)
        self.sub_format = sub_format
        self.scale_focal = scale_focal
        self.max_imgs = max_imgs
        self.z_near, self.z_far = z_near, z_far
        self.lindisp = False
        if sub_format == "dtu":
            self._world = _FLIP_YZ.clone()
            self._cam = _FLIP_YZ.clone()
        else:  # ShapeNet renders: z-up world -> y-up world, OpenCV camera -> renderer camera
            self._world = torch.tensor([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
            self._cam = _FLIP_YZ.clone()
        self.skip_step = skip_step

    def __len__(self):
        return len(self.all_objs)

    def __getitem__(self, index):
        import cv2

        cat, root_dir = self.all_objs[index]
        rgb_paths = sorted(x for x in glob.glob(os.path.join(root_dir, "image", "*")) if x.lower().endswith((".jpg", ".png")))
        if not rgb_paths:
            rgb_paths = sorted(x for x in glob.glob(os.path.join(root_dir, "rgb", "*")) if x.lower().endswith((".jpg", ".png")))
        mask_paths = sorted(glob.glob(os.path.join(root_dir, "mask", "*.png")))
        if len(mask_paths) == 0:
            mask_paths = [None] * len(rgb_paths)
        sel = np.arange(len(rgb_paths))
        if len(rgb_paths) > self.max_imgs:
            sel = np.sort(np.random.choice(len(rgb_paths), self.max_imgs, replace=False))
        elif self.skip_step:
            sel = sel[:: self.skip_step]
        cams = np.load(os.path.join(root_dir, "cameras.npz"))
        images, poses, masks, bboxes = [], [], [], []
        fx = fy = cx = cy = 0.0
        for i in sel:
            img = _read_rgb(rgb_paths[i])
            h, w = img.shape[:2]
            x_scale, y_scale = (w / 2.0, h / 2.0) if self.scale_focal else (1.0, 1.0)
            if self.sub_format == "dtu":
                P = cams["world_mat_%d" % i][:3]
                K, R, t = cv2.decomposeProjectionMatrix(P)[:3]
                K = K / K[2, 2]
                pose = np.eye(4, dtype=np.float32)
                pose[:3, :3] = R.transpose()
                pose[:3, 3] = (t[:3] / t[3])[:, 0]
                key = "scale_mat_%d" % i
                if key in cams.files:
                    sm = cams[key]
                    pose[:3, 3:] = (pose[:3, 3:] - sm[:3, 3:]) / np.diagonal(sm[:3, :3])[:, None]
                fx += float(K[0, 0]) * x_scale
                fy += float(K[1, 1]) * y_scale
                cx += (float(K[0, 2]) + 1.0) * x_scale if self.scale_focal else float(K[0, 2])
                cy += (float(K[1, 2]) + 1.0) * y_scale if self.scale_focal else float(K[1, 2])
            else:
                key = "world_mat_inv_%d" % i
                pose = cams[key] if key in cams.files else np.linalg.inv(cams["world_mat_%d" % i])
                intr = cams["camera_mat_%d" % i]
                assert abs(float(intr[0, 0]) - float(intr[1, 1])) < 1e-6, "DVR ShapeNet renders have fx == fy"
                fx += float(intr[0, 0]) * x_scale
                fy += float(intr[1, 1]) * y_scale
                cx += (float(intr[0, 2]) + 1.0) * x_scale
                cy += (float(intr[1, 2]) + 1.0) * y_scale
            pose = self._world @ torch.tensor(np.asarray(pose, dtype=np.float32)) @ self._cam
            images.append(_to_tensor_balanced(img))
            poses.append(pose)
            if mask_paths[i] is not None:
                m = _read_rgb(mask_paths[i])[..., 0] > 127
                masks.append(torch.from_numpy(m.astype(np.float32))[None])
                bboxes.append(_bbox_of(m))
        n = float(len(sel))
        fx, fy, cx, cy = fx / n, fy / n, cx / n, cy / n
        images, poses = torch.stack(images), torch.stack(poses)
        item = {"path": root_dir, "img_id": index}
        if self.image_size is not None and tuple(images.shape[-2:]) != self.image_size:
            sy, sx = self.image_size[0] / images.shape[-2], self.image_size[1] / images.shape[-1]
            fx, cx, fy, cy = fx * sx, cx * sx, fy * sy, cy * sy
            images = torch.nn.functional.interpolate(images, size=self.image_size, mode="area")
            if masks:
                masks = [torch.nn.functional.interpolate(m[None], size=self.image_size, mode="area")[0] for m in masks]
                bboxes = [b * torch.tensor([sx, sy, sx, sy]) for b in bboxes]
        item["focal"] = torch.tensor(fx, dtype=torch.float32) if abs(fx - fy) < 1e-6 * max(1.0, abs(fx)) else \
            torch.tensor([fx, fy], dtype=torch.float32)
        if self.sub_format == "dtu" or abs(cx - images.shape[-1] / 2.0) > 1e-4 or abs(cy - images.shape[-2] / 2.0) > 1e-4:
            item["c"] = torch.tensor([cx, cy], dtype=torch.float32)
        item["images"], item["poses"] = images, poses
        if masks:
            item["masks"], item["bbox"] = torch.stack(masks), torch.stack(bboxes)
        return item


class MultiObjectDataset(torch.utils.data.Dataset):
    """Two-object ShapeNet scenes in NeRF-synthetic layout: ``<path>/<stage>/<scene>/transforms.json`` with
    ``camera_angle_x`` and ``frames[{file_path, transform_matrix}]`` (camera-to-world, already in the renderer's
    convention), RGBA PNGs next to it (composited on white; the alpha channel is the mask)."""

    def __init__(self, path, stage="train", z_near=4.0, z_far=9.0, n_views=None):
        super().__init__()
        base = os.path.join(path, stage)
        assert os.path.exists(base), "multi_obj split directory %s does not exist" % base
        self.scenes = sorted(os.path.dirname(x) for x in glob.glob(os.path.join(base, "*", "transforms.json")))
        self.stage, self.n_views = stage, n_views
        self.z_near, self.z_far, self.lindisp = z_near, z_far, False

    def __len__(self):
        return len(self.scenes)

    def __getitem__(self, index):
        import cv2
        import json

        root = self.scenes[index]
        with open(os.path.join(root, "transforms.json"), "r") as f:
            meta = json.load(f)
        frames = meta["frames"]
        if self.n_views is not None and self.n_views < len(frames):
            frames = [frames[i] for i in np.sort(np.random.choice(len(frames), self.n_views, replace=False))]
        images, masks, bboxes, poses = [], [], [], []
        for fr in frames:
            fpath = os.path.join(root, os.path.basename(fr["file_path"]) + ".png")
            raw = cv2.imread(fpath, cv2.IMREAD_UNCHANGED)
            if raw is None:
                raise FileNotFoundError(fpath)
            fg = raw[..., 3] > 127 if raw.shape[2] == 4 else np.ones(raw.shape[:2], dtype=bool)
            images.append(_to_tensor_balanced(_read_rgb(fpath)))
            masks.append(torch.from_numpy(fg.astype(np.float32))[None])
            bboxes.append(_bbox_of(fg))
            poses.append(torch.tensor(fr["transform_matrix"], dtype=torch.float32))
        images = torch.stack(images)
        focal = 0.5 * images.shape[-1] / np.tan(0.5 * float(meta["camera_angle_x"]))
        return {"path": root, "img_id": index, "focal": torch.tensor(focal, dtype=torch.float32), "images": images,
                "masks": torch.stack(masks), "bbox": torch.stack(bboxes), "poses": torch.stack(poses)}


def get_split_dataset(dataset_type, datadir, want_split="all", training=True, **kwargs):
    """
    Dataset(s) of the requested split.
    :param dataset_type  srn | multi_obj | dvr | dvr_gen | dvr_dtu
    :param want_split    train | val | test | all (-> (train, val, test))
    :param training      only affects augmentation / image caps of upstream's training mode
    """
    flags = {}
    if dataset_type == "srn":
        cls = SRNDataset
    elif dataset_type.startswith("dvr"):
        cls = DVRDataset
        if dataset_type == "dvr_gen":       # category-generalisation split lists
            flags["list_prefix"] = "gen_"
        elif dataset_type == "dvr_dtu":     # DTU in DVR's format
            flags.update(list_prefix="new_", sub_format="dtu", scale_focal=False, z_near=0.1, z_far=5.0)
            if training:
                flags["max_imgs"] = 49
    elif dataset_type == "multi_obj":
        cls = MultiObjectDataset
    else:
        raise NotImplementedError("Unsupported dataset type %r" % (dataset_type,))
    flags.update(kwargs)
    want = ("train", "val", "test") if want_split == "all" else (want_split,)
    sets = tuple(cls(datadir, stage=s, **flags) for s in want)
    return sets if want_split == "all" else sets[0]
