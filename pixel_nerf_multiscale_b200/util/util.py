"""
Host-side helpers the ray-rendering path and its callers need (the reference keeps them in
src/util/util.py).  Only what the path and its drivers (eval/gen_video.py, eval/eval.py,
eval/eval_real.py, eval/eval_approx.py) touch is provided: ray generation (util.py:118-148,243-281),
spherical poses (:314-328), repeat_interleave (:58-65), combine_interleaved (:466-476), psnr (:479-486),
get_cuda (:198-207), quat_to_rot / rot_to_quat (:489-533), the Blender coordinate flips (:151-176),
cmap (:13-30), batched_index_select_nd (:33-55), get_image_to_tensor_balanced / get_mask_to_tensor
(:68-86), get_module (:536-543) -- plus the device-side output tail (finalize_frames, frame_metrics,
normalize_depth).
"""
import math

import torch


def repeat_interleave(input, repeats, dim=0):
    """(N, ...) -> (N*repeats, ...) with each row repeated consecutively."""
    if dim != 0:
        raise NotImplementedError("repeat_interleave: only dim=0 (as in the reference)")
    return input.unsqueeze(1).expand(-1, repeats, *input.shape[1:]).reshape(-1, *input.shape[1:])


def combine_interleaved(t, inner_dims=(1,), agg_type="average"):
    """View pooling of (SB*NS*B, ...) rows; no-op for inner_dims == (1,)."""
    if len(inner_dims) == 1 and inner_dims[0] == 1:
        return t
    t = t.reshape(-1, *inner_dims, *t.shape[1:])
    if agg_type == "average":
        return t.mean(dim=1)
    if agg_type == "max":
        return t.max(dim=1)[0]
    raise NotImplementedError("Unsupported combine type " + agg_type)


def unproj_map(width, height, f, c=None, device="cpu"):
    """Unit camera-space ray direction per pixel, (H, W, 3); camera looks down -z, +y up."""
    if c is None:
        cx, cy = width * 0.5, height * 0.5
    else:
        cc = torch.as_tensor(c, dtype=torch.float32).flatten()
        cx, cy = float(cc[0]), float(cc[-1]) if cc.numel() > 1 else float(cc[0])
    ft = torch.as_tensor(f, dtype=torch.float32).flatten()
    fx, fy = (float(ft[0]), float(ft[0])) if ft.numel() == 1 else (float(ft[0]), float(ft[1]))
    ys = torch.arange(height, dtype=torch.float32) - cy
    xs = torch.arange(width, dtype=torch.float32) - cx
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    X = X.to(device=device) / fx
    Y = Y.to(device=device) / fy
    d = torch.stack((X, -Y, -torch.ones_like(X)), dim=-1)
    return d / torch.norm(d, dim=-1).unsqueeze(-1)


def gen_rays(poses, width, height, focal, z_near, z_far, c=None, ndc=False):
    """(N,4,4) camera-to-world poses -> (N,H,W,8) rays [origin(3) dir(3) near far]."""
    if ndc:
        raise NotImplementedError("ndc rays (undefined in the reference as well: util.py:265)")
    n = poses.shape[0]
    dev = poses.device
    if poses.is_cuda:  # generated on the device by the native library (no (N,H,W,8) host tensor, no H2D)
        from .. import _native as N

        ft = torch.as_tensor(focal, dtype=torch.float32).flatten()
        fx, fy = (float(ft[0]), float(ft[0])) if ft.numel() == 1 else (float(ft[0]), float(ft[1]))
        if c is None:
            cx, cy = width * 0.5, height * 0.5
        else:
            cc = torch.as_tensor(c, dtype=torch.float32).flatten()
            cx, cy = float(cc[0]), float(cc[-1])
        p = poses.detach().float().contiguous()
        rays = torch.empty(n, height, width, 8, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib().pnr_gen_rays(N.ptr(p), n, width, height, fx, fy, cx, cy, float(z_near), float(z_far),
                                         N.ptr(rays), N.stream_ptr(dev)), "pnr_gen_rays")
        return rays
    cam = unproj_map(width, height, torch.as_tensor(focal).squeeze(), c=c, device=dev)
    dirs = torch.matmul(poses[:, None, None, :3, :3], cam[None, ..., None])[..., 0]
    cen = poses[:, None, None, :3, 3].expand(-1, height, width, -1)
    near = torch.full((n, height, width, 1), float(z_near), device=dev)
    far = torch.full((n, height, width, 1), float(z_far), device=dev)
    return torch.cat((cen, dirs, near, far), dim=-1)


def pose_spherical(theta, phi, radius):
    """Camera-to-world pose on a sphere (degrees), NeRF convention."""
    t, p = theta / 180.0 * math.pi, phi / 180.0 * math.pi
    trans = torch.eye(4)
    trans[2, 3] = radius
    rp = torch.tensor([[1, 0, 0, 0], [0, math.cos(p), -math.sin(p), 0], [0, math.sin(p), math.cos(p), 0],
                       [0, 0, 0, 1]], dtype=torch.float32)
    rt = torch.tensor([[math.cos(t), 0, -math.sin(t), 0], [0, 1, 0, 0], [math.sin(t), 0, math.cos(t), 0],
                       [0, 0, 0, 1]], dtype=torch.float32)
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    return flip @ (rt @ (rp @ trans))


def psnr(pred, target):
    mse = ((pred - target) ** 2).mean()
    return -10 * math.log10(mse)


def get_cuda(gpu_id):
    """torch.device of GPU `gpu_id`, or the CPU device when CUDA is unavailable (the rendering path itself
    then raises: there is no CPU implementation)."""
    return torch.device("cuda:%d" % gpu_id) if torch.cuda.is_available() else torch.device("cpu")


def get_module(net):
    """net.module for DataParallel-style wrappers (incl. parallel.MultiDeviceRenderer), else net."""
    return net.module if hasattr(net, "module") and isinstance(net.module, torch.nn.Module) else net


def quat_to_rot(q):
    """(B,4) quaternions (w,x,y,z; normalised here) -> (B,3,3) rotation matrices."""
    q = torch.nn.functional.normalize(q, dim=1)
    w, x, y, z = q.unbind(dim=1)
    rows = (1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y))
    return torch.stack(rows, dim=1).reshape(-1, 3, 3)


def rot_to_quat(R):
    """(B,3,3) rotations -> (B,4) quaternions (w,x,y,z), the trace > -1 branch the reference uses."""
    w = torch.sqrt(1.0 + R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]) / 2
    return torch.stack((w, (R[:, 2, 1] - R[:, 1, 2]) / (4 * w), (R[:, 0, 2] - R[:, 2, 0]) / (4 * w),
                        (R[:, 1, 0] - R[:, 0, 1]) / (4 * w)), dim=1)


def coord_from_blender(dtype=torch.float32, device="cpu"):
    """Blender (x right, y in, z up) -> standard (x right, y up, z out) transform, (4,4)."""
    return torch.tensor([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], dtype=dtype, device=device)


def coord_to_blender(dtype=torch.float32, device="cpu"):
    """Standard -> Blender coordinate transform, (4,4)."""
    return torch.tensor([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=dtype, device=device)


def batched_index_select_nd(t, inds):
    """t (B, N, ...), inds (B, M) -> (B, M, ...): per-batch row selection along dim 1."""
    idx = inds.reshape(inds.shape[0], inds.shape[1], *([1] * (t.dim() - 2))).expand(-1, -1, *t.shape[2:])
    return t.gather(1, idx)


def get_image_to_tensor_balanced(image_size=0):
    """PIL image -> (3,H,W) tensor in [0,1] (the fork dropped upstream's [-1,1] normalisation)."""
    from torchvision import transforms

    ops = [transforms.Resize(image_size)] if image_size > 0 else []
    return transforms.Compose(ops + [transforms.ToTensor()])


def get_mask_to_tensor():
    from torchvision import transforms

    return transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.0,), (1.0,))])


def image_float_to_uint8(img):
    """float image -> uint8 after min/max stretching (numpy)."""
    import numpy as np

    lo, hi = float(np.min(img)), float(np.max(img))
    if hi - lo < 1e-10:
        hi += 1e-10
    return ((img - lo) / (hi - lo) * 255.0).astype(np.uint8)


def cmap(img, color_map=None):
    """'HOT' colour map of a float image (eval.py --write_depth)."""
    import cv2

    return cv2.applyColorMap(image_float_to_uint8(img), cv2.COLORMAP_HOT if color_map is None else color_map)


def normalize_depth(depth, z_near, z_far):
    """(depth - z_near) / (z_far - z_near), the depth image eval/eval.py:284 writes -- on the depth's device."""
    return (depth - float(z_near)) / (float(z_far) - float(z_near))


def frame_metrics(rgb, target, data_range=1.0, win_size=7):
    """Per-view PSNR and SSIM of rendered frames against ground truth, on the device and without a
    host sync -- what eval/eval.py:314-343 computes per view with skimage's compare_psnr /
    compare_ssim(multichannel=True, data_range=1) after clamping the render to [0,1].
    :param rgb, target (NV,H,W,C) or (H,W,C) CUDA tensors
    :return (psnr (NV,), ssim (NV,)) fp64 device tensors"""
    from .. import _native as N

    if not rgb.is_cuda:
        raise RuntimeError("pixelnerf_b200: frame_metrics needs CUDA tensors")
    x = rgb.detach().float().contiguous()
    gt = target.detach().float().contiguous().to(x.device)
    assert x.shape == gt.shape and x.dim() in (3, 4)
    if x.dim() == 3:
        x, gt = x[None], gt[None]
    NV, H, W, Cc = x.shape
    sums = torch.zeros(NV, 2, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().pnr_frame_metrics(N.ptr(x), N.ptr(gt), NV, H, W, Cc, int(win_size), float(data_range),
                                          N.ptr(sums), N.stream_ptr(x.device)), "pnr_frame_metrics")
    ssim = sums[:, 0] / float(Cc * (H - win_size + 1) * (W - win_size + 1))
    mse = sums[:, 1] / float(H * W * Cc)
    return 10.0 * torch.log10(float(data_range) ** 2 / mse), ssim


def finalize_frames(rgb, target=None):
    """Device-side tail of the eval drivers: clamp to [0,1], quantise to uint8 (what gen_video writes)
    and, with a ground-truth image, PSNR of the clamped frame -- all without a host sync.
    :return (uint8 tensor like rgb, psnr 0-dim fp64 device tensor | None)"""
    from .. import _native as N

    if not rgb.is_cuda:
        raise RuntimeError("pixelnerf_b200: finalize_frames needs CUDA tensors")
    x = rgb.detach().float().contiguous()
    u8 = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    gt = sse = None
    if target is not None:
        gt = target.detach().float().contiguous().to(x.device)
        assert gt.shape == x.shape
        sse = torch.zeros((), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().pnr_finalize_rgb(N.ptr(x), N.ptr(gt), x.numel(), N.ptr(u8), N.ptr(sse), N.stream_ptr(x.device)),
                "pnr_finalize_rgb")
    return u8, (None if sse is None else -10.0 * torch.log10(sse / x.numel()))


def assemble_frames(rgb, depth, n_views, height, width, z_near, z_far):
    """Frame assembly of the eval drivers on the tensors' own device (eval/eval.py:278-292,
    eval/gen_video.py:219-222): flat per-ray outputs -> rgb (NV,H,W,3) clamped to [0,1] and the normalised
    depth image (NV,H,W) = (depth - z_near) / (z_far - z_near) -- no per-batch .cpu() round trip."""
    rgb = torch.clamp(rgb.reshape(n_views, height, width, 3), 0.0, 1.0)
    return rgb, normalize_depth(depth.reshape(n_views, height, width), z_near, z_far)
