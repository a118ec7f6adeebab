"""
Host-side helpers the ray-rendering path and its callers need (the reference keeps them in
src/util/util.py).  Only what the path touches is provided: ray generation
(util.py:118-148,243-281), spherical poses (:314-328), repeat_interleave (:58-65),
combine_interleaved (:466-476) and psnr (:479-486).
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
