"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

A plain-torch fp32 restatement (explicit arithmetic, no nn.Module, no grid_sample) of the
reference's ray-rendering hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file, and only as the checker or the
reported CPU baseline -- never as the product path.

Parity status: PINNED BY EXECUTION, not by reference tests.  The reference holds no golden
vectors or tests for this path.  This file is pinned against outputs of
the UNMODIFIED reference run in the build container (oracle/ref_loader.py loads
/root/reference/src with model.models := models.py.backup2); the generating script is
tests/golden/make_golden.py and the vectors live in tests/golden/*.pt.
tests/test_oracle_golden.py checks every stage of this file against those vectors.

Every function cites the reference lines it restates (paths relative to /root/reference).
All tensors fp32; works on any torch device.
"""
import math

import torch

# --------------------------------------------------------------------------------------
# camera block                                                   models.py.backup2:98-153
# --------------------------------------------------------------------------------------


def encode_cameras(poses_c2w, focal, c, width, height):
    """
    Restates PixelNeRFNet.encode's camera bookkeeping (src/model/models.py.backup2:120-150).

    poses_c2w (V,4,4) camera->world.  Returns world->camera (V,3,4) = [R^T | -R^T t],
    focal (F,2) = (fx, -fy)  (fy negated, :139), c (C,2) (defaults to the image centre,
    :141-143).  F and C are 1 or V exactly as the reference leaves them.
    """
    rot = poses_c2w[:, :3, :3].transpose(1, 2)
    trans = -torch.bmm(rot, poses_c2w[:, :3, 3:])
    w2c = torch.cat((rot, trans), dim=-1)
    focal = torch.as_tensor(focal, dtype=torch.float32, device=poses_c2w.device)
    if focal.dim() == 0:
        focal = focal[None, None].repeat(1, 2)
    elif focal.dim() == 1:
        focal = focal.unsqueeze(-1).repeat(1, 2)
    else:
        focal = focal.clone()
    focal = focal.float()
    focal[..., 1] *= -1.0
    if c is None:
        c = torch.tensor([[width * 0.5, height * 0.5]], dtype=torch.float32, device=poses_c2w.device)
    else:
        c = torch.as_tensor(c, dtype=torch.float32, device=poses_c2w.device)
        if c.dim() == 0:
            c = c[None, None].repeat(1, 2)
        elif c.dim() == 1:
            c = c.unsqueeze(-1).repeat(1, 2)
    return w2c, focal, c


def _expand_views(t, sb, ns):
    """focal / c rows are per object; repeat-interleave x NS only when there is more than
    one row (models.py.backup2:216-221)."""
    if t.shape[0] == 1:
        return t.expand(sb * ns, -1)
    if t.shape[0] == sb * ns:
        return t
    return t.unsqueeze(1).expand(-1, ns, -1).reshape(sb * ns, -1)


# --------------------------------------------------------------------------------------
# bilinear feature gather                                             encoder.py:138-205
# --------------------------------------------------------------------------------------


def index_level(fmap, uv):
    """
    One pyramid level of SpatialEncoder.index (src/model/encoder.py:174-188) with
    bilinear / border / align_corners=True, written out explicitly.

    The reference maps pixel uv to the grid with uv/(W_i-1)*2-1 where W_i is the FEATURE MAP
    width (image_size is ignored); with align_corners=True grid_sample maps
    that back to texel x = ((g+1)/2)*(W_i-1).  The round trip is restated literally (not
    simplified to x=u) so that fp32 rounding matches.  Border padding clamps the texel
    coordinate to [0, W_i-1]; the +1 tap of a coordinate sitting exactly on the last texel
    falls outside and contributes zero (its weight is zero too).

    fmap (V,C,H,W), uv (V,N,2) pixel coords -> (V,C,N)
    """
    V, C, H, W = fmap.shape
    gx = (uv[..., 0] / (W - 1)) * 2 - 1
    gy = (uv[..., 1] / (H - 1)) * 2 - 1
    ix = ((gx + 1) / 2) * (W - 1)
    iy = ((gy + 1) / 2) * (H - 1)
    ix = torch.clamp(ix, 0, W - 1)
    iy = torch.clamp(iy, 0, H - 1)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = fmap.reshape(V, C, H * W)

    def tap(xx, yy, ww):
        ok = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
        xi = xx.clamp(0, W - 1).long()
        yi = yy.clamp(0, H - 1).long()
        idx = (yi * W + xi).unsqueeze(1).expand(-1, C, -1)
        val = torch.gather(flat, 2, idx)
        return val * (ww * ok.to(ww.dtype)).unsqueeze(1)

    return tap(x0, y0, w_nw) + tap(x1, y0, w_ne) + tap(x0, y1, w_sw) + tap(x1, y1, w_se)


def index_features(latents, uv):
    """SpatialEncoder.index (encoder.py:138-205): per-level gather, channel concat (:193).
    latents: list of (V,C_i,H_i,W_i) (one entry for the single-scale encoder)."""
    if uv.shape[0] == 1 and latents[0].shape[0] > 1:  # encoder.py:148-149
        uv = uv.expand(latents[0].shape[0], -1, -1)
    return torch.cat([index_level(f, uv) for f in latents], dim=1)


# --------------------------------------------------------------------------------------
# positional encoding                                                     code.py:30-47
# --------------------------------------------------------------------------------------


def positional_encoding(x, num_freqs=6, freq_factor=math.pi, include_input=True):
    """
    PositionalEncoding.forward (src/model/code.py:41-46): for each frequency f_k =
    freq_factor*2^k emit sin(f_k x + 0), sin(f_k x + pi/2) -- each d_in wide, in that
    order -- after the raw input.  The phase add is an fp32 addcmul, so cos is computed as
    sin(x*f + fl32(pi/2)), not as cos.
    """
    n, d = x.shape
    freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)  # code.py:15 (fp32 tensor)
    fr = torch.repeat_interleave(freqs, 2).view(1, -1, 1).to(x.device)
    ph = torch.zeros(2 * num_freqs)
    ph[1::2] = math.pi * 0.5
    ph = ph.view(1, -1, 1).to(x.device)
    emb = x.unsqueeze(1).repeat(1, num_freqs * 2, 1)
    emb = torch.sin(torch.addcmul(ph, emb, fr)).reshape(n, -1)
    if include_input:
        emb = torch.cat((x, emb), dim=-1)
    return emb


# --------------------------------------------------------------------------------------
# ResnetFC                                                           resnetfc.py:173-236
# --------------------------------------------------------------------------------------


def _linear(x, sd, name):
    return torch.addmm(sd[name + ".bias"], x, sd[name + ".weight"].t())


def resnetfc_forward(sd, zx, d_latent, n_blocks, combine_layer, combine_inner_dims, combine_type="average"):
    """
    ResnetFC.forward (src/model/resnetfc.py:193-235) with ReLU activations (beta=0) and no
    SPADE: x=lin_in(code); per block: view-pool at combine_layer (util.py:466-476), then
    x += lin_z[b](z) while b < combine_layer, then x + fc_1(relu(fc_0(relu(x))))
    (resnetfc.py:53-62); out = lin_out(relu(x)).
    sd: state-dict slice of one MLP (keys 'lin_in.weight', 'blocks.0.fc_0.weight', ...).
    """
    z = zx[..., :d_latent]
    x = _linear(zx[..., d_latent:], sd, "lin_in")
    for b in range(n_blocks):
        if b == combine_layer:
            if not (len(combine_inner_dims) == 1 and combine_inner_dims[0] == 1):
                x = x.reshape(-1, *combine_inner_dims, x.shape[-1])
                if combine_type == "average":
                    x = x.mean(dim=1)
                elif combine_type == "max":
                    x = x.max(dim=1)[0]
                else:
                    raise NotImplementedError(combine_type)
                x = x.reshape(-1, x.shape[-1]) if x.dim() > 2 else x
        if d_latent > 0 and b < combine_layer:
            x = x + _linear(z, sd, "lin_z.%d" % b)
        net = _linear(torch.relu(x), sd, "blocks.%d.fc_0" % b)
        dx = _linear(torch.relu(net), sd, "blocks.%d.fc_1" % b)
        x = x + dx
    return _linear(torch.relu(x), sd, "lin_out")


# --------------------------------------------------------------------------------------
# PixelNeRFNet.forward                                       models.py.backup2:155-282
# --------------------------------------------------------------------------------------


class Scene:
    """Everything PixelNeRFNet.encode leaves behind + the static model hyper-parameters."""

    def __init__(self, latents, w2c, focal, c, ns, mlp_coarse, mlp_fine, *, d_latent, n_blocks=5,
                 combine_layer=3, combine_type="average", use_viewdirs=True, use_code=True,
                 use_code_viewdirs=False, normalize_z=True, use_xyz=True, num_freqs=6,
                 freq_factor=1.5, include_input=True):
        self.latents = latents          # list of (SB*NS, C_i, H_i, W_i)
        self.w2c = w2c                  # (SB*NS, 3, 4)
        self.focal = focal              # (1|SB|SB*NS, 2), fy already negated
        self.c = c                      # (1|SB|SB*NS, 2)
        self.ns = ns
        self.mlp_coarse = mlp_coarse    # state-dict slices
        self.mlp_fine = mlp_fine        # or None
        self.d_latent = d_latent
        self.n_blocks = n_blocks
        self.combine_layer = combine_layer
        self.combine_type = combine_type
        self.use_viewdirs = use_viewdirs
        self.use_code = use_code
        self.use_code_viewdirs = use_code_viewdirs
        self.normalize_z = normalize_z
        self.use_xyz = use_xyz
        self.num_freqs = num_freqs
        self.freq_factor = freq_factor
        self.include_input = include_input


def mlp_input(scene, xyz, viewdirs):
    """
    The (SB*NS*P, d_latent + d_in) rows fed to ResnetFC: models.py.backup2:166-243.
    Row order (sb, ns, p); columns [latent | code(xyz_rot) | viewdir_cam] (or
    code([xyz_rot, viewdir_cam]) when use_code_viewdirs).
    """
    SB, P, _ = xyz.shape
    NS = scene.ns
    R = scene.w2c[:, None, :3, :3]
    xyz_r = xyz.unsqueeze(1).expand(-1, NS, -1, -1).reshape(SB * NS, P, 3)       # :170
    xyz_rot = torch.matmul(R, xyz_r.unsqueeze(-1))[..., 0]                        # :171-173
    xyz_cam = xyz_rot + scene.w2c[:, None, :3, 3]                                 # :174
    if scene.use_xyz:
        zf = (xyz_rot if scene.normalize_z else xyz_cam).reshape(-1, 3)           # :178-182
    else:
        zf = -(xyz_rot if scene.normalize_z else xyz_cam)[..., 2].reshape(-1, 1)  # :183-187
    pe = lambda t: positional_encoding(t, scene.num_freqs, scene.freq_factor, scene.include_input)
    if scene.use_code and not scene.use_code_viewdirs:
        zf = pe(zf)                                                               # :189-191
    if scene.use_viewdirs:
        vd = viewdirs.reshape(SB, P, 3, 1)
        vd = vd.unsqueeze(1).expand(-1, NS, -1, -1, -1).reshape(SB * NS, P, 3, 1)
        vd = torch.matmul(R, vd).reshape(-1, 3)                                   # :197-202
        zf = torch.cat((zf, vd), dim=1)                                           # :203-205
    if scene.use_code and scene.use_code_viewdirs:
        zf = pe(zf)                                                               # :207-209
    uv = -xyz_cam[:, :, :2] / xyz_cam[:, :, 2:]                                   # :215
    uv = uv * _expand_views(scene.focal, SB, NS).unsqueeze(1)                     # :216-218
    uv = uv + _expand_views(scene.c, SB, NS).unsqueeze(1)                         # :219-221
    latent = index_features(scene.latents, uv)                                    # :222-224
    latent = latent.transpose(1, 2).reshape(-1, scene.d_latent)                   # :235-237
    return torch.cat((latent, zf), dim=-1), uv                                    # :243


def net_forward(scene, xyz, coarse=True, viewdirs=None):
    """PixelNeRFNet.forward: (SB,P,3) -> (SB,P,4) = [sigmoid(rgb), relu(sigma)] (:274-281)."""
    SB, P, _ = xyz.shape
    zx, _ = mlp_input(scene, xyz, viewdirs)
    sd = scene.mlp_coarse if (coarse or scene.mlp_fine is None) else scene.mlp_fine     # :258
    out = resnetfc_forward(sd, zx, scene.d_latent, scene.n_blocks, scene.combine_layer,
                           (scene.ns, P), scene.combine_type)
    out = out.reshape(-1, P, 4)
    return torch.cat((torch.sigmoid(out[..., :3]), torch.relu(out[..., 3:4])), dim=-1).reshape(SB, P, 4)


# --------------------------------------------------------------------------------------
# NeRFRenderer                                                         nerf.py:98-303
# --------------------------------------------------------------------------------------


def sample_coarse(rays, n_coarse, lindisp, rand):
    """nerf.py:98-118.  rand = the U[0,1) draw of shape (B,Kc) (RNG #1)."""
    near, far = rays[:, -2:-1], rays[:, -1:]
    step = 1.0 / n_coarse
    z_steps = torch.linspace(0, 1 - step, n_coarse, device=rays.device)
    z_steps = z_steps.unsqueeze(0).repeat(rays.shape[0], 1)
    z_steps = z_steps + rand * step
    if not lindisp:
        return near * (1 - z_steps) + far * z_steps
    return 1 / (1 / near * (1 - z_steps) + 1 / far * z_steps)


def fine_cdf(weights):
    """nerf.py:129-133: pdf=(w+1e-5)/sum, cdf=[0, cumsum] -> (B,Kc+1)."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)


def fine_indices(cdf, u):
    """nerf.py:138-139: searchsorted(right=True)-1 = (number of cdf entries <= u) - 1,
    clamped below at 0 and NOT above (cdf[-1] may round below 1)."""
    inds = torch.searchsorted(cdf, u, right=True).float() - 1.0
    return torch.clamp_min(inds, 0.0)


def sample_fine(rays, weights, n_coarse, lindisp, u, rand):
    """nerf.py:120-148.  u (RNG #3) and rand (RNG #4) have shape (B, Kf-Kfd)."""
    inds = fine_indices(fine_cdf(weights), u)
    z_steps = (inds + rand) / n_coarse
    near, far = rays[:, -2:-1], rays[:, -1:]
    if not lindisp:
        return near * (1 - z_steps) + far * z_steps
    return 1 / (1 / near * (1 - z_steps) + 1 / far * z_steps)


def sample_fine_depth(rays, depth, depth_std, randn):
    """nerf.py:150-161.  randn (RNG #5) has shape (B, Kfd)."""
    z = depth.unsqueeze(1).repeat(1, randn.shape[1])
    z = z + randn * depth_std
    return torch.max(torch.min(z, rays[:, -1:]), rays[:, -2:-1])


def composite_weights(rays, z_samp, out, white_bkgd):
    """
    The arithmetic of NeRFRenderer.composite after the model call (nerf.py:178-182, 223-244).
    out (B,K,4) = [rgb, sigma]  ->  weights (B,K), rgb (B,3), depth (B)
    """
    deltas = z_samp[:, 1:] - z_samp[:, :-1]
    deltas = torch.cat([deltas, rays[:, -1:] - z_samp[:, -1:]], -1)
    rgbs, sigmas = out[..., :3], out[..., 3]
    alphas = 1 - torch.exp(-deltas * torch.relu(sigmas))
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10], -1)
    T = torch.cumprod(shifted, -1)
    weights = alphas * T[:, :-1]
    rgb = torch.sum(weights.unsqueeze(-1) * rgbs, -2)
    depth = torch.sum(weights * z_samp, -1)
    if white_bkgd:
        rgb = rgb + 1 - weights.sum(dim=1).unsqueeze(-1)
    return weights, rgb, depth


def composite(scene, rays, z_samp, coarse, sb, white_bkgd, eval_batch_size=None):
    """NeRFRenderer.composite (nerf.py:163-249) for sb >= 1 objects: points o+z*d (:185),
    viewdirs = ray dirs broadcast over samples (:203-208), one model call (the reference
    only chunks it for memory, :195-216)."""
    B, K = z_samp.shape
    pts = rays[:, None, :3] + z_samp.unsqueeze(2) * rays[:, None, 3:6]
    pts = pts.reshape(sb, -1, 3)
    vd = rays[:, None, 3:6].expand(-1, K, -1).reshape(sb, -1, 3) if scene.use_viewdirs else None
    if eval_batch_size is None:
        out = net_forward(scene, pts, coarse=coarse, viewdirs=vd)
    else:
        chunk = (eval_batch_size - 1) // sb + 1
        outs = []
        for s in range(0, pts.shape[1], chunk):
            outs.append(net_forward(scene, pts[:, s:s + chunk], coarse=coarse,
                                    viewdirs=None if vd is None else vd[:, s:s + chunk]))
        out = torch.cat(outs, dim=1)
    out = out.reshape(B, K, 4)
    return composite_weights(rays, z_samp, out, white_bkgd) + (out,)


class RngTape:
    """
    The renderer's random draws in reference order, shape and dtype:
      #1 rand_like (B,Kc)        nerf.py:111
      #3 rand      (B,Kf-Kfd)    nerf.py:135-137
      #4 rand_like (B,Kf-Kfd)    nerf.py:141
      #5 randn_like(B,Kfd)       nerf.py:158
    Drawn from torch's global generator on ``device`` so that, after the same
    torch.manual_seed, the values equal what the reference itself draws (rand_like(x) and
    rand(x.shape) consume the generator identically).
    """

    def __init__(self, B, n_coarse, n_fine, n_fine_depth, device):
        self.shape = (B, n_coarse, n_fine, n_fine_depth)
        self.device = device
        self.coarse = self.u = self.jit = self.nrm = None

    def draw_coarse(self):
        B, Kc, _, _ = self.shape
        self.coarse = torch.rand(B, Kc, dtype=torch.float32, device=self.device)
        return self.coarse

    def draw_fine(self):
        B, _, Kf, Kfd = self.shape
        if Kf - Kfd > 0:
            self.u = torch.rand(B, Kf - Kfd, dtype=torch.float32, device=self.device)
            self.jit = torch.rand(B, Kf - Kfd, dtype=torch.float32, device=self.device)
        if Kfd > 0:
            self.nrm = torch.randn(B, Kfd, dtype=torch.float32, device=self.device)
        return self.u, self.jit, self.nrm


def render(scene, rays_sb, *, n_coarse=64, n_fine=32, n_fine_depth=16, depth_std=0.01,
           white_bkgd=False, lindisp=False, tape=None, eval_batch_size=None):
    """
    NeRFRenderer.forward (nerf.py:251-303) in eval mode.  rays_sb (SB,B,8).
    Returns a dict with coarse/fine {weights, rgb, depth, z, out} (all flattened over SB*B).
    """
    sb = rays_sb.shape[0]
    rays = rays_sb.reshape(-1, 8)
    B = rays.shape[0]
    if tape is None:
        tape = RngTape(B, n_coarse, n_fine, n_fine_depth, rays.device)
    z_c = sample_coarse(rays, n_coarse, lindisp, tape.draw_coarse())
    w_c, rgb_c, d_c, out_c = composite(scene, rays, z_c, True, sb, white_bkgd, eval_batch_size)
    res = {"coarse": {"weights": w_c, "rgb": rgb_c, "depth": d_c, "z": z_c, "out": out_c}}
    if n_fine > 0:
        u, jit, nrm = tape.draw_fine()
        samps = [z_c]
        if n_fine - n_fine_depth > 0:
            samps.append(sample_fine(rays, w_c, n_coarse, lindisp, u, jit))
        if n_fine_depth > 0:
            samps.append(sample_fine_depth(rays, d_c, depth_std, nrm))
        z_f, _ = torch.sort(torch.cat(samps, dim=-1), dim=-1)
        w_f, rgb_f, d_f, out_f = composite(scene, rays, z_f, False, sb, white_bkgd, eval_batch_size)
        res["fine"] = {"weights": w_f, "rgb": rgb_f, "depth": d_f, "z": z_f, "out": out_f}
    return res


# --------------------------------------------------------------------------------------
# synthetic scene factory shared by tests / smoke / bench (no dataset exists)
# --------------------------------------------------------------------------------------


def pose_spherical(theta, phi, radius):
    """util.pose_spherical (src/util/util.py:314-328): camera-to-world on a sphere."""
    def trans_t(t):
        return torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], dtype=torch.float32)

    def rot_phi(p):
        return torch.tensor([[1, 0, 0, 0], [0, math.cos(p), -math.sin(p), 0],
                             [0, math.sin(p), math.cos(p), 0], [0, 0, 0, 1]], dtype=torch.float32)

    def rot_theta(t):
        return torch.tensor([[math.cos(t), 0, -math.sin(t), 0], [0, 1, 0, 0],
                             [math.sin(t), 0, math.cos(t), 0], [0, 0, 0, 1]], dtype=torch.float32)

    c2w = trans_t(radius)
    c2w = rot_phi(phi / 180.0 * math.pi) @ c2w
    c2w = rot_theta(theta / 180.0 * math.pi) @ c2w
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    return flip @ c2w


def gen_rays(poses, width, height, focal, z_near, z_far, c=None):
    """util.gen_rays + unproj_map (src/util/util.py:118-148, 243-281), ndc=False.
    poses (N,4,4) c2w -> (N,H,W,8) = [origin, unit dir, near, far]."""
    dev = poses.device
    if c is None:
        cx, cy = width * 0.5, height * 0.5
    else:
        cc = torch.as_tensor(c).flatten()
        cx, cy = float(cc[0]), float(cc[1])
    f = torch.as_tensor(focal, dtype=torch.float32).flatten()
    fx, fy = (float(f[0]), float(f[0])) if f.numel() == 1 else (float(f[0]), float(f[1]))
    Y, X = torch.meshgrid(torch.arange(height, dtype=torch.float32) - cy,
                          torch.arange(width, dtype=torch.float32) - cx, indexing="ij")
    X = X.to(dev) / fx
    Y = Y.to(dev) / fy
    unproj = torch.stack((X, -Y, -torch.ones_like(X)), dim=-1)
    unproj = unproj / torch.norm(unproj, dim=-1).unsqueeze(-1)
    N = poses.shape[0]
    dirs = torch.matmul(poses[:, None, None, :3, :3], unproj[None, ..., None])[..., 0]
    cen = poses[:, None, None, :3, 3].expand(-1, height, width, -1)
    near = torch.full((N, height, width, 1), float(z_near), device=dev)
    far = torch.full((N, height, width, 1), float(z_far), device=dev)
    return torch.cat((cen, dirs, near, far), dim=-1)


# ---- per-view metrics of the eval driver (eval/eval.py:314-343) -------------------------------------
# The reference calls skimage.measure.compare_ssim / compare_psnr (scikit-image, NOT vendored in
# /root/reference and not installed here; the `compare_*` names exist up to scikit-image 0.17, no
# version is pinned by the reference).  Restated from the published algorithm of
# skimage.metrics.structural_similarity (Wang et al. 2004 with skimage's defaults): per channel,
# float64, uniform win x win filter, sample covariance (NP/(NP-1)), K1=0.01, K2=0.03, mean of the
# SSIM map with a (win-1)//2 border cropped; multichannel = mean over channels.  compare_psnr =
# 10 log10(data_range^2 / mse).  No golden vector exists in the reference for these ("parity
# unpinned" for this function); tests pin the restatement against a brute-force window loop.
def frame_metrics(rgb, gt, data_range=1.0, win=7):
    """rgb, gt (H,W,C) arrays/tensors; rgb is clamped to [0,1] first (eval.py:290-292).
    :return (psnr, ssim) python floats"""
    import numpy as np
    from scipy.ndimage import uniform_filter

    a = np.clip(np.asarray(rgb, dtype=np.float32), 0.0, 1.0).astype(np.float64)
    b = np.asarray(gt, dtype=np.float32).astype(np.float64)
    mse = np.mean((a - b) ** 2)
    psnr = 10.0 * np.log10(data_range ** 2 / mse)
    npix = win * win
    cov_norm = npix / (npix - 1.0)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    pad = (win - 1) // 2
    vals = []
    for ch in range(a.shape[2]):
        x, y = a[..., ch], b[..., ch]
        ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
        uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
        vals.append(s[pad:-pad, pad:-pad].mean())
    return float(psnr), float(np.mean(vals))
