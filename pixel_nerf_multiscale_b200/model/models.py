"""
PixelNeRFNet: drop-in for the reference's functional model (src/model/models.py.backup2 -- the
live src/model/models.py cannot be constructed with any shipped conf).

Same constructor, attributes, ``encode`` / ``forward`` / ``load_weights`` / ``save_weights``
signatures and state-dict keys.  ``encode`` stays in PyTorch (ResNet encoder + camera
bookkeeping); ``forward`` runs the fused native kernels:

    (a) csrc/features.*  camera transform, projection, multiscale bilinear gather, positional code
    (b) csrc/mlp_*.cu    ResnetFC with latent injection, residual blocks, view mean-pool, head

There is no PyTorch/CPU fallback for ``forward``: CPU tensors or an unbuilt extension raise.
"""
import os
import os.path as osp
import warnings

import torch

from .. import _native as N
from .code import PositionalEncoding
from .encoder import ImageEncoder
from .model_util import make_encoder, make_mlp

_PRECISIONS = {"bf16": N.BF16, "fp32": N.FP32, "fp16": N.FP16}
_TORCH_DTYPE = {N.FP32: torch.float32, N.BF16: torch.bfloat16, N.FP16: torch.float16}


class PixelNeRFNet(torch.nn.Module):
    def __init__(self, conf, stop_encoder_grad=False):
        """:param conf config subtree 'model' (schema of conf/default.conf)"""
        super().__init__()
        self.encoder = make_encoder(conf["encoder"])
        self.use_encoder = conf.get_bool("use_encoder", True)
        self.use_xyz = conf.get_bool("use_xyz", False)
        assert self.use_encoder or self.use_xyz  # must use some feature
        self.normalize_z = conf.get_bool("normalize_z", True)
        self.stop_encoder_grad = stop_encoder_grad
        self.use_code = conf.get_bool("use_code", False)
        self.use_code_viewdirs = conf.get_bool("use_code_viewdirs", True)
        self.use_viewdirs = conf.get_bool("use_viewdirs", False)
        self.use_global_encoder = conf.get_bool("use_global_encoder", False)
        if not self.use_encoder:
            raise NotImplementedError("use_encoder=False is not supported by the native path")
        if self.use_global_encoder:
            raise NotImplementedError("use_global_encoder is not supported by the native path")

        d_latent = self.encoder.latent_size
        d_in = 3 if self.use_xyz else 1
        if self.use_viewdirs and self.use_code_viewdirs:
            d_in += 3
        if self.use_code and d_in > 0:
            self.code = PositionalEncoding.from_conf(conf["code"], d_in=d_in)
            d_in = self.code.d_out
        if self.use_viewdirs and not self.use_code_viewdirs:
            d_in += 3
        d_out = 4
        # the fork keeps d_latent as the per-level list for multi-scale encoders and the
        # summed width in latent_size (models.py.backup2:48,70-78)
        self.latent_size = sum(int(x) for x in d_latent) if isinstance(d_latent, (list, tuple)) else int(d_latent)

        self.mlp_coarse = make_mlp(conf["mlp_coarse"], d_in, d_latent, d_out=d_out)
        self.mlp_fine = make_mlp(conf["mlp_fine"], d_in, d_latent, d_out=d_out, allow_empty=True)
        # world -> camera, bottom row omitted
        self.register_buffer("poses", torch.empty(1, 3, 4), persistent=False)
        self.register_buffer("image_shape", torch.empty(2), persistent=False)
        self.d_in, self.d_out, self.d_latent = d_in, d_out, d_latent
        self.register_buffer("focal", torch.empty(1, 2), persistent=False)
        self.register_buffer("c", torch.empty(1, 2), persistent=False)
        self.num_objs = 0
        self.num_views_per_obj = 1

        # native-path state
        # "fp16" (default) | "bf16": tensor-core path with f16 / bf16 operands; "fp32": validation path
        self.precision = os.environ.get("PIXELNERF_B200_PRECISION", "fp16")
        self.texel_scale = None  # None = the fork's pixel==texel behaviour
        self._scene_cache = {}
        self._scene_version = 0
        self._workspace = None

    # ------------------------------------------------------------------------------------------
    def encode(self, images, poses, focal, z_bounds=None, c=None):
        """
        :param images (NS,3,H,W) or (SB,NS,3,H,W) source views
        :param poses (NS,4,4) or (SB,NS,4,4) camera-to-world
        :param focal () | (2) | (N) | (N,2)   :param c None | () | (2) | (N) | (N,2)
        """
        # the fork's eval/gen_video.py:204 passes the source images on the CPU while the model lives on the
        # GPU (upstream moves them first); accept that instead of failing inside the first convolution
        dev = self.poses.device
        images, poses, focal = images.to(dev), poses.to(dev), focal.to(dev)
        c = c if c is None else c.to(dev)
        self.num_objs = images.size(0)
        if images.dim() == 5:
            assert poses.dim() == 4
            assert poses.size(1) == images.size(1)  # consistent number of source views
            self.num_views_per_obj = images.size(1)
            images = images.reshape(-1, *images.shape[2:])
            poses = poses.reshape(-1, 4, 4)
        else:
            self.num_views_per_obj = 1
        self.encoder(images)
        rot = poses[:, :3, :3].transpose(1, 2)
        trans = -torch.bmm(rot, poses[:, :3, 3:])
        self.poses = torch.cat((rot, trans), dim=-1)
        self.image_shape[0] = images.shape[-1]
        self.image_shape[1] = images.shape[-2]
        if focal.dim() == 0:
            focal = focal[None, None].repeat((1, 2))
        elif focal.dim() == 1:
            focal = focal.unsqueeze(-1).repeat((1, 2))
        else:
            focal = focal.clone()
        self.focal = focal.float()
        self.focal[..., 1] *= -1.0  # image y points down, camera y up
        if c is None:
            c = (self.image_shape * 0.5).unsqueeze(0)
        elif c.dim() == 0:
            c = c[None, None].repeat((1, 2))
        elif c.dim() == 1:
            c = c.unsqueeze(-1).repeat((1, 2))
        self.c = c
        self.invalidate_scene()

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_scene_cache"], state["_workspace"] = {}, None  # caches of device pointers: never copied
        return state

    def invalidate_scene(self):
        """Drop the packed feature pyramid / camera block (after encode() or after the caller
        replaces ``encoder.latent(s)``)."""
        self._scene_cache = {}
        self._scene_version += 1

    # ------------------------------------------------------------------------------------------
    def _native_precision(self, precision=None):
        p = self.precision if precision is None else precision
        if p not in _PRECISIONS:
            raise ValueError("precision must be 'bf16', 'fp16' or 'fp32', got %r" % (p,))
        return _PRECISIONS[p]

    def _per_view(self, t, n_views):
        """focal / c rows -> one row per (object, view)  (models.py.backup2:216-221)."""
        t = t.to(self.poses.device).float()
        if t.shape[0] == 1:
            return t.expand(n_views, -1)
        if t.shape[0] == n_views:
            return t
        ns = self.num_views_per_obj
        assert t.shape[0] * ns == n_views, "focal/c rows (%d) do not match the encoded objects" % t.shape[0]
        return t.unsqueeze(1).expand(-1, ns, -1).reshape(n_views, -1)

    def native_scene(self, precision=None):
        """ctypes Scene descriptor + keepalive list for the currently encoded source views."""
        prec = self._native_precision(precision)
        hit = self._scene_cache.get(prec)
        if hit is not None:
            return hit
        maps = self.encoder.level_maps()
        if len(maps) == 0 or maps[0] is None:
            raise RuntimeError("PixelNeRFNet.forward called before encode()")
        # the native gather implements grid_sample(bilinear, border, align_corners=True) only
        # (src/model/encoder.py:182-188 with the shipped confs); other conf values would render differently
        interp = getattr(self.encoder, "index_interp", "bilinear")
        padding = getattr(self.encoder, "index_padding", "border")
        if interp != "bilinear" or padding != "border":
            raise NotImplementedError("encoder.index_interp=%r / index_padding=%r are not supported by the native "
                                      "gather (only bilinear / border)" % (interp, padding))
        device = self.poses.device
        if device.type != "cuda":
            raise RuntimeError("pixelnerf_b200: the rendering path needs the model on a CUDA device (no CPU path)")
        n_views = self.poses.shape[0]
        keep = []
        sc = N.Scene()
        sc.n_views, sc.ns, sc.n_levels = n_views, self.num_views_per_obj, len(maps)
        sc.feat_dtype = prec
        lib = N.lib()
        off = 0
        with torch.cuda.device(device):
            for i, fm in enumerate(maps):
                fm = fm.detach().float().contiguous()
                v, ch, h, w = fm.shape
                assert v == n_views, "feature maps have %d views, cameras %d" % (v, n_views)
                packed = torch.empty((v, h, w, ch), dtype=_TORCH_DTYPE[prec], device=device)
                N.check(lib.pnr_pack_level(N.ptr(fm), v, ch, h, w, N.ptr(packed), prec, N.stream_ptr(device)),
                        "pnr_pack_level")
                keep += [fm, packed]
                sc.C[i], sc.H[i], sc.W[i], sc.ch_off[i] = ch, h, w, off
                sx, sy = (1.0, 1.0) if self.texel_scale is None else self.texel_scale(i, (h, w))
                sc.kx[i], sc.ky[i] = sx, sy
                sc.level[i] = N.ptr(packed)
                off += ch
        sc.d_latent = off
        assert off == self.latent_size, "encoder produced %d channels, model expects %d" % (off, self.latent_size)
        cams = torch.cat((self.poses[:, :, :3].reshape(n_views, 9), self.poses[:, :, 3],
                          self._per_view(self.focal, n_views), self._per_view(self.c, n_views)), dim=1)
        cams = cams.float().contiguous()
        keep.append(cams)
        sc.cams = N.ptr(cams)
        sc.use_xyz, sc.normalize_z = int(self.use_xyz), int(self.normalize_z)
        sc.use_viewdirs, sc.use_code = int(self.use_viewdirs), int(self.use_code)
        sc.use_code_viewdirs = int(self.use_code_viewdirs)
        if self.use_code:
            sc.num_freqs, sc.include_input = self.code.num_freqs, int(self.code.include_input)
            sc.freq_factor = self.code.freq_factor
        sc.d_in = self.d_in
        self._scene_cache[prec] = (sc, keep)
        return sc, keep

    def native_mlp(self, coarse=True, precision=None):
        prec = self._native_precision(precision)
        mlp = self.mlp_coarse if (coarse or self.mlp_fine is None) else self.mlp_fine
        return mlp.native(prec)

    def workspace(self, nbytes, device):
        """Grow-only scratch buffer (torch owns all memory handed to the C ABI)."""
        ws = self._workspace
        if ws is None or ws.numel() < nbytes or ws.device != device:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self._workspace = ws
        return ws

    def forward(self, xyz, coarse=True, viewdirs=None, far=False):
        """
        (r, g, b, sigma) at world-space points.  Call encode() first.
        :param xyz (SB, B, 3)   :param viewdirs (SB, B, 3) when use_viewdirs
        :return (SB, B, 4) [sigmoid(rgb), relu(sigma)]
        """
        SB, B, _ = xyz.shape
        if not xyz.is_cuda:
            raise RuntimeError("pixelnerf_b200: PixelNeRFNet.forward needs CUDA tensors (no CPU path)")
        assert SB == self.num_objs or SB * self.num_views_per_obj == self.poses.shape[0], \
            "xyz has %d objects, encode() saw %d" % (SB, self.num_objs)
        if self.use_viewdirs:
            assert viewdirs is not None
            viewdirs = viewdirs.reshape(SB, B, 3).float().contiguous()
        else:
            viewdirs = None
        xyz = xyz.float().contiguous()
        prec = self._native_precision()
        sc, _keep_s = self.native_scene()
        m, _keep_m = self.native_mlp(coarse)
        out = torch.empty(SB, B, self.d_out, dtype=torch.float32, device=xyz.device)
        lib = N.lib()
        # same scope name as models.py.backup2:165
        with torch.autograd.profiler.record_function("model_inference"), torch.cuda.device(xyz.device):
            nbytes = lib.pnr_net_forward_workspace(sc, m, SB, B, prec)
            ws = self.workspace(nbytes, xyz.device)
            N.check(lib.pnr_net_forward(sc, m, N.ptr(xyz), N.ptr(viewdirs), SB, B, prec, N.ptr(out), N.ptr(ws),
                                        ws.numel(), N.stream_ptr(xyz.device)), "pnr_net_forward")
        return out

    # ------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts the model's own keys, the fork trainer's checkpoint dict, DataParallel saves and
        upstream pixelNeRF checkpoints (model/checkpoint.py)."""
        from .checkpoint import normalize_state_dict

        return super().load_state_dict(normalize_state_dict(state_dict, self.state_dict()), strict=strict, **kw)

    def load_weights(self, args, opt_init=False, strict=True, device=None):
        """Loads checkpoints/<name>/pixel_nerf_{latest,init} like the reference
        (models.py.backup2:284-314); returns self."""
        if opt_init and not args.resume:
            return
        ckpt_name = "pixel_nerf_init" if opt_init or not args.resume else "pixel_nerf_latest"
        model_path = "%s/%s/%s" % (args.checkpoints_path, args.name, ckpt_name)
        if device is None:
            device = self.poses.device
        if os.path.exists(model_path):
            print("Load", model_path)
            self.load_state_dict(torch.load(model_path, map_location=device), strict=strict)
        elif not opt_init:
            warnings.warn("WARNING: {} does not exist, not loaded!! Model will be re-initialized.".format(model_path))
        return self

    def save_weights(self, args, opt_init=False):
        from shutil import copyfile

        ckpt_name = "pixel_nerf_init" if opt_init else "pixel_nerf_latest"
        backup_name = "pixel_nerf_init_backup" if opt_init else "pixel_nerf_backup"
        ckpt_path = osp.join(args.checkpoints_path, args.name, ckpt_name)
        if osp.exists(ckpt_path):
            copyfile(ckpt_path, osp.join(args.checkpoints_path, args.name, backup_name))
        torch.save(self.state_dict(), ckpt_path)
        return self
