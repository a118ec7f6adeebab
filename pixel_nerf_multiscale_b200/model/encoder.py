"""
Spatial (pixel-aligned) image encoder.  Stays in PyTorch -- out of scope for the native
kernels (BASELINE.json north_star) -- but its OUTPUT LAYOUT is the input of the fused
gather kernel, so it mirrors the fork's src/model/encoder.py exactly:

* single-scale: only the LAST ResNet stage is kept (256 ch for resnet34/num_layers=4),
* ``use_multi_scale``: a LIST of per-stage maps at their native resolutions,
* ``latent`` / ``latents`` are plain attributes,
* module/parameter names (``model.*`` and the aliasing ``layers.*``) match the fork so its
  checkpoints load with strict=True.

``index`` (bilinear/border/align_corners gather in which the PIXEL coordinate is used as the
TEXEL coordinate of every level -- SURVEY.md F4b) is kept for API compatibility; the renderer
does not call it, the same arithmetic runs inside csrc/features.cuh.
"""
import torch
import torch.nn.functional as F
from torch import nn

_STAGE_CHANNELS = {"resnet18": [64, 64, 128, 256, 512], "resnet34": [64, 64, 128, 256, 512],
                   "resnet50": [64, 256, 512, 1024, 2048]}


class SpatialEncoder(nn.Module):
    def __init__(self, backbone="resnet34", pretrained=True, num_layers=4, index_interp="bilinear",
                 index_padding="border", upsample_interp="bilinear", feature_scale=1.0, use_first_pool=True,
                 norm_type="batch", use_multi_scale=False):
        super().__init__()
        import torchvision

        if backbone not in _STAGE_CHANNELS:
            raise NotImplementedError("Backbone %s not supported" % backbone)
        self.use_multi_scale = use_multi_scale
        self.num_layers = num_layers
        self.feature_scale = feature_scale
        self.use_first_pool = use_first_pool
        self.index_interp = index_interp
        self.index_padding = index_padding
        self.upsample_interp = upsample_interp
        self.align_corners = True if index_interp == "bilinear" else None
        weights = "DEFAULT" if pretrained else None
        self.model = getattr(torchvision.models, backbone)(weights=weights)
        stem = [self.model.conv1, self.model.bn1, self.model.relu]
        if use_first_pool:
            stem.append(self.model.maxpool)
        stages = [nn.Sequential(*stem)]
        for i, name in enumerate(("layer1", "layer2", "layer3", "layer4")):
            if num_layers > i + 1:
                stages.append(getattr(self.model, name))
        self.layers = nn.ModuleList(stages)
        chans = _STAGE_CHANNELS[backbone][:num_layers]
        self.latent_size = chans if use_multi_scale else chans[-1]
        self.latent = None
        self.latents = []

    def forward(self, x):
        x = x * self.feature_scale
        feats = []
        for stage in self.layers:
            x = stage(x)
            feats.append(x)
        self.latent = feats[-1]
        if self.use_multi_scale:
            self.latents = feats
            return feats
        return x

    def level_maps(self):
        """The maps the point-feature gather reads: one per pyramid level."""
        return list(self.latents) if self.use_multi_scale else [self.latent]

    def index(self, uv, cam_z=None, image_size=(), z_bounds=None):
        """(B,N,2) pixel coords -> (B,L,N).  Compatibility path (torch ops)."""
        maps = self.level_maps()
        if uv.shape[0] == 1 and maps[0].shape[0] > 1:
            uv = uv.expand(maps[0].shape[0], -1, -1)
        outs = []
        for fmap in maps:
            h, w = fmap.shape[-2:]
            scale = torch.tensor([w - 1, h - 1], dtype=uv.dtype, device=uv.device)
            grid = (uv / scale * 2 - 1).unsqueeze(1)
            s = F.grid_sample(fmap, grid, align_corners=self.align_corners, mode=self.index_interp,
                              padding_mode=self.index_padding)
            outs.append(s.squeeze(2))
        return torch.cat(outs, dim=1)

    @classmethod
    def from_conf(cls, conf, **kwargs):
        g = conf.get
        return cls(backbone=g("backbone", "resnet34"), pretrained=g("pretrained", True),
                   num_layers=g("num_layers", 4), index_interp=g("index_interp", "bilinear"),
                   index_padding=g("index_padding", "border"), upsample_interp=g("upsample_interp", "bilinear"),
                   feature_scale=g("feature_scale", 1.0), use_first_pool=g("use_first_pool", True),
                   norm_type=g("norm_type", "batch"), use_multi_scale=g("use_multi_scale", False), **kwargs)


ImageEncoder = SpatialEncoder
