"""
Multi-GPU plumbing for the ray-rendering path.

Rays are independent (no reduction across rays anywhere in NeRFRenderer.forward,
src/render/nerf.py:251-303), so the only data that has to reach every GPU is the source-view state
left by PixelNeRFNet.encode (feature maps + camera block) and the MLP weights -- once per
encode / weight update -- and per call a slice of the rays in and rgb/depth out.  There is no
per-ray or per-layer communication.

Two front-ends:

* one process per GPU under torch.distributed (``broadcast_scene`` / ``shard_rays`` /
  ``gather_outputs``): rank 0 encodes, the state is broadcast over NCCL (NVLink/NVSwitch), each
  rank renders its contiguous slice of dim 1 -- what bench.py's multi-GPU arm uses;
* single process, several devices (``MultiDeviceRenderer``): what
  ``NeRFRenderer.bind_parallel(net, gpus)`` returns for len(gpus) > 1, the drop-in for the
  reference's ``torch.nn.DataParallel(wrapped, gpus, dim=1)`` (nerf.py:367-371) without its
  per-call parameter broadcast (and without its dependence on ``encoder.latent`` being a buffer,
  which breaks the reference on >1 GPU -- SURVEY.md F4c).
"""
import copy

import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------
# torch.distributed front-end
# ---------------------------------------------------------------------------------------------
def shard_bounds(n, rank, world):
    """Contiguous [lo, hi) slice of n items for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays, rank=None, world=None):
    """(SB, B, 8) -> this rank's contiguous slice along dim 1."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(rays.shape[1], rank, world)
    return rays[:, lo:hi]


def gather_outputs(t, total, dim=1, group=None):
    """All-gather per-rank slices (possibly of unequal length along `dim`) back to `total` items.
    Collective: one ``ncclAllGather`` (``dist.all_gather_into_tensor``) of the slices padded to the longest."""
    world = dist.get_world_size(group)
    if world == 1:
        return t
    if dim != 0:
        return gather_outputs(t.transpose(0, dim).contiguous(), total, 0, group).transpose(0, dim)
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    if t.shape[0] == longest and t.is_contiguous():
        buf = t
    else:
        buf = t.new_zeros((longest,) + tuple(t.shape[1:]))
        buf[: t.shape[0]].copy_(t)
    out = t.new_empty((world * longest,) + tuple(t.shape[1:]))
    dist.all_gather_into_tensor(out, buf, group=group)
    if total == world * longest:
        return out
    return torch.cat([out[r * longest: r * longest + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def render_views(render_par, poses, width, height, focal, z_near, z_far, c=None, ray_batch_size=50000,
                 rank=None, world=None, gather=True, group=None, ray_range=None, host_rays=None, packed=False):
    """The frame loop of the reference's drivers (eval/gen_video.py:174-237, eval/eval.py:250-292)
    with the rays generated where they are rendered: every rank derives ITS contiguous
    range of the NV*H*W rays from (poses, focal, c) -- no (NV,H,W,8) host tensor, no scatter --
    renders it in ``ray_batch_size`` batches through ``render_par`` (``bind_parallel(...,
    simple_output=True)``) and, with ``gather``, all ranks receive the whole frames through ONE
    all-gather of the packed (rgb, depth) rows.

    :param poses (NV,4,4) camera-to-world on the rendering device
    :param ray_range optional (g0, g1): render only rays [g0, g1) of the flattened NV*H*W rays (the
           drivers' ``torch.split(render_rays.view(-1, 8), ray_batch_size)`` batches cross frame borders)
    :param host_rays optional (NV*H*W, 8) CPU tensor (pinned): take the rays from the host like the
           reference's drivers do (``util.gen_rays(...).to(device)``), copying only this rank's slice
    :param packed return the gathered rows as ONE contiguous (n, 4) tensor [r g b depth] (a single D2H copy
           moves a frame to the host)
    :return rgb (NV,H,W,3), depth (NV,H,W); with ``ray_range`` flat (n,3), (n,); without ``gather``
            this rank's flat slices (n,3), (n,) and its [lo, hi) ray range
    """
    from . import util

    distributed = dist.is_available() and dist.is_initialized()
    if rank is None:
        rank = dist.get_rank(group) if distributed else 0
    if world is None:
        world = dist.get_world_size(group) if distributed else 1
    nv, per_frame = poses.shape[0], width * height
    g0, g1 = (0, nv * per_frame) if ray_range is None else ray_range
    total = g1 - g0
    lo, hi = shard_bounds(total, rank, world)
    lo, hi = lo + g0, hi + g0
    dev = poses.device
    parts = []
    if hi > lo:
        if host_rays is not None:
            rays = host_rays[lo:hi].to(dev, non_blocking=True)
        else:
            f0, f1 = lo // per_frame, (hi - 1) // per_frame + 1   # frames this range touches
            rays = util.gen_rays(poses[f0:f1], width, height, focal, z_near, z_far, c=c).reshape(-1, 8)
            rays = rays[lo - f0 * per_frame: hi - f0 * per_frame]
        for batch in torch.split(rays, ray_batch_size, dim=0):
            rgb, depth = render_par(batch[None])
            parts.append(torch.cat((rgb[0], depth[0].unsqueeze(-1)), dim=-1))   # packed (n, 4) rows
    out4 = (parts[0] if len(parts) == 1 else torch.cat(parts)) if parts else torch.zeros(0, 4, device=dev)
    if not gather:
        return out4[:, :3], out4[:, 3], (lo, hi)
    if world > 1:
        out4 = gather_outputs(out4.contiguous(), total, dim=0, group=group)
    if packed:
        return out4
    if ray_range is not None:
        return out4[:, :3], out4[:, 3]
    return out4[:, :3].reshape(nv, height, width, 3), out4[:, 3].reshape(nv, height, width)


def broadcast_scene(net, src=0, group=None):
    """Make every rank's PixelNeRFNet hold rank `src`'s encoded source views and MLP weights.
    Call after ``net.encode`` on `src` (other ranks need not have encoded anything)."""
    rank = dist.get_rank(group)
    device = net.poses.device
    if rank == src:
        maps = [m.detach().float().contiguous() for m in net.encoder.level_maps()]
        meta = dict(shapes=[tuple(m.shape) for m in maps], poses=tuple(net.poses.shape), focal=tuple(net.focal.shape),
                    c=tuple(net.c.shape), num_objs=net.num_objs, ns=net.num_views_per_obj,
                    multi=bool(net.encoder.use_multi_scale))
    else:
        maps, meta = None, None
    box = [meta]
    dist.broadcast_object_list(box, src=src, group=group)
    meta = box[0]
    if rank != src:
        maps = [torch.empty(s, dtype=torch.float32, device=device) for s in meta["shapes"]]
        net.poses = torch.empty(meta["poses"], dtype=torch.float32, device=device)
        net.focal = torch.empty(meta["focal"], dtype=torch.float32, device=device)
        net.c = torch.empty(meta["c"], dtype=torch.float32, device=device)
        net.num_objs, net.num_views_per_obj = meta["num_objs"], meta["ns"]
    net.c = net.c.to(device).float().contiguous()
    for t in maps + [net.poses, net.focal, net.c, net.image_shape]:
        dist.broadcast(t, src=src, group=group)
    with torch.no_grad():
        for p in list(net.mlp_coarse.parameters()) + (list(net.mlp_fine.parameters()) if net.mlp_fine is not None else []):
            dist.broadcast(p.data, src=src, group=group)
    if rank != src:
        net.encoder.latent = maps[-1]
        net.encoder.latents = list(maps) if meta["multi"] else []
        for m in (net.mlp_coarse, net.mlp_fine):
            if m is not None:
                m._native_cache = {}  # weights changed through .data: re-pack the native image lazily
    net.invalidate_scene()
    return net


# ---------------------------------------------------------------------------------------------
# single-process, multi-device front-end
# ---------------------------------------------------------------------------------------------
class MultiDeviceRenderer(torch.nn.Module):
    """forward(rays (SB,B,8), want_weights=False): dim 1 is split across `gpus`; every device runs
    the fused renderer on its own stream; results are concatenated on gpus[0]."""

    def __init__(self, wrapped, gpus):
        super().__init__()
        self.module = wrapped
        self.gpus = [torch.device("cuda", g) if isinstance(g, int) else torch.device(g) for g in gpus]
        self._replicas = {}

    def _replica(self, dev):
        net = self.module.net
        if dev == net.poses.device:
            return self.module
        rep = self._replicas.get(dev)
        key = (net._scene_version, net.mlp_coarse._fingerprint(),
               None if net.mlp_fine is None else net.mlp_fine._fingerprint())
        if rep is None:
            enc, net.encoder = net.encoder, None   # the ResNet encoder is not needed on replicas
            scache, net._scene_cache = net._scene_cache, {}
            ws, net._workspace = net._workspace, None
            try:
                rnet = copy.deepcopy(net)
            finally:
                net.encoder, net._scene_cache, net._workspace = enc, scache, ws
            rnet.encoder = _LatentHolder(enc)
            rnet = rnet.to(dev)
            wrapper = type(self.module)(rnet, self.module.renderer, self.module.simple_output)
            rep = [wrapper, None]
            self._replicas[dev] = rep
        if rep[1] != key:
            rnet = rep[0].net
            with torch.no_grad():
                rnet.mlp_coarse.load_state_dict(net.mlp_coarse.state_dict())
                if net.mlp_fine is not None and rnet.mlp_fine is not None:
                    rnet.mlp_fine.load_state_dict(net.mlp_fine.state_dict())
                elif net.mlp_fine is None:
                    rnet.mlp_fine = None
            rnet.poses = net.poses.to(dev)
            rnet.focal = net.focal.to(dev)
            rnet.c = net.c.to(dev)
            rnet.image_shape = net.image_shape.to(dev)
            rnet.num_objs, rnet.num_views_per_obj = net.num_objs, net.num_views_per_obj
            rnet.precision, rnet.texel_scale = net.precision, net.texel_scale
            rnet.encoder.set_maps([m.to(dev) for m in net.encoder.level_maps()])
            rnet.invalidate_scene()
            rep[1] = key
        return rep[0]

    def forward(self, rays, want_weights=False):
        n = rays.shape[1]
        primary = self.gpus[0]
        parts = []
        renderer = self.module.renderer
        tape, renderer.rng_tape = renderer.rng_tape, None   # a pre-drawn tape covers the whole batch: split it per shard
        for i, dev in enumerate(self.gpus):
            lo, hi = shard_bounds(n, i, len(self.gpus))
            if hi == lo:
                continue
            with torch.cuda.device(dev):
                wrapper = self._replica(dev)
                if tape is not None:
                    assert rays.shape[0] == 1, "rng_tape with several devices needs SB == 1 (rows are ray-major)"
                    renderer.rng_tape = {k: v[lo:hi] for k, v in tape.items()}
                parts.append(wrapper(rays[:, lo:hi].to(dev, non_blocking=True), want_weights=want_weights))
        if len(parts) == 0:
            return self.module(rays, want_weights=want_weights)

        def cat(items):
            return torch.cat([t.to(primary, non_blocking=True) for t in items], dim=1)

        if isinstance(parts[0], tuple):
            return tuple(cat([p[k] for p in parts]) for k in range(len(parts[0])))
        out = {}
        for lvl in parts[0]:
            out[lvl] = {k: cat([p[lvl][k] for p in parts]) for k in parts[0][lvl]}
        return out


class _LatentHolder(torch.nn.Module):
    """Stand-in for the image encoder on render replicas: holds the per-level feature maps only."""

    def __init__(self, enc):
        super().__init__()
        self.use_multi_scale = enc.use_multi_scale
        self.latent_size = enc.latent_size
        self.latent, self.latents = None, []

    def set_maps(self, maps):
        self.latent = maps[-1]
        self.latents = list(maps) if self.use_multi_scale else []

    def level_maps(self):
        return list(self.latents) if self.use_multi_scale else [self.latent]
