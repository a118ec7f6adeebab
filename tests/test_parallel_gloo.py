"""N>1 host logic on CPU: world_size-2 gloo.  Rank 0 'encodes', the scene state and MLP weights are
broadcast, rays are sharded along dim 1 and outputs gathered back in order."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import REPO


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pixel_nerf_multiscale_b200 as pk
    from pixel_nerf_multiscale_b200.parallel import broadcast_scene, gather_outputs, shard_rays
    from pixel_nerf_multiscale_b200.util.conf import ConfigFactory

    conf = ConfigFactory.parse_file(os.path.join(REPO, "conf", "exp", "sn64_multiscale.conf"))
    conf["model"]["encoder"].put("pretrained", False)
    torch.manual_seed(100 + rank)  # different weights per rank before the broadcast
    net = pk.make_model(conf["model"]).eval()
    if rank == 0:
        with torch.no_grad():
            poses = torch.stack([pk.util.pose_spherical(30.0 * i, -20.0, 2.6) for i in range(2)])[None]
            net.encode(torch.rand(1, 2, 3, 32, 32), poses, torch.tensor(40.0))
    broadcast_scene(net, src=0)
    sig = [float(net.mlp_coarse.lin_in.weight.sum()), float(net.mlp_fine.blocks[2].fc_0.bias.abs().sum() + net.mlp_fine.lin_out.weight.sum()),
           float(net.poses.sum()), float(net.focal.sum()), float(net.c.sum()), net.num_views_per_obj,
           [tuple(m.shape) for m in net.encoder.level_maps()], float(sum(m.sum() for m in net.encoder.level_maps()))]
    rays = torch.arange(1 * 11 * 8, dtype=torch.float32).reshape(1, 11, 8)
    mine = shard_rays(rays)
    back = gather_outputs(mine[..., :3].contiguous(), 11)
    # render_views: each rank generates and "renders" only its own ray range; frames come back whole
    from pixel_nerf_multiscale_b200.parallel import render_views

    def fake_render(r):  # deterministic function of the ray record, shapes of _RenderWrapper(simple_output=True)
        return r[..., 3:6] * 0.5 + r[..., :3], r[..., 6] + r[..., 3]

    vposes = torch.stack([pk.util.pose_spherical(40.0 * i, -10.0, 2.0) for i in range(3)])
    rgb, depth = render_views(fake_render, vposes, 7, 5, 30.0, 0.5, 3.0, ray_batch_size=13)
    full = pk.util.gen_rays(vposes, 7, 5, 30.0, 0.5, 3.0)
    exp_rgb, exp_depth = fake_render(full)
    frames_ok = rgb.shape == (3, 5, 7, 3) and torch.equal(rgb, exp_rgb) and torch.equal(depth, exp_depth)
    _, _, (lo, hi) = render_views(fake_render, vposes, 7, 5, 30.0, 0.5, 3.0, gather=False)
    # a ray batch that crosses frame borders (the drivers' torch.split batches), device-generated and host-fed
    flat = full.reshape(-1, 8)
    for host in (None, flat):
        rr, dd = render_views(fake_render, vposes, 7, 5, 30.0, 0.5, 3.0, ray_batch_size=13, ray_range=(10, 81),
                              host_rays=host)
        frames_ok = frames_ok and torch.equal(rr, exp_rgb.reshape(-1, 3)[10:81]) and torch.equal(dd, exp_depth.reshape(-1)[10:81])
    q.put((rank, sig, mine.shape[1], torch.equal(back, rays[..., :3]) and frames_ok and (hi - lo) in (52, 53)))
    dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, n0, ok0), (r1, s1, n1, ok1) = res
    assert s0 == s1, "ranks disagree after broadcast_scene"
    assert s0[5] == 2 and len(s0[6]) == 4
    assert (n0, n1) == (6, 5) and ok0 and ok1
