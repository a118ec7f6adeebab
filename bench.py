#!/usr/bin/env python
"""
bench.py -- rays/s of the ray-rendering hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4] [--impl ours|reference]

A "step" renders one ray batch (default 50 000 rays = the reference's --ray_batch_size,
src/util/args.py:19) through render_par(rays) = NeRFRenderer.bind_parallel(net)(rays): 64 coarse +
32 fine (16 importance + 16 depth) samples per ray, both MLPs.  Workloads (synthetic, seeded;
SURVEY.md section 8d -- there is no dataset and no checkpoint):
    c1  sn64  64x64   1 source view   conf/exp/sn64.conf  (single-scale, L=256)
    c2  SRN   128x128 2 source views  conf/exp/srn.conf   (single-scale, L=256)     [default, N=1]
    c3  DTU   300x400 3 source views  conf/exp/dtu.conf   (single-scale, L=256)
    c4  DTU   300x400 3 source views  dtu.conf + encoder.use_multi_scale (pyramid, L=512)
Source images are random, the ResNet34 encoder is random-init (eval mode; each feature level rescaled
to unit RMS, as a trained encoder would emit), the two ResnetFC MLPs are re-randomised (at default
init every block is the identity).

Printed JSON (one line, rank 0): value = whole-job rays/s with rays resident in HBM; e2e = the same
through the public API with rays in pinned host memory and rgb/depth read back every step;
roofline = the dominant kernel (fused ResnetFC phase A) timed with CUDA events inside the timed
region against the measured dense-bf16 peak; cpu_baseline = the oracle (a torch restatement of the
reference's path, oracle/pixelnerf_oracle.py) on the host cores on a bounded ray sample.
--impl reference times that CPU path alone (the reference is Python and cannot travel to the
GPU box; the oracle port is pinned against it by tests/golden).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    "c1": dict(conf="conf/exp/sn64.conf", H=64, W=64, ns=1, focal=119.4, c=None, z_near=1.2, z_far=4.0, radius=2.6,
               multi_scale=False, rays=4096, desc="sn64 64x64 NS=1 single-scale L=256"),
    "c2": dict(conf="conf/exp/srn.conf", H=128, W=128, ns=2, focal=131.25, c=None, z_near=0.8, z_far=1.8, radius=1.3,
               multi_scale=False, rays=50000, desc="SRN car 128x128 NS=2 (views 64 104) single-scale L=256, 40-frame orbit"),
    "c3": dict(conf="conf/exp/dtu.conf", H=300, W=400, ns=3, focal=(723.0, 723.0), c=(200.0, 150.0), z_near=0.1,
               z_far=5.0, radius=2.2, multi_scale=False, rays=50000, desc="DTU 300x400 NS=3 (22 25 28) single-scale L=256"),
    "c4": dict(conf="conf/exp/dtu.conf", H=300, W=400, ns=3, focal=(723.0, 723.0), c=(200.0, 150.0), z_near=0.1,
               z_far=5.0, radius=2.2, multi_scale=True, rays=50000, desc="DTU 300x400 NS=3 multi-scale pyramid L=512"),
}


def rerandomise(mlp, seed):
    """Seeded re-randomisation of one ResnetFC: fan-in scaled weights everywhere,
    small biases, positive density offset."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in mlp.named_parameters():
            if name.endswith("weight"):
                gain = 0.5 if "fc_1" in name else (0.7 if "lin_z" in name else (0.25 if "lin_out" in name else 1.0))
                p.copy_(torch.randn(p.shape, generator=g) * gain * math.sqrt(2.0 / p.shape[1]))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
        mlp.lin_out.bias[3] = 1.0


def build_scene(wl, device, precision):
    import pixel_nerf_multiscale_b200 as pk
    from pixel_nerf_multiscale_b200.util.conf import ConfigFactory

    conf = ConfigFactory.parse_file(os.path.join(REPO, wl["conf"]))
    conf["model"]["encoder"].put("pretrained", False)
    conf["model"]["encoder"].put("use_multi_scale", bool(wl["multi_scale"]))
    torch.manual_seed(0)
    net = pk.make_model(conf["model"]).eval()
    rerandomise(net.mlp_coarse, 1)
    rerandomise(net.mlp_fine, 2)
    net = net.to(device)
    net.precision = precision
    g = torch.Generator().manual_seed(7)
    images = (torch.rand(1, wl["ns"], 3, wl["H"], wl["W"], generator=g) * 2 - 1).to(device)
    poses = torch.stack([pk.util.pose_spherical(30.0 * i, -20.0, wl["radius"]) for i in range(wl["ns"])])[None].to(device)
    focal = torch.tensor(wl["focal"], dtype=torch.float32)
    focal = focal[None] if focal.dim() == 1 else focal
    c = None if wl["c"] is None else torch.tensor(wl["c"], dtype=torch.float32)[None]
    with torch.no_grad():
        net.encode(images, poses, focal.to(device), c=None if c is None else c.to(device))
        # A random-init ResNet34 in eval mode has identity BatchNorm, so its activations grow ~10x per
        # stage (std 0.1 -> 30); a trained encoder emits O(1) features.  Rescale every level to unit RMS
        # so the synthetic scene has non-degenerate density (otherwise sigma saturates to 0 or huge).
        for m in net.encoder.level_maps():
            m.div_(m.pow(2).mean().sqrt().clamp_min(1e-6))
        net.invalidate_scene()
    renderer = pk.NeRFRenderer.from_conf(conf["renderer"], lindisp=False, eval_batch_size=wl["rays"])
    return net, renderer, conf, dict(poses=poses, focal=focal, c=c)


def orbit_rays(wl, cam, n_frames, device):
    """Rays of a 360-degree orbit (eval/gen_video.py:157-183 shape), flattened to (n, 8)."""
    import pixel_nerf_multiscale_b200 as pk

    angles = torch.linspace(-180, 180, n_frames + 1)[:-1]
    poses = torch.stack([pk.util.pose_spherical(float(a), -10.0, wl["radius"]) for a in angles]).to(device)
    rays = pk.util.gen_rays(poses, wl["W"], wl["H"], cam["focal"], wl["z_near"], wl["z_far"], cam["c"])
    return rays.reshape(-1, 8)


def oracle_scene(net, cam, conf, device="cpu"):
    """Oracle Scene sharing this net's weights / feature maps / cameras (CPU copies)."""
    from oracle import pixelnerf_oracle as po
    from oracle import synth

    hp = synth.model_hparams(conf["model"])
    lat = [t.detach().float().to(device) for t in net.encoder.level_maps()]
    sd = lambda m: {k: v.detach().float().to(device) for k, v in m.state_dict().items()}
    n_views = net.poses.shape[0]
    return po.Scene(lat, net.poses.detach().float().to(device), net._per_view(net.focal, n_views).to(device),
                    net._per_view(net.c, n_views).to(device), net.num_views_per_obj, sd(net.mlp_coarse),
                    sd(net.mlp_fine), d_latent=net.latent_size, **hp)


def cpu_render_rate(scene, conf, rays_cpu, steps, warmup):
    """rays/s of the oracle port on the host cores (all threads)."""
    from oracle import pixelnerf_oracle as po

    r = conf["renderer"]
    kw = dict(n_coarse=r.get_int("n_coarse", 128), n_fine=r.get_int("n_fine", 0), n_fine_depth=r.get_int("n_fine_depth", 0),
              depth_std=r.get_float("depth_std", 0.01), white_bkgd=bool(r.get_float("white_bkgd", False)), lindisp=False,
              eval_batch_size=50000)
    n = rays_cpu.shape[0]
    with torch.no_grad():
        for _ in range(warmup):
            po.render(scene, rays_cpu[None, : max(64, n // 8)], **kw)
        t0 = time.perf_counter()
        for i in range(steps):
            po.render(scene, rays_cpu[None], **kw)
        dt = time.perf_counter() - t0
    return steps * n / dt, dt / steps


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], None, set(), []
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = float(s[1])
                pw.append(float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        busy = [x for x in sm if mx and x > 0.5 * mx] or sm
        pw.sort()
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_median": pw[len(pw) // 2] if pw else None, "power_w_max": pw[-1] if pw else None}


_json_out = sys.stdout


def main():
    # stdout carries exactly ONE line, the JSON: everything else any library writes to fd 1 (NCCL prints its
    # version there on multi-GPU runs) is sent to stderr; the JSON goes to a duplicate of the original stdout
    global _json_out
    sys.stdout.flush()
    _json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # same workload at every N so that the per-N values form one (weak) scaling series
    wl_name = args.workload or "c2"
    wl = WORKLOADS[wl_name]
    metric, unit = "rays/sec", "rays/s"
    config = {"workload": "%s: %s; %d rays/step, 64 coarse + 32 fine (16 depth) samples, 2 MLPs" % (wl_name, wl["desc"], wl["rays"]),
              "rays_per_step": wl["rays"], "points_per_ray": 160, "precision": args.precision}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        net, renderer, conf, cam = build_scene(wl, "cpu", "fp32")
        scene = oracle_scene(net, cam, conf)
        sample = 256
        rays = orbit_rays(wl, cam, 1, "cpu")
        pick = torch.randperm(rays.shape[0], generator=torch.Generator().manual_seed(3))[:sample]
        torch.manual_seed(123)
        rate, sec = cpu_render_rate(scene, conf, rays[pick], max(1, args.steps), max(1, min(args.warmup, 1)))
        cores = torch.get_num_threads()
        what = "%d-ray sample of %s per step, torch %s CPU, %d threads" % (sample, wl_name, torch.__version__, cores)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": rate, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": rate, "unit": unit, "cores": cores, "kind": "port", "sample": what},
            "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), file=_json_out, flush=True)
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch.distributed as dist

    from pixel_nerf_multiscale_b200 import _native as N

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one line (the JSON); NCCL's version / debug lines go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    net, renderer, conf, cam = build_scene(wl, device, args.precision)
    if world > 1:
        # source-view state is produced once (rank 0) and broadcast over NVLink; no per-ray traffic
        from pixel_nerf_multiscale_b200.parallel import broadcast_scene

        broadcast_scene(net, src=0)
    render_par = renderer.bind_parallel(net, [local_rank], simple_output=True).eval()
    n_frames = 40 if wl_name == "c2" else (1 if wl_name == "c1" else 2)
    all_rays = orbit_rays(wl, cam, n_frames, device)
    B = wl["rays"]
    n_batches = max(1, all_rays.shape[0] // B)
    batches = [all_rays[i * B:(i + 1) * B].contiguous() for i in range(n_batches)]
    # rank r renders its own slice of the batch list (rays are independent: no data-path collective)
    mine = [batches[(rank + i * world) % n_batches] for i in range(args.steps + args.warmup)]
    lib = N.lib()

    def sync_all():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    torch.manual_seed(100 + rank)
    with torch.no_grad():
        for i in range(args.warmup):
            render_par(mine[i][None])
        N.check(lib.pnr_tc_check(N.stream_ptr(device)), "pnr_tc_check")
        sync_all()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        lib.pnr_launch_count(1)
        lib.pnr_profile_begin()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.steps):
            rgb, depth = render_par(mine[args.warmup + i][None])
        ev1.record()
        sync_all()
        ms = ev0.elapsed_time(ev1)
        launches = int(lib.pnr_launch_count(1))
        import ctypes as C

        pms, pl, pf, pb = (C.c_double * 3)(), (C.c_int64 * 3)(), (C.c_double * 3)(), (C.c_double * 3)()
        lib.pnr_profile_end(pms, pl, pf, pb)
        clocks = sampler.stop() if rank == 0 else None
        N.check(lib.pnr_tc_check(N.stream_ptr(device)), "pnr_tc_check")

        # ---- end-to-end: rays from pinned host memory, rgb+depth read back every step
        host_rays = [b.cpu().pin_memory() for b in mine[args.warmup:]]
        out_rgb = torch.empty(1, B, 3).pin_memory()
        out_d = torch.empty(1, B).pin_memory()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            r = host_rays[i].to(device, non_blocking=True)
            rgb, depth = render_par(r[None])
            out_rgb.copy_(rgb, non_blocking=True)
            out_d.copy_(depth, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
        e1.record()
        sync_all()
        ms_e2e = e0.elapsed_time(e1)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    total_rays = world * args.steps * B
    value = total_rays / (ms * 1e-3)
    e2e = total_rays / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
        roof = None
        if pl[1] > 0 and pms[1] > 0:
            ach = pf[1] / (pms[1] * 1e-3) / 1e12
            traffic = None
            try:  # DRAM bytes of one launch of this kernel from the committed ncu --set full capture
                prof = json.load(open(os.path.join(REPO, "profiles", "r01_v13_ncu_summary.json")))["mlp_phaseA_kernel"]
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                traffic = sum(float(prof[k][0]) * scale[prof[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            except Exception:
                pass
            roof = {"bound": "tensor", "kernel": "mlp_phaseA_kernel (fused gather + ResnetFC blocks 0..combine_layer-1 + view pool, tcgen05 cta_group::2)",
                    "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                    "traffic_note": "dram read+write bytes of ONE launch (~0.5 Mi points) from profiles/r01_v13_ncu_summary.json",
                    "peak_source": peak_src, "launches": int(pl[1]), "avg_launch_ms": pms[1] / pl[1],
                    "share_of_step": pms[1] / ms,
                    "other_kernels": {
                        "point_features_bf16_kernel": {"ms": pms[0], "launches": int(pl[0]),
                                                       "algorithmic_GBps": (pb[0] / (pms[0] * 1e-3) / 1e9) if pms[0] > 0 else None},
                        "mlp_phaseB_kernel": {"ms": pms[2], "launches": int(pl[2]),
                                              "TFLOPs": (pf[2] / (pms[2] * 1e-3) / 1e12) if pms[2] > 0 else None}}}
        cpu = None
        if not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            scene = oracle_scene(net, cam, conf)
            sample = 1024 if wl_name == "c1" else 512
            pick = torch.randperm(all_rays.shape[0], generator=torch.Generator().manual_seed(3))[:sample]
            torch.manual_seed(123)
            rate, sec = cpu_render_rate(scene, conf, all_rays[pick.to(device)].cpu(), 2, 1)
            cpu = {"value": rate, "unit": unit, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": "%d-ray sample of %s, 2 timed calls after 1 warm-up, oracle (torch %s CPU)" % (sample, wl_name, torch.__version__)}
        eager = None
        if not args.no_cpu_baseline:
            # honest software bar: the same oracle arithmetic as PyTorch eager ops on this
            # B200 (fp32, TF32 off), bounded sample; reported only, never the measured arm
            try:
                from oracle import pixelnerf_oracle as po

                torch.backends.cuda.matmul.allow_tf32 = False
                gscene = oracle_scene(net, cam, conf, device=device)
                sample = 8192
                rs = all_rays[torch.randperm(all_rays.shape[0], generator=torch.Generator().manual_seed(3))[:sample].to(device)]
                r = conf["renderer"]
                kw = dict(n_coarse=r.get_int("n_coarse", 128), n_fine=r.get_int("n_fine", 0),
                          n_fine_depth=r.get_int("n_fine_depth", 0), depth_std=r.get_float("depth_std", 0.01),
                          white_bkgd=bool(r.get_float("white_bkgd", False)), lindisp=False, eval_batch_size=200000)
                with torch.no_grad():
                    po.render(gscene, rs[None, :1024], **kw)
                    torch.cuda.synchronize(device)
                    t0 = time.perf_counter()
                    po.render(gscene, rs[None], **kw)
                    torch.cuda.synchronize(device)
                    dt = time.perf_counter() - t0
                eager = {"value": sample / dt, "unit": unit, "sample": "%d rays, oracle ops on cuda (torch %s eager, fp32)" % (sample, torch.__version__)}
                del gscene
            except Exception as ex:  # pragma: no cover
                eager = {"error": str(ex)[:200]}
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": dict(
                config, l2="per-step working set (operand scratch + samples, >1 GB) far exceeds the 126 MB L2; "
                           "consecutive steps render different ray batches"),
            "points_per_sec": value * 160, "clocks": clocks,
            "e2e": {"value": e2e, "unit": unit, "h2d_bytes_per_step": B * 8 * 4, "d2h_bytes_per_step": B * 4 * 4},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "torch_eager_gpu": eager}
        print(json.dumps(line), file=_json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
