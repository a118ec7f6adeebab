"""
NeRFRenderer: drop-in for src/render/nerf.py.  Same constructor / from_conf / sched_step /
bind_parallel surface, same mutable attributes (n_coarse, n_fine, using_fine, ...), same
outputs (nested rgb/depth[/weights] for coarse and fine, or the (rgb, depth) tuple with
simple_output).

Execution: one call of the C ABI's ``pnr_render_rays`` per ray batch when the model is this
package's PixelNeRFNet (no host loop over point chunks, no host sync); for any other model
object the per-ray kernels (sample_coarse / composite / sample_fine+sort) are used around
``model(points, coarse=..., viewdirs=...)`` calls, like the reference's composite().

Random numbers: the five draws of the reference (nerf.py:111,135-137,141,158) are made with
torch on the rays' device in the same order, shapes and dtype, so a seeded run consumes the
generator exactly as the reference does; a pre-drawn tape can be supplied for parity tests.
"""
import torch

from .. import _native as N
from ..model.models import PixelNeRFNet


class RenderOutput(dict):
    """Attribute-style nested dict (stands in for dotmap.DotMap in the reference's outputs)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value

    def toDict(self):
        return {k: (v.toDict() if isinstance(v, RenderOutput) else v) for k, v in self.items()}


class _RenderWrapper(torch.nn.Module):
    def __init__(self, net, renderer, simple_output):
        super().__init__()
        self.net = net
        self.renderer = renderer
        self.simple_output = simple_output

    def forward(self, rays, want_weights=False):
        if rays.shape[0] == 0:
            return torch.zeros(0, 3, device=rays.device), torch.zeros(0, device=rays.device)
        outputs = self.renderer(self.net, rays, want_weights=want_weights and not self.simple_output)
        if self.simple_output:
            lvl = outputs.fine if self.renderer.using_fine else outputs.coarse
            return lvl.rgb, lvl.depth
        return outputs.toDict()


class NeRFRenderer(torch.nn.Module):
    """
    :param n_coarse coarse (stratified) samples per ray
    :param n_fine fine samples per ray, INCLUDING n_fine_depth depth-guided ones
    :param noise_std training-time sigma noise (unused: inference path)
    :param depth_std std of the depth-guided samples
    :param eval_batch_size points per model call when the generic (non-fused) path is used
    :param white_bkgd white instead of black background
    :param lindisp sample linearly in disparity
    :param sched [[iters], [n_coarse], [n_fine]] sampling schedule
    """

    def __init__(self, n_coarse=128, n_fine=0, n_fine_depth=0, noise_std=0.0, depth_std=0.01,
                 eval_batch_size=100000, white_bkgd=False, lindisp=False, sched=None):
        super().__init__()
        self.n_coarse, self.n_fine, self.n_fine_depth = n_coarse, n_fine, n_fine_depth
        self.noise_std, self.depth_std = noise_std, depth_std
        self.eval_batch_size = eval_batch_size
        self.white_bkgd = white_bkgd
        self.lindisp = lindisp
        if lindisp:
            print("Using linear displacement rays")
        self.using_fine = n_fine > 0
        self.sched = sched if sched is not None and len(sched) > 0 else None
        self.register_buffer("iter_idx", torch.tensor(0, dtype=torch.long), persistent=True)
        self.register_buffer("last_sched", torch.tensor(0, dtype=torch.long), persistent=True)
        self.rng_tape = None  # optional dict(coarse=, u=, jitter=, normal=) consumed by the next forward

    # ---- random draws ------------------------------------------------------------------------
    def _draw(self, B, device):
        """Reference order: coarse jitter; then (after the coarse pass) u, jitter, normal."""
        tape = self.rng_tape
        self.rng_tape = None
        kc, kf, kd = int(self.n_coarse), int(self.n_fine), int(self.n_fine_depth)
        if tape is not None:
            # a supplied tape is materialised as contiguous fp32 on the rays' device (the locals keep the
            # copies alive across the launch) and must cover every ray: the kernels index it by ray
            want = {"coarse": (B, kc)}
            if self.using_fine and kf - kd > 0:
                want["u"] = want["jitter"] = (B, kf - kd)
            if self.using_fine and kd > 0:
                want["normal"] = (B, kd)
            out = {}
            for name, shape in want.items():
                t = tape.get(name)
                assert t is not None, "rng_tape is missing %r" % name
                assert tuple(t.shape) == shape, "rng_tape[%r] has shape %s, expected %s" % (name, tuple(t.shape), shape)
                out[name] = t.to(device=device, dtype=torch.float32).contiguous()
            return out

        t = {}
        t["coarse"] = torch.rand(B, kc, dtype=torch.float32, device=device)
        if self.using_fine:
            if kf - kd > 0:
                t["u"] = torch.rand(B, kf - kd, dtype=torch.float32, device=device)
                t["jitter"] = torch.rand(B, kf - kd, dtype=torch.float32, device=device)
            if kd > 0:
                t["normal"] = torch.randn(B, kd, dtype=torch.float32, device=device)
        return t

    # ---- fused path ----------------------------------------------------------------------------
    def _render_fused(self, net, rays, want_weights, taps=False):
        sb, b = rays.shape[0], rays.shape[1]
        device = rays.device
        R = sb * b
        flat = rays.reshape(-1, 8).float().contiguous()
        kc = int(self.n_coarse)
        kf = int(self.n_fine) if self.using_fine else 0
        kd = int(self.n_fine_depth) if self.using_fine else 0
        tape = self._draw(R, device)
        prec = net._native_precision()
        sc, keep_s = net.native_scene()
        mc, keep_c = net.native_mlp(True)
        mf = None
        if net.mlp_fine is not None and kf > 0:
            mf, keep_f = net.native_mlp(False)
        cfg = N.RenderCfg(kc, kf, kd, int(bool(self.white_bkgd)), int(bool(self.lindisp)), float(self.depth_std),
                          prec, int(want_weights))
        f32 = dict(dtype=torch.float32, device=device)
        res = {"coarse": {"rgb": torch.empty(R, 3, **f32), "depth": torch.empty(R, **f32)}}
        if want_weights:
            res["coarse"]["weights"] = torch.empty(R, kc, **f32)
        if taps:
            res["coarse"]["z"] = torch.empty(R, kc, **f32)
        if kf > 0:
            res["fine"] = {"rgb": torch.empty(R, 3, **f32), "depth": torch.empty(R, **f32)}
            if want_weights:
                res["fine"]["weights"] = torch.empty(R, kc + kf, **f32)
            if taps:
                res["fine"]["z"] = torch.empty(R, kc + kf, **f32)
        g = lambda lvl, k: N.ptr(res.get(lvl, {}).get(k))
        out = N.RenderOut(g("coarse", "rgb"), g("coarse", "depth"), g("coarse", "weights"), g("fine", "rgb"),
                          g("fine", "depth"), g("fine", "weights"), g("coarse", "z"), g("fine", "z"))
        ctape = N.RngTape(N.ptr(tape["coarse"]), N.ptr(tape.get("u")), N.ptr(tape.get("jitter")),
                          N.ptr(tape.get("normal")))
        lib = N.lib()
        with torch.cuda.device(device):
            nbytes = lib.pnr_render_workspace(sc, mc, mf, cfg, sb, b)
            ws = net.workspace(nbytes, device)
            N.check(lib.pnr_render_rays(sc, mc, mf, cfg, N.ptr(flat), sb, b, ctape, out, N.ptr(ws), ws.numel(),
                                        N.stream_ptr(device)), "pnr_render_rays")
        return res

    # ---- generic path (arbitrary model callable) -----------------------------------------------
    def _eval_model(self, model, flat_rays, z, coarse, sb):
        R, K = z.shape
        pts = flat_rays[:, None, :3] + z.unsqueeze(2) * flat_rays[:, None, 3:6]
        use_viewdirs = hasattr(model, "use_viewdirs") and model.use_viewdirs
        pts = pts.reshape(sb, -1, 3)
        chunk = (self.eval_batch_size - 1) // sb + 1
        dirs = flat_rays[:, None, 3:6].expand(-1, K, -1).reshape(sb, -1, 3) if use_viewdirs else None
        outs = []
        for s in range(0, pts.shape[1], chunk):
            if use_viewdirs:
                outs.append(model(pts[:, s:s + chunk], coarse=coarse, viewdirs=dirs[:, s:s + chunk]))
            else:
                outs.append(model(pts[:, s:s + chunk], coarse=coarse))
        return torch.cat(outs, dim=1).reshape(R, K, -1)[..., :4].float().contiguous()

    def _render_generic(self, model, rays, want_weights, taps=False):
        sb, b = rays.shape[0], rays.shape[1]
        device = rays.device
        R = sb * b
        flat = rays.reshape(-1, 8).float().contiguous()
        kc = int(self.n_coarse)
        kf = int(self.n_fine) if self.using_fine else 0
        kd = int(self.n_fine_depth) if self.using_fine else 0
        tape = self._draw(R, device)
        lib = N.lib()
        f32 = dict(dtype=torch.float32, device=device)

        def composite(z, out4, K):
            w = torch.empty(R, K, **f32)
            rgb = torch.empty(R, 3, **f32)
            depth = torch.empty(R, **f32)
            N.check(lib.pnr_composite(N.ptr(flat), N.ptr(z), N.ptr(out4), R, K, int(bool(self.white_bkgd)), N.ptr(w),
                                      N.ptr(rgb), N.ptr(depth), N.stream_ptr(device)), "pnr_composite")
            return w, rgb, depth

        with torch.cuda.device(device):
            z_c = torch.empty(R, kc, **f32)
            N.check(lib.pnr_sample_coarse(N.ptr(flat), N.ptr(tape["coarse"]), R, kc,
                                          int(bool(self.lindisp)), N.ptr(z_c), N.stream_ptr(device)),
                    "pnr_sample_coarse")
            w_c, rgb_c, d_c = composite(z_c, self._eval_model(model, flat, z_c, True, sb), kc)
            res = {"coarse": {"rgb": rgb_c, "depth": d_c}}
            if want_weights:
                res["coarse"]["weights"] = w_c
            if taps:
                res["coarse"]["z"] = z_c
            if kf > 0:
                z_f = torch.empty(R, kc + kf, **f32)
                N.check(lib.pnr_sample_fine_sorted(N.ptr(flat), N.ptr(z_c), N.ptr(w_c), N.ptr(d_c),
                                                   N.ptr(tape.get("u")), N.ptr(tape.get("jitter")),
                                                   N.ptr(tape.get("normal")), R, kc, kf, kd, float(self.depth_std),
                                                   int(bool(self.lindisp)), N.ptr(z_f), N.stream_ptr(device)),
                        "pnr_sample_fine_sorted")
                w_f, rgb_f, d_f = composite(z_f, self._eval_model(model, flat, z_f, False, sb), kc + kf)
                res["fine"] = {"rgb": rgb_f, "depth": d_f}
                if want_weights:
                    res["fine"]["weights"] = w_f
                if taps:
                    res["fine"]["z"] = z_f
        return res

    # ---- public API ----------------------------------------------------------------------------
    def forward(self, model, rays, want_weights=False, taps=False):
        """
        :param model PixelNeRFNet (fused path) or any callable model(xyz (SB,B,3), coarse=, viewdirs=)
        :param rays (SB, B, 8) [origin(3) dir(3) near far]
        :return RenderOutput: .coarse/.fine each with rgb (SB,B,3), depth (SB,B)[, weights (SB,B,K)]
        """
        if self.sched is not None and self.last_sched.item() > 0:
            self.n_coarse = self.sched[1][self.last_sched.item() - 1]
            self.n_fine = self.sched[2][self.last_sched.item() - 1]
        assert len(rays.shape) == 3
        if not rays.is_cuda:
            raise RuntimeError("pixelnerf_b200: NeRFRenderer needs CUDA rays (there is no CPU path)")
        sb = rays.shape[0]
        with torch.autograd.profiler.record_function("renderer_forward"):   # same scope name as nerf.py:264
            if isinstance(model, PixelNeRFNet):
                res = self._render_fused(model, rays, want_weights, taps)
            else:
                res = self._render_generic(model, rays, want_weights, taps)
        out = RenderOutput()
        for lvl, d in res.items():
            o = RenderOutput()
            o.rgb = d["rgb"].reshape(sb, -1, 3)
            o.depth = d["depth"].reshape(sb, -1)
            if "weights" in d:
                o.weights = d["weights"].reshape(sb, -1, d["weights"].shape[-1])
            if "z" in d:
                o.z = d["z"].reshape(sb, -1, d["z"].shape[-1])
            out[lvl] = o
        return out

    def sched_step(self, steps=1):
        """Advance the sampling schedule (called once per training iteration)."""
        if self.sched is None:
            return
        self.iter_idx += steps
        while self.last_sched.item() < len(self.sched[0]) and self.iter_idx.item() >= self.sched[0][self.last_sched.item()]:
            self.n_coarse = self.sched[1][self.last_sched.item()]
            self.n_fine = self.sched[2][self.last_sched.item()]
            print("INFO: NeRF sampling resolution changed on schedule ==> c", self.n_coarse, "f", self.n_fine)
            self.last_sched += 1

    @classmethod
    def from_conf(cls, conf, white_bkgd=False, lindisp=False, eval_batch_size=100000):
        return cls(conf.get_int("n_coarse", 128), conf.get_int("n_fine", 0),
                   n_fine_depth=conf.get_int("n_fine_depth", 0), noise_std=conf.get_float("noise_std", 0.0),
                   depth_std=conf.get_float("depth_std", 0.01), white_bkgd=conf.get_float("white_bkgd", white_bkgd),
                   lindisp=lindisp, eval_batch_size=conf.get_int("eval_batch_size", eval_batch_size),
                   sched=conf.get_list("sched", None))

    def bind_parallel(self, net, gpus=None, simple_output=False):
        """
        Module that renders rays with this renderer and the given network:
        forward(rays (SB,B,8), want_weights=False) -> (rgb, depth) | nested dict.
        With several ``gpus`` the ray batch is split along dim 1 across the devices
        (parallel.MultiDeviceRenderer); the packed source-view features and MLP operands are
        copied to each device once per encode()/weight update, only rays and outputs move per call.
        """
        wrapped = _RenderWrapper(net, self, simple_output=simple_output)
        if gpus is not None and len(gpus) > 1:
            from ..parallel import MultiDeviceRenderer

            print("Using multi-GPU", gpus)
            wrapped = MultiDeviceRenderer(wrapped, gpus)
        return wrapped
