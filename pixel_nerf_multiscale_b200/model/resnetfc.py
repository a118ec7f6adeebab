"""
ResnetFC: parameter container + native dispatch for the point MLP
(API/state-dict mirror of src/model/resnetfc.py; the arithmetic runs in csrc/mlp_f32.cu
[fp32 validation] or csrc/mlp_tc.cu [bf16 tcgen05]).

Natively supported: ReLU activations (beta == 0), combine_type == "average", no SPADE -- every
shipped conf.  Anything else raises NotImplementedError (there is no PyTorch/CPU fallback).
"""
import math

import torch
from torch import nn

from .. import _native as N


def _as_int(v, reduce_sum=False):
    if isinstance(v, (list, tuple)):
        if len(v) == 0:
            return 0
        return sum(int(x) for x in v) if reduce_sum else int(v[0])
    return int(v)


class ResnetBlockFC(nn.Module):
    """x + fc_1(relu(fc_0(relu(x)))) -- parameters only; see ResnetFC for execution."""

    def __init__(self, size_in, size_out=None, size_h=None, beta=0.0):
        super().__init__()
        size_out = size_in if size_out is None else size_out
        size_h = min(size_in, size_out) if size_h is None else size_h
        if size_in != size_out:
            raise NotImplementedError("ResnetBlockFC with a projection shortcut is not supported")
        self.size_in, self.size_h, self.size_out = size_in, size_h, size_out
        self.fc_0 = nn.Linear(size_in, size_h)
        self.fc_1 = nn.Linear(size_h, size_out)
        nn.init.zeros_(self.fc_0.bias)
        nn.init.kaiming_normal_(self.fc_0.weight, a=0, mode="fan_in")
        nn.init.zeros_(self.fc_1.bias)
        nn.init.zeros_(self.fc_1.weight)
        self.shortcut = None


class ResnetFC(nn.Module):
    def __init__(self, d_in, d_out=4, n_blocks=5, d_latent=0, d_hidden=128, beta=0.0, combine_layer=1000,
                 combine_type="average", use_spade=False):
        super().__init__()
        d_in, d_out = _as_int(d_in), _as_int(d_out)
        d_latent = _as_int(d_latent, reduce_sum=True)  # multi-scale encoders report a per-level list
        d_hidden, n_blocks, combine_layer = _as_int(d_hidden), _as_int(n_blocks), _as_int(combine_layer)
        if beta > 0:
            raise NotImplementedError("Softplus activations (beta > 0) are not supported by the native MLP")
        if use_spade:
            raise NotImplementedError("use_spade is not supported by the native MLP")
        if combine_type != "average":
            raise NotImplementedError("combine_type=%s is not supported by the native MLP" % combine_type)
        if d_in <= 0:
            raise NotImplementedError("d_in == 0 is not supported by the native MLP")
        self.d_in, self.d_out, self.d_latent, self.d_hidden = d_in, d_out, d_latent, d_hidden
        self.n_blocks, self.combine_layer, self.combine_type, self.use_spade = n_blocks, combine_layer, combine_type, False
        self.lin_in = nn.Linear(d_in, d_hidden)
        self.lin_out = nn.Linear(d_hidden, d_out)
        for lin in (self.lin_in, self.lin_out):
            nn.init.zeros_(lin.bias)
            nn.init.kaiming_normal_(lin.weight, a=0, mode="fan_in")
        self.blocks = nn.ModuleList([ResnetBlockFC(d_hidden, beta=beta) for _ in range(n_blocks)])
        if d_latent != 0:
            n_lin_z = min(combine_layer, n_blocks)
            self.lin_z = nn.ModuleList([nn.Linear(d_latent, d_hidden) for _ in range(n_lin_z)])
            for lin in self.lin_z:
                nn.init.zeros_(lin.bias)
                nn.init.kaiming_normal_(lin.weight, a=0, mode="fan_in")
        self.activation = nn.ReLU()
        self._native_cache = {}

    def __getstate__(self):
        # the native descriptors (ctypes structs with device pointers) are a cache: never copied/pickled
        state = self.__dict__.copy()
        state["_native_cache"] = {}
        state["_param_list"] = None
        return state

    # ---- native operand bookkeeping -------------------------------------------------------
    def _fingerprint(self):
        """Cheap identity of the current parameter values: storage pointer + in-place version counter of
        every parameter (torch bumps ``_version`` on copy_/load_state_dict/optimizer steps; ``.to()`` /
        ``_apply`` replace the storages).  The parameter list itself is cached (it only changes when a
        submodule is replaced, which ``_apply`` / ``load_state_dict`` below invalidate)."""
        ps = self.__dict__.get("_param_list")
        if ps is None:
            ps = self.__dict__["_param_list"] = list(self.parameters())
        return tuple([(p.data_ptr(), p._version) for p in ps])

    def _apply(self, fn, *a, **kw):
        self.__dict__["_param_list"] = None
        self._native_cache = {}
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self.__dict__["_param_list"] = None
        self._native_cache = {}
        return super().load_state_dict(*a, **kw)

    def native(self, precision):
        """ctypes Mlp descriptor for the current parameters (re-packed lazily after any
        in-place update / load_state_dict).  Returns (struct, keepalive)."""
        fp = self._fingerprint()
        hit = self._native_cache.get(precision)
        if hit is not None and hit[0] == fp:
            return hit[1], hit[2]
        keep = []

        def dev(t):
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            if not t.is_cuda:
                raise RuntimeError("pixelnerf_b200: the native MLP needs its parameters on a CUDA device")
            keep.append(t)
            return N.ptr(t)

        m = N.Mlp()
        m.d_in, m.d_latent, m.d_hidden, m.d_out = self.d_in, self.d_latent, self.d_hidden, self.d_out
        m.n_blocks, m.combine_layer = self.n_blocks, self.combine_layer
        m.n_lin_z = len(self.lin_z) if self.d_latent != 0 else 0
        m.combine_type = 0
        m.lin_in_w, m.lin_in_b = dev(self.lin_in.weight), dev(self.lin_in.bias)
        m.lin_out_w, m.lin_out_b = dev(self.lin_out.weight), dev(self.lin_out.bias)
        for i in range(m.n_lin_z):
            m.lin_z_w[i], m.lin_z_b[i] = dev(self.lin_z[i].weight), dev(self.lin_z[i].bias)
        for i, blk in enumerate(self.blocks):
            m.fc0_w[i], m.fc0_b[i] = dev(blk.fc_0.weight), dev(blk.fc_0.bias)
            m.fc1_w[i], m.fc1_b[i] = dev(blk.fc_1.weight), dev(blk.fc_1.bias)
        if precision != N.FP32:
            device = self.lin_in.weight.device
            nbytes = N.lib().pnr_mlp_packed_bytes(m)
            packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                N.check(N.lib().pnr_mlp_pack(m, N.ptr(packed), nbytes, precision, N.stream_ptr(device)), "pnr_mlp_pack")
            keep.append(packed)
            m.packed, m.packed_bytes, m.packed_dtype = N.ptr(packed), nbytes, precision
        self._native_cache[precision] = (fp, m, keep)
        return m, keep

    def forward(self, zx, combine_inner_dims=(1,), combine_index=None, dim_size=None, precision=None):
        """
        :param zx (..., d_latent + d_in) rows in reference order
        :param combine_inner_dims (NS, B): rows are reshaped to (-1, NS, B, .) and mean-pooled
               over NS at combine_layer
        :return (..., d_out) raw outputs (leading dims pooled over NS)
        """
        assert zx.size(-1) == self.d_latent + self.d_in, \
            "Input size %d != d_latent (%d) + d_in (%d)" % (zx.size(-1), self.d_latent, self.d_in)
        if not zx.is_cuda:
            raise RuntimeError("pixelnerf_b200: ResnetFC.forward needs CUDA tensors (no CPU path)")
        precision = N.BF16 if precision is None else precision
        rows = zx.reshape(-1, zx.size(-1)).float().contiguous()
        if len(combine_inner_dims) == 1 and combine_inner_dims[0] == 1:
            ns, p = 1, rows.shape[0]
        else:
            ns, p = int(combine_inner_dims[0]), int(math.prod(combine_inner_dims[1:]))
        if ns > 1 and self.combine_layer >= self.n_blocks:
            ns, p = 1, rows.shape[0]  # never pooled
        assert rows.shape[0] % (ns * p) == 0, "rows do not divide into combine_inner_dims"
        sb = rows.shape[0] // (ns * p)
        m, keep = self.native(precision)
        out = torch.empty(sb * p, self.d_out, dtype=torch.float32, device=zx.device)
        lib = N.lib()
        # same scope name as resnetfc.py:180
        with torch.autograd.profiler.record_function("resnetfc_infer"), torch.cuda.device(zx.device):
            nbytes = lib.pnr_mlp_forward_workspace(m, sb, ns, p, precision)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=zx.device)
            N.check(lib.pnr_mlp_forward(m, N.ptr(rows), sb, ns, p, precision, N.ptr(out), N.ptr(ws), nbytes,
                                        N.stream_ptr(zx.device)), "pnr_mlp_forward")
        if ns > 1 or zx.dim() == 2:
            return out if zx.dim() == 2 and ns == 1 else out.reshape(sb, p, self.d_out)
        return out.reshape(*zx.shape[:-1], self.d_out)

    @classmethod
    def from_conf(cls, conf, d_in, **kwargs):
        return cls(d_in, n_blocks=conf.get_int("n_blocks", 5), d_hidden=conf.get_int("d_hidden", 128),
                   beta=conf.get_float("beta", 0.0), combine_layer=conf.get_int("combine_layer", 1000),
                   combine_type=conf.get_string("combine_type", "average"),
                   use_spade=conf.get_bool("use_spade", False), **kwargs)
