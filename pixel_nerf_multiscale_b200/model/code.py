"""
Positional encoding module (API/state-dict mirror of src/model/code.py).

On the rendering path the encoding is computed inside the fused point-feature kernel
(csrc/features.cuh: code_entry); this module exists so that checkpoints keep their
``code._freqs`` / ``code._phases`` buffers and so that callers can still evaluate the
encoding on a tensor directly.
"""
import math

import torch


class PositionalEncoding(torch.nn.Module):
    def __init__(self, num_freqs=6, d_in=3, freq_factor=math.pi, include_input=True):
        super().__init__()
        self.num_freqs = num_freqs
        self.d_in = d_in
        self.freq_factor = float(freq_factor)
        self.include_input = include_input
        self.freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)
        self.d_out = num_freqs * 2 * d_in + (d_in if include_input else 0)
        self.register_buffer("_freqs", torch.repeat_interleave(self.freqs, 2).view(1, -1, 1))
        phases = torch.zeros(2 * num_freqs)
        phases[1::2] = math.pi * 0.5
        self.register_buffer("_phases", phases.view(1, -1, 1))

    def forward(self, x):
        """(N, d_in) -> (N, d_out) = [x, sin(f0 x), cos(f0 x), sin(f1 x), ...]."""
        if x.numel() == 0:
            return torch.empty(x.shape[0], self.d_out, device=x.device, dtype=x.dtype)
        arg = x.unsqueeze(1) * self._freqs + self._phases  # (N, 2F, d_in)
        emb = torch.sin(arg).reshape(x.shape[0], -1)
        return torch.cat((x, emb), dim=-1) if self.include_input else emb

    @classmethod
    def from_conf(cls, conf, d_in=3):
        return cls(conf.get_int("num_freqs", 6), d_in, conf.get_float("freq_factor", math.pi),
                   conf.get_bool("include_input", True))
