"""Drop-in for reference model/encodings.py: ``get_encoder`` returning modules that look like
``tcnn.Encoding`` (flat ``params`` Parameter, ``n_output_dims``, fp32 output, picklable) but run on
the hand-written sm_100a kernels of this package.  No tiny-cuda-nn, no CPU fallback."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def grid_meta(log2_hashmap_size, n_levels=16, n_features=2, base_resolution=16, per_level_scale=2.0):
    meta = L.GridMeta()
    L.call("mf_hashgrid_meta", int(log2_hashmap_size), int(n_levels), int(n_features), int(base_resolution),
           float(per_level_scale), C.byref(meta))
    return meta


class _HashGridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, module):
        meta = module.meta
        N = x.shape[0]
        out = torch.empty(N, module.n_output_dims, device=x.device, dtype=torch.float32)
        L.call("mf_hashgrid_fwd", L.ptr(x), L.ptr(params), C.byref(meta), L.ptr(out), None, N, L.stream())
        ctx.module = module
        ctx.save_for_backward(x, params)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, params = ctx.saved_tensors
        meta = ctx.module.meta
        dy = dy.contiguous()
        g_params = torch.zeros_like(params) if ctx.needs_input_grad[1] else None
        g_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        if g_params is None:                      # the kernel always scatters; give it somewhere to go
            g_params = torch.zeros_like(params)
        L.call("mf_hashgrid_bwd", L.ptr(x), L.ptr(dy), L.ptr(params), C.byref(meta), L.ptr(g_params), L.ptr(g_x),
               x.shape[0], L.stream())
        return g_x, (g_params if ctx.needs_input_grad[1] else None), None


class HashGridEncoding(torch.nn.Module):
    """tcnn.Encoding(otype="HashGrid") replacement (reference model/encodings.py:14-25)."""

    def __init__(self, n_input_dims=3, n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16,
                 per_level_scale=2.0, seed=1337):
        super().__init__()
        if n_input_dims != 3:
            raise L.MipsFusionB200Error("HashGridEncoding: only 3-D inputs are supported")
        self.cfg = dict(n_levels=n_levels, n_features=n_features_per_level, log2_hashmap_size=log2_hashmap_size,
                        base_resolution=base_resolution, per_level_scale=float(per_level_scale))
        meta = self.meta
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_levels * n_features_per_level
        n_params = int(meta.offset[n_levels]) * n_features_per_level
        g = torch.Generator().manual_seed(seed)                 # tcnn default init: U(-1e-4, 1e-4)
        self.params = torch.nn.Parameter((torch.rand(n_params, generator=g, dtype=torch.float32) * 2 - 1) * 1e-4)

    @property
    def meta(self):                                             # not stored: keeps the module picklable
        m = self.__dict__.get("_meta")
        if m is None:
            c = self.cfg
            m = grid_meta(c["log2_hashmap_size"], c["n_levels"], c["n_features"], c["base_resolution"], c["per_level_scale"])
            self.__dict__["_meta"] = m
        return m

    def __getstate__(self):
        s = self.__dict__.copy()
        s.pop("_meta", None)
        return s


    def forward(self, x):
        x = L.f32c(x.reshape(-1, 3), self.params.device)
        return _HashGridFn.apply(x, self.params, self)

    def indices(self, x):
        """Debug / parity: (N, L, 8) int64 table indices of the 8 corners of every level."""
        x = L.f32c(x.reshape(-1, 3), self.params.device)
        N = x.shape[0]
        out = torch.empty(N, self.n_output_dims, device=x.device, dtype=torch.float32)
        idx = torch.empty(N, self.cfg["n_levels"], 8, device=x.device, dtype=torch.int32)
        L.call("mf_hashgrid_fwd", L.ptr(x), L.ptr(self.params.data), C.byref(self.meta), L.ptr(out), L.ptr(idx), N, L.stream())
        return idx.to(torch.int64) & 0xFFFFFFFF, out


class _FreqFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_freq):
        N, D = x.shape
        out = torch.empty(N, D * n_freq * 2, device=x.device, dtype=torch.float32)
        L.call("mf_freq_fwd", L.ptr(x), L.ptr(out), D, n_freq, N, L.stream())
        ctx.n_freq = n_freq
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        dy = dy.contiguous()
        L.call("mf_freq_bwd", L.ptr(x), L.ptr(dy), L.ptr(dx), x.shape[1], ctx.n_freq, x.shape[0], L.stream())
        return dx, None


class FrequencyEncoding(torch.nn.Module):
    """tcnn.Encoding(otype="Frequency") replacement (reference model/encodings.py:31-38)."""

    def __init__(self, n_input_dims=3, n_frequencies=12):
        super().__init__()
        self.n_input_dims, self.n_frequencies = n_input_dims, n_frequencies
        self.n_output_dims = n_input_dims * n_frequencies * 2
        self.params = torch.nn.Parameter(torch.zeros(0, dtype=torch.float32))      # tcnn keeps an empty params tensor

    def forward(self, x):
        x = L.f32c(x.reshape(-1, self.n_input_dims))
        if not x.is_cuda:
            raise L.MipsFusionB200Error("FrequencyEncoding needs CUDA tensors (no CPU fallback)")
        return _FreqFn.apply(x, self.n_frequencies)


class IdentityEncoding(torch.nn.Module):
    """tcnn.Encoding(otype="Identity") replacement (reference model/encodings.py:43-49)."""

    def __init__(self, n_input_dims=3):
        super().__init__()
        self.n_input_dims = self.n_output_dims = n_input_dims
        self.params = torch.nn.Parameter(torch.zeros(0, dtype=torch.float32))

    def forward(self, x):
        return x.to(torch.float32)


def get_encoder(encoding, input_dim=3, n_bins=16, n_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                desired_resolution=512):
    """Same signature and return value as reference model/encodings.py:6-52."""
    if "hash" in encoding.lower() or "tiled" in encoding.lower():
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (n_levels - 1))
        embed = HashGridEncoding(input_dim, n_levels, level_dim, log2_hashmap_size, base_resolution, per_level_scale)
    elif "freq" in encoding.lower():
        embed = FrequencyEncoding(input_dim, n_bins)
    elif "identity" in encoding.lower():
        embed = IdentityEncoding(input_dim)
    else:
        raise L.MipsFusionB200Error(f"unknown encoding {encoding!r}")
    return embed, embed.n_output_dims
