"""Drop-in for reference model/decoder.py: ``MLP_reg`` with the same constructor, parameter names
(state_dict keys) and ``forward(embed, embed_pos, query_pts) -> (N, 10)``, evaluated by the fused
decoder kernels (mf_mlp_fwd / mf_mlp_bwd)."""
import torch
import torch.nn as nn

from . import _lib as L

PARAM_ORDER = ["pts_linear.0.weight", "pts_linear.0.bias", "pts_linear.2.weight", "pts_linear.2.bias",
               "rgb_linear.0.weight", "rgb_linear.0.bias", "sdf_linear.0.weight", "sdf_linear.0.bias",
               "sdf_linear.2.weight", "sdf_linear.2.bias"]


class _Workspace:
    """Per-device scratch for the per-CTA partial parameter gradients (never pickled)."""
    _bufs = {}
    _sizes = {}                                      # (points, ray grads, extra) -> floats (the size functions are pure)

    @classmethod
    def get(cls, device, extra_floats=0, field_points=None, want_ray_grads=False):
        """field_points: size the buffer for a fused field backward over that many points
        (mf_field_bwd_workspace_size), otherwise for the stand-alone decoder backward."""
        skey = (field_points, bool(want_ray_grads), extra_floats)
        n = cls._sizes.get(skey)
        if n is None:
            if field_points is not None:
                n = int(L.lib().mf_field_bwd_workspace_size(int(field_points), int(bool(want_ray_grads))))
            else:
                n = int(L.lib().mf_mlp_grad_workspace_size()) + int(extra_floats)
            if len(cls._sizes) < 256:
                cls._sizes[skey] = n
        # one buffer per (device, stream): backward passes issued on different streams must not share scratch
        key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, device=device, dtype=torch.float32)
            cls._bufs[key] = buf
        return buf


class _MLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, embed, embed_pos, pts, prep, *params):
        N = embed.shape[0]
        out = torch.empty(N, L.MF_RAW_DIM, device=embed.device, dtype=torch.float32)
        L.call("mf_mlp_fwd", L.ptr(embed), L.ptr(embed_pos), L.ptr(pts), L.ptr(prep), L.ptr(out), N, L.stream())
        ctx.save_for_backward(embed, embed_pos, pts, prep)
        ctx.shapes = [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, d_out):
        embed, embed_pos, pts, prep = ctx.saved_tensors
        N = embed.shape[0]
        dev = embed.device
        g_mlp = torch.zeros(L.MF_MLP_PARAMS, device=dev, dtype=torch.float32)
        d_embed = torch.empty_like(embed) if ctx.needs_input_grad[0] else None
        want_dx = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        d_pos = torch.empty_like(embed_pos) if want_dx else None
        d_pts = torch.empty_like(pts) if want_dx else None
        d_out = d_out.contiguous()
        ws = _Workspace.get(dev)
        L.call("mf_mlp_bwd", L.ptr(embed), L.ptr(embed_pos), L.ptr(pts), L.ptr(prep), L.ptr(d_out), L.ptr(g_mlp),
               L.ptr(d_embed), L.ptr(d_pos), L.ptr(d_pts), L.ptr(ws), N, L.stream())
        grads, o = [], 0
        for shp in ctx.shapes:
            n = shp.numel()
            grads.append(g_mlp[o:o + n].view(shp))
            o += n
        return (d_embed, d_pos if ctx.needs_input_grad[1] else None, d_pts if ctx.needs_input_grad[2] else None, None, *grads)


class MLP_reg(nn.Module):
    """MLP with SDF classification head -- reference model/decoder.py:6-75."""

    def __init__(self, cfg, input_ch=3, input_ch_pos=12, n_hidden=128, n_hidden_rgb=64, n_hidden_sdf=64,
                 n_hidden_branch=128, n_class=5, beta=80.):
        super().__init__()
        self.cfg = cfg
        self.input_ch = input_ch
        self.input_ch_pos = input_ch_pos + 3
        self.n_hidden, self.n_hidden_rgb, self.n_hidden_sdf, self.n_hidden_branch = n_hidden, n_hidden_rgb, n_hidden_sdf, n_hidden_branch
        self.n_class, self.max_class_Id, self.beta = n_class, n_class - 1, beta
        if (self.input_ch, self.input_ch_pos, n_hidden, n_hidden_rgb, n_hidden_sdf, n_hidden_branch, n_class) != (32, 51, 128, 64, 64, 128, 5):
            raise L.MipsFusionB200Error(
                "MLP_reg kernels are built for the reference instantiation (input_ch=32, input_ch_pos=48, widths "
                "128/64/64/128, 5 classes; model/scene_rep.py:45); got a different shape")
        self.pts_linear = nn.Sequential(nn.Linear(self.input_ch_pos, n_hidden), nn.ReLU(), nn.Linear(n_hidden, n_hidden_sdf + n_hidden_rgb))
        self.rgb_linear = nn.Sequential(nn.Linear(n_hidden_rgb + self.input_ch_pos, 3))
        self.sdf_linear = nn.Sequential(nn.Linear(n_hidden_sdf + self.input_ch, n_hidden_branch), nn.ReLU(),
                                        nn.Linear(n_hidden_branch, n_class), nn.Softmax(dim=-1))

    def ordered_params(self):
        """The ten state_dict tensors in order.  Cached (walking the nn.Sequential containers costs ~25 us per call); the
        cache is dropped when the first parameter object is no longer the cached one (copied module, re-assigned weights)."""
        cached = self.__dict__.get("_ordered")
        if cached is None or cached[0] is not self._modules["pts_linear"]._modules["0"]._parameters["weight"]:
            lins = (self.pts_linear[0], self.pts_linear[2], self.rgb_linear[0], self.sdf_linear[0], self.sdf_linear[2])
            cached = [t for lin in lins for t in (lin.weight, lin.bias)]
            self.__dict__["_ordered"] = cached
        return cached

    def flat_weights(self):
        """The state_dict tensors concatenated in order (MF_MLP_PARAMS floats)."""
        flat = self.flat_storage()
        if flat is not None:
            return flat[:L.MF_MLP_PARAMS]
        return torch.cat([p.detach().reshape(-1) for p in self.ordered_params()])

    def flatten_parameters(self):
        """Re-seat the ten parameters as views of ONE flat buffer in state_dict order (padded to a multiple of four floats:
        the layout of the kernels' weight blob), so that the kernel-layout weights are rebuilt without a concatenation and a
        fused optimiser steps the whole decoder in one launch (``create_map_optimizer``).  Values, shapes, ``state_dict``
        and ``load_state_dict`` are unchanged; ``.to()`` / ``deepcopy`` give ordinary separate tensors again (and the slower
        generic routes).  -> the flat buffer."""
        ps = self.ordered_params()
        dev = ps[0].device
        if dev.type != "cuda":
            raise L.MipsFusionB200Error("MLP_reg parameters must live on a CUDA device (no CPU fallback)")
        flat = torch.zeros((L.MF_MLP_PARAMS + 3) // 4 * 4, device=dev, dtype=torch.float32)
        o = 0
        with torch.no_grad():
            for p in ps:
                n = p.numel()
                flat[o:o + n].copy_(p.detach().reshape(-1))
                p.data = flat[o:o + n].view(p.shape)
                o += n
        self.__dict__["_flat"] = (flat, tuple(p.data_ptr() for p in ps))
        self.__dict__.pop("_prep_cache", None)
        return flat

    def flat_storage(self):
        """The flat buffer of :meth:`flatten_parameters` while every parameter still is its view, else None."""
        f = self.__dict__.get("_flat")
        if f is None:
            return None
        if tuple(p.data_ptr() for p in self.ordered_params()) != f[1]:
            self.__dict__.pop("_flat")
            return None
        return f[0]

    def prepared(self):
        """Kernel-layout weights, rebuilt only when a parameter changed (mf_mlp_prepare)."""
        ps = self.ordered_params()
        ext = self.__dict__.get("_ext_prep")
        if ext is not None:
            # a FusedMapper steps these parameters in place (they are views of its flat master copy) and refreshes its own
            # kernel-layout image after every step: hand that image out, so every route sees the stepped weights
            ptrs, owner = ext
            if tuple(p.data_ptr() for p in ps) == ptrs:
                vers = tuple(p._version for p in ps)
                if vers != owner._seen_versions:         # written by somebody else (load_state_dict, another optimiser)
                    L.call("mf_mlp_prepare", L.ptr(owner.mlp), L.ptr(owner.prep), L.stream())
                    owner._seen_versions = vers
                return owner.prep
            self.__dict__.pop("_ext_prep")
        flat = self.flat_storage()
        if flat is not None:
            # flat storage: the image is rebuilt IN PLACE (same buffer, so the cached field descriptor stays valid) whenever a
            # parameter's version moved; FusedAdam does it right after its step (refresh_prepared)
            key = tuple(p._version for p in ps)
            cache = self.__dict__.get("_prep_cache")
            if cache is None or cache[2] is not flat:
                cache = [None, torch.empty(int(L.lib().mf_mlp_prep_size()), device=flat.device, dtype=torch.float32), flat]
                self.__dict__["_prep_cache"] = cache
            if cache[0] != key:
                L.call("mf_mlp_prepare", L.ptr(flat), L.ptr(cache[1]), L.stream())
                cache[0] = key
            return cache[1]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        cache = self.__dict__.get("_prep_cache")
        if cache is None or len(cache) != 2 or cache[0] != key:
            dev = ps[0].device
            if dev.type != "cuda":
                raise L.MipsFusionB200Error("MLP_reg parameters must live on a CUDA device (no CPU fallback)")
            flat = self.flat_weights()
            prep = torch.empty(int(L.lib().mf_mlp_prep_size()), device=dev, dtype=torch.float32)
            L.call("mf_mlp_prepare", L.ptr(flat), L.ptr(prep), L.stream())
            cache = (key, prep)
            self.__dict__["_prep_cache"] = cache
        return cache[1]

    def __getstate__(self):
        s = self.__dict__.copy()
        s.pop("_prep_cache", None)
        s.pop("_ext_prep", None)
        s.pop("_flat_grad", None)
        s.pop("_flat", None)
        s.pop("_ordered", None)
        return s

    def flat_grad_target(self):
        """Flat (MF_MLP_PARAMS) gradient buffer the backward kernels can accumulate into directly, or None.

        The ten parameter ``.grad`` tensors are kept as views of one flat buffer (the layout of the kernels' gradient
        blob), so a backward pass adds into it in place instead of returning ten tensors that autograd then adds one by
        one.  Usable when every ``.grad`` is None (fresh / after ``zero_grad(set_to_none=True)``: the buffer is cleared
        and the views installed) or already is the matching view; any other state (grads assigned by the user, a copied
        module) returns None and the caller falls back to returning gradients."""
        ps = self.ordered_params()
        dev = ps[0].device
        flat = self.__dict__.get("_flat_grad")
        if flat is None or flat.device != dev:
            flat = torch.zeros((L.MF_MLP_PARAMS + 3) // 4 * 4, device=dev, dtype=torch.float32)     # padded: float4 optimiser kernels
            self.__dict__["_flat_grad"] = flat
        if all(p.grad is None for p in ps):
            flat.zero_()
            o = 0
            for p in ps:
                n = p.numel()
                p.grad = flat[o:o + n].view(p.shape)
                o += n
            return flat
        o, base = 0, flat.data_ptr()
        for p in ps:
            g = p.grad
            if g is None or g.data_ptr() != base + 4 * o or g.shape != p.shape or not g.is_contiguous() or g.dtype != torch.float32:
                return None
            o += p.numel()
        return flat


    def forward(self, embed, embed_pos, query_pts):
        dev = self.pts_linear[0].weight.device
        return _MLPFn.apply(L.f32c(embed, dev), L.f32c(embed_pos, dev), L.f32c(query_pts, dev), self.prepared(), *self.ordered_params())
