"""Drop-in for reference model/scene_rep.py: ``JointEncoding`` (one submap's neural field).

Same constructor, attributes (``embed_fn``, ``embedpos_fn``, ``decoder``, ``bounding_box``,
``coords_norm_factor``, ``config``), methods and returned dict keys as the reference
(model/scene_rep.py:11-238), so that ActiveMap / InactiveMap / RandomOptimizer / Mesher call it
unchanged.  ``forward`` / ``render_rays`` / ``run_network`` / ``query_*`` run as fused sm_100a
kernels: sample z -> (normalise + hash-grid + frequency encode + MLP) -> SDF-to-weight render ->
losses, with a hand-written backward to the grid, the decoder and the rays (pose gradients).

The reference draws the stratified jitter with CPU ``torch.rand`` (scene_rep.py:176); here it is
drawn on the device unless the caller passes ``u`` explicitly (the parity tests do).
"""
import copy
import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L
from .decoder import MLP_reg, _Workspace
from .encodings import get_encoder


def _linspace(a, b, n, device):
    # evaluated by torch on the CPU exactly as the reference does (scene_rep.py:158-166), then shipped
    return torch.linspace(a, b, steps=n).to(device=device, dtype=torch.float32).contiguous()


class _FieldQueryFn(torch.autograd.Function):
    """run_network / query_color_sdf: points (N,3) -> (N,10)."""

    @staticmethod
    def forward(ctx, pts, normalize, model, grid, *mlp_params):
        # gradients w.r.t. the points (pose path): fp32 decoder for the forward AND the backward.  Callers of this route
        # difference the outputs of two submaps (InactiveMap.get_SDF_dif): the residual is small by design, so the 1e-4 of
        # the tensor-core forward would be a percent-level error of the pose gradient (measured 3e-2 on
        # tests/golden/overlap.npz); the ray route (_RenderFn) keeps its tensor-core forward, its residuals are against targets
        ctx.impl = 1 if pts.requires_grad else 0
        field = model._field(impl=ctx.impl)
        N = pts.shape[0]
        out = torch.empty(N, L.MF_RAW_DIM, device=pts.device, dtype=torch.float32)
        L.call("mf_field_query", L.ptr(pts), C.byref(field), int(normalize), L.ptr(out), N, L.stream())
        ctx.model, ctx.normalize = model, normalize
        ctx.keep = field._keepalive
        ctx.save_for_backward(pts, grid)
        ctx.shapes = [p.shape for p in mlp_params]
        return out

    @staticmethod
    def backward(ctx, d_out):
        pts, grid = ctx.saved_tensors
        model = ctx.model
        field = model._field(ctx.keep, impl=ctx.impl)
        N = pts.shape[0]
        g_grid, ret_grid, g_mlp, ret_mlp = _grad_targets(model, grid, pts.device, ctx.needs_input_grad[3], all(ctx.needs_input_grad[4:]))
        d_pts = torch.empty_like(pts) if ctx.needs_input_grad[0] else None
        d_out = d_out.contiguous()                     # bound to a name: a temporary could be recycled before the launch
        ws = _Workspace.get(pts.device, field_points=N)
        L.call("mf_field_query_bwd", L.ptr(pts), C.byref(field), int(ctx.normalize), L.ptr(d_out), L.ptr(g_grid),
               L.ptr(g_mlp), L.ptr(d_pts), L.ptr(ws), N, L.stream())
        mlp_grads = _split(g_mlp, ctx.shapes) if ret_mlp else [None] * len(ctx.shapes)
        return (d_pts, None, None, g_grid if ret_grid else None, *mlp_grads)


def _split(flat, shapes):
    out, o = [], 0
    for shp in shapes:
        n = shp.numel()
        out.append(flat[o:o + n].view(shp))
        o += n
    return out


def _grad_targets(model, grid, dev, need_grid, need_mlp):
    """Where the backward kernels accumulate: straight into the parameters' ``.grad`` when that is possible (saves the
    zero-fill of a 36 MB temporary plus autograd's add into ``.grad``, and ten small adds for the decoder), else into fresh
    tensors that are returned to autograd.  OFF by default: plain autograd semantics (``torch.autograd.grad``, gradient
    hooks, ``backward(inputs=...)`` all see the parameter gradients).  ``mipsfusion_b200.create_map_optimizer`` (the fused
    stand-in for the reference's ``create_optimizer``, mipsfusion.py:580-584) switches it on for the model it is given:
    that loop is ``loss.backward(); optimizer.step(); zero_grad()``, for which the two are indistinguishable.  A parameter
    with registered hooks always gets its gradient returned.  -> (g_grid, return_grid, g_mlp, return_mlp)"""
    in_place = bool(getattr(model, "accumulate_grads_in_place", False))
    if in_place and (model.embed_fn.params._backward_hooks or any(q._backward_hooks for q in model.decoder.ordered_params())):
        in_place = False
    p = model.embed_fn.params
    g = p.grad if (in_place and need_grid and p.data_ptr() == grid.data_ptr()) else None
    if g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.shape == grid.shape and g.device == grid.device \
            and g.data_ptr() % 16 == 0:
        g_grid, ret_grid = g, False
    else:
        g_grid, ret_grid = torch.zeros_like(grid), True
    flat = model.decoder.flat_grad_target() if (in_place and need_mlp) else None
    if flat is not None:
        return g_grid, ret_grid, flat, False
    return g_grid, ret_grid, torch.zeros(L.MF_MLP_PARAMS, device=dev, dtype=torch.float32), True


def _row_view(t, dev, R, w):
    """fp32 device tensor of R rows x w columns with unit column stride (any row stride >= w): usable in place."""
    return (t.dtype == torch.float32 and t.device == dev and t.dim() == 2 and t.shape[0] == R and t.shape[1] == w and R > 0
            and (w == 1 or t.stride(1) == 1) and t.stride(0) >= w and t.data_ptr() % 4 == 0)


def _ld(t, w):
    """Row stride (floats) of a target tensor: contiguous (R,), (R,1) or (R,w), or a row-strided view accepted by _row_view."""
    return w if (t is None or t.dim() == 1) else int(t.stride(0))


_ZERO = {}                                      # device -> cached 0-dim zero (stand-in for an undefined loss gradient)


class _RenderFn(torch.autograd.Function):
    """render_rays (+ losses when targets are given).  Outputs:
    rgb (R,3), depth (R), aux (R,3)=[depth_var, disp, acc], z (R,S), raw (R,S,10), counts (2), then the scalars rgb_loss,
    depth_loss, sdf_loss, fs_loss, psnr (views of one 8-float buffer the loss kernel fills: separate outputs, so that the
    caller's weighted sum back-propagates four scalars instead of three select-backward nodes with a zero-fill each)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, target_rgb, target_d, u, emd_w, model, grid, *mlp_params):
        dev = rays_o.device
        R = rays_o.shape[0]
        cfg, lins = model._render_cfg(target_d is not None, emd_w, dev)
        S = cfg.n_samples_d + cfg.n_range_d
        # pose gradients (d loss / d rays) are ill-conditioned w.r.t. forward rounding (loss residuals cancel, the
        # frequency encoding multiplies by up to 2^7 pi): that route runs the fp32 decoder end to end
        ctx.impl = 1 if (rays_o.requires_grad or rays_d.requires_grad) else 0
        # ... its FORWARD can still use the tensor cores: measured pose gradients are as close to the oracle with the
        # tcgen05 forward (5e-7 .. 1e-6) as with the fp32 one -- the sensitivity is in the backward's recomputation
        fwd_impl = 0 if (ctx.impl == 1 and getattr(model, "pose_route_tc_forward", True)) else ctx.impl
        field = model._field(impl=fwd_impl)
        st = L.stream()
        z = torch.empty(R, S, device=dev, dtype=torch.float32)
        counts = torch.empty(2, device=dev, dtype=torch.int64)
        ld_d, ld_rgb = _ld(target_d, 1), _ld(target_rgb, 3)
        p_d = None if target_d is None else target_d.data_ptr()
        p_rgb = None if target_rgb is None else target_rgb.data_ptr()
        L.call("mf_sample_z_ld", p_d, ld_d, L.ptr(u), L.ptr(lins[0]), L.ptr(lins[1]), L.ptr(lins[2]), C.byref(cfg), L.ptr(z),
               L.ptr(counts), R, st)
        raw = torch.empty(R, S, L.MF_RAW_DIM, device=dev, dtype=torch.float32)
        # encoded-feature cache for the backward (tensor-core route, only when a backward can follow)
        feat = None
        if ctx.impl == 0 and fwd_impl == 0 and any(ctx.needs_input_grad):       # (grad mode is off inside Function.forward; this is the signal)
            feat = torch.empty(int(L.lib().mf_feat_cache_size(R * S)), device=dev, dtype=torch.uint8)
        ctx.feat = feat
        L.call("mf_field_query_rays", L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), C.byref(field), L.ptr(raw), L.ptr(feat), R, S, st)
        rgb = torch.empty(R, 3, device=dev, dtype=torch.float32)
        depth = torch.empty(R, device=dev, dtype=torch.float32)
        aux = torch.empty(R, 3, device=dev, dtype=torch.float32)
        # (the loss kernels write all eight entries when targets are given)
        losses = torch.empty(8, device=dev, dtype=torch.float32) if target_d is not None else torch.zeros(8, device=dev, dtype=torch.float32)
        scratch = torch.empty(R * 8, device=dev, dtype=torch.float32) if target_d is not None else None
        L.call("mf_render_loss_fwd_ld", L.ptr(raw), L.ptr(z), p_rgb, ld_rgb, p_d, ld_d, L.ptr(counts), C.byref(cfg),
               L.ptr(rgb), L.ptr(depth), L.ptr(aux), None, None, L.ptr(losses), L.ptr(scratch), R, S, st)
        ctx.model, ctx.cfg, ctx.S = model, cfg, S
        ctx.keep = field._keepalive
        ctx.has_t = target_d is not None
        ctx.save_for_backward(rays_o, rays_d, target_rgb, target_d, z, raw, counts, losses, grid)
        ctx.shapes = [p.shape for p in mlp_params]
        l_rgb, l_depth, l_sdf, l_fs, psnr = losses[0], losses[1], losses[2], losses[3], losses[4]
        ctx.mark_non_differentiable(aux, z, counts, psnr)
        ctx.set_materialize_grads(False)          # unused outputs arrive as None in backward (no dense zero buffers)
        return rgb, depth, aux, z, raw, counts, l_rgb, l_depth, l_sdf, l_fs, psnr

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_aux, g_z, g_raw, g_counts, g_l0, g_l1, g_l2, g_l3, g_psnr):
        rays_o, rays_d, target_rgb, target_d, z, raw, counts, losses, grid = ctx.saved_tensors
        model, cfg, S = ctx.model, ctx.cfg, ctx.S
        dev, R = rays_o.device, rays_o.shape[0]
        st = L.stream()
        field = model._field(ctx.keep, impl=ctx.impl)
        d_raw = torch.empty_like(raw)
        g_rgb = g_rgb.contiguous() if g_rgb is not None else None
        g_depth = g_depth.contiguous() if g_depth is not None else None
        gls = (g_l0, g_l1, g_l2, g_l3)
        if all(g is None or (g.dtype == torch.float32 and g.is_cuda and g.numel() == 1) for g in gls):
            # the four upstream loss gradients go to the kernel as they arrive (device scalars; None = 0): no packing kernel
            L.call("mf_render_loss_bwd_scalars", L.ptr(raw), L.ptr(z), None if target_rgb is None else target_rgb.data_ptr(), _ld(target_rgb, 3),
                   None if target_d is None else target_d.data_ptr(), _ld(target_d, 1), L.ptr(losses), C.byref(cfg),
                   *[None if g is None else g.data_ptr() for g in gls], L.ptr(g_rgb), L.ptr(g_depth), L.ptr(d_raw), R, S, st)
        else:
            zero = _ZERO.get(dev)
            if zero is None:
                zero = _ZERO[dev] = torch.zeros((), device=dev, dtype=torch.float32)
            gl = torch.stack([zero if g is None else g.reshape(()) for g in gls]).to(torch.float32)
            target_rgb = None if target_rgb is None else target_rgb.contiguous()
            target_d = None if target_d is None else target_d.contiguous()
            L.call("mf_render_loss_bwd", L.ptr(raw), L.ptr(z), L.ptr(target_rgb), L.ptr(target_d), L.ptr(counts), L.ptr(losses),
                   C.byref(cfg), L.ptr(gl), L.ptr(g_rgb), L.ptr(g_depth), L.ptr(d_raw), R, S, st)
        if g_raw is not None:
            d_raw = d_raw + g_raw
        want_rays = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        g_grid, ret_grid, g_mlp, ret_mlp = _grad_targets(model, grid, dev, ctx.needs_input_grad[7], all(ctx.needs_input_grad[8:]))
        d_o = torch.empty_like(rays_o) if want_rays else None
        d_d = torch.empty_like(rays_d) if want_rays else None
        ws = _Workspace.get(dev, field_points=R * S, want_ray_grads=want_rays)
        L.call("mf_field_query_rays_bwd", L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), C.byref(field), L.ptr(d_raw), L.ptr(ctx.feat),
               L.ptr(g_grid), L.ptr(g_mlp), L.ptr(d_o), L.ptr(d_d), L.ptr(ws), R, S, st)
        mlp_grads = _split(g_mlp, ctx.shapes) if ret_mlp else [None] * len(ctx.shapes)
        return (d_o, d_d, None, None, None, None, None, g_grid if ret_grid else None, *mlp_grads)


class JointEncoding(nn.Module):
    def __init__(self, config, bound_box, coords_norm_factor):
        super().__init__()
        self.config = config
        self.bounding_box = bound_box
        self.coords_norm_factor = coords_norm_factor
        self.get_resolution()
        self.get_encoding(config)
        self.get_decoder(config)
        self.save_initial_param()

    # ---- construction: model/scene_rep.py:24-55 -------------------------------------------------
    def get_resolution(self):
        dim_max = (self.bounding_box[:, 1] - self.bounding_box[:, 0]).max()
        if self.config["grid"]["voxel_sdf"] > 10:
            self.resolution_sdf = self.config["grid"]["voxel_sdf"]
        else:
            self.resolution_sdf = int(dim_max / self.config["grid"]["voxel_sdf"])

    def get_encoding(self, config):
        self.embedpos_fn, self.input_ch_pos = get_encoder(config["pos"]["enc"], n_bins=self.config["pos"]["n_bins"])
        self.embed_fn, self.input_ch = get_encoder(config["grid"]["enc"], log2_hashmap_size=config["grid"]["hash_size"],
                                                   desired_resolution=256)

    def get_decoder(self, config):
        self.decoder = MLP_reg(config, input_ch=self.input_ch, input_ch_pos=self.input_ch_pos)

    def save_initial_param(self):
        self.initial_dict = copy.deepcopy(self.state_dict())

    def recover_initial_param(self):
        self.load_state_dict(self.initial_dict)

    # ---- host-side descriptors -------------------------------------------------------------------
    def _norm_host(self):
        """(a, b) fp64 lists with x_n = (x - a) / b, cached per bound tensor (scene_rep.py:138-142)."""
        key = (id(self.bounding_box), id(self.coords_norm_factor), bool(self.config["grid"]["use_bound_normalize"]),
               bool(self.config["grid"]["tcnn_encoding"]))
        cache = self.__dict__.get("_norm_cache")
        if cache is None or cache[0] != key:
            if not self.config["grid"]["tcnn_encoding"]:
                a, b = [0.0] * 3, [1.0] * 3
            elif self.config["grid"]["use_bound_normalize"]:
                bb = torch.as_tensor(self.bounding_box).detach().to("cpu", torch.float64)
                a = bb[:, 0].tolist()
                b = (bb[:, 1] - bb[:, 0]).tolist()
            else:
                nf = torch.as_tensor(self.coords_norm_factor).detach().to("cpu", torch.float64)
                a = (-nf).tolist()
                b = (2 * nf).tolist()
            cache = (key, a, b)
            self.__dict__["_norm_cache"] = cache
        return cache[1], cache[2]

    def _field(self, keep=None, impl=0):
        """mf_field descriptor of the current weights (impl: 0 process default, 1 fp32 CUDA cores, 2 tcgen05)."""
        if self.config["pos"]["enc"].lower().find("freq") < 0 or self.config["pos"]["n_bins"] != 8:
            raise L.MipsFusionB200Error("fused field kernels are built for pos.enc=Frequency, n_bins=8")
        grid = self.embed_fn.params
        if not grid.is_cuda:
            raise L.MipsFusionB200Error("JointEncoding must live on a CUDA device (no CPU fallback); call .to('cuda')")
        prep = keep[1] if keep is not None else self.decoder.prepared()
        # the descriptor only depends on the two device pointers and the decoder selection: reuse it between calls
        key = (grid.data_ptr(), prep.data_ptr(), int(impl))
        cache = self.__dict__.setdefault("_field_cache", {})
        f = cache.get(key)
        if f is None:
            if len(cache) > 16:
                cache.clear()
            f = L.Field()
            f.grid, f.mlp_prep = grid.data_ptr(), prep.data_ptr()
            a, b = self._norm_host()
            for k in range(3):
                f.norm_a[k], f.norm_b[k] = a[k], b[k]
            f.norm_factor = float(self.config["training"]["norm_factor"])
            f.decoder_impl = int(impl)
            f.meta = self.embed_fn.meta
            cache[key] = f
        f._keepalive = (grid, prep)
        return f

    def _render_cfg(self, has_depth, emd_w, device):
        tr, cam = self.config["training"], self.config["cam"]
        cfg = L.RenderCfg()
        if has_depth:
            cfg.n_samples_d, cfg.n_range_d = int(tr["n_samples_d"]), int(tr["n_range_d"])
        else:
            cfg.n_samples_d, cfg.n_range_d = int(tr["n_samples"]), 0
        if not 0 < cfg.n_samples_d + cfg.n_range_d <= L.MF_MAX_SAMPLES:
            raise L.MipsFusionB200Error(f"render kernels hold a ray's samples in registers: 1..{L.MF_MAX_SAMPLES} samples per ray, "
                                        f"got {cfg.n_samples_d + cfg.n_range_d}")
        cfg.perturb = 1 if tr["perturb"] > 0.0 else 0
        cfg.rgb_missing_nz = 1 if tr["rgb_missing"] != 0 else 0
        cfg.trunc, cfg.sc_factor = float(tr["trunc"]), float(self.config["data"]["sc_factor"])
        cfg.depth_trunc, cfg.emd_w = float(cam["depth_trunc"]), float(emd_w)
        key = (has_depth, cfg.n_samples_d, cfg.n_range_d, float(cam["near"]), float(cam["far"]), float(tr["range_d"]), str(device))
        cache = self.__dict__.get("_lin_cache")
        if cache is None or cache[0] != key:
            lu = _linspace(cam["near"], cam["far"], cfg.n_samples_d, device) if cfg.n_samples_d > 0 else None
            lr = _linspace(-tr["range_d"], tr["range_d"], cfg.n_range_d, device) if cfg.n_range_d > 0 else None
            lf = _linspace(cam["near"], cam["far"], cfg.n_range_d, device) if cfg.n_range_d > 0 else None
            cache = (key, (lu, lr, lf))
            self.__dict__["_lin_cache"] = cache
        return cfg, cache[1]

    def __getstate__(self):
        s = self.__dict__.copy()
        for k in ("_norm_cache", "_lin_cache", "_field_cache"):
            s.pop(k, None)
        return s

    @property
    def _device(self):
        return self.embed_fn.params.device

    # ---- queries: model/scene_rep.py:106-146 -----------------------------------------------------
    def _query(self, pts, normalize):
        shape = pts.shape
        flat = L.f32c(pts.reshape(-1, shape[-1]), self._device)
        out = _FieldQueryFn.apply(flat, normalize, self, self.embed_fn.params, *self.decoder.ordered_params())
        return out.reshape(*shape[:-1], L.MF_RAW_DIM)

    def query_color_sdf(self, query_points):
        """(…,3) already-normalised points -> (N,10) rgb_raw + sdf + entropy + prob."""
        return self._query(query_points, False).reshape(-1, L.MF_RAW_DIM)

    def query_sdf(self, query_points):
        return self.query_color_sdf(query_points)[..., 3:4]

    def query_color(self, query_points):
        return torch.sigmoid(self.query_color_sdf(query_points)[..., :3])

    def query_sdf_entropy_prob(self, query_points):
        return self.query_color_sdf(query_points)[..., 3:]

    def run_network(self, inputs):
        """(…,3) points in the submap frame -> (…,10); normalisation fused (fp64, as the reference)."""
        return self._query(inputs, bool(self.config["grid"]["tcnn_encoding"]))

    # ---- rendering: model/scene_rep.py:58-103,153-187 ---------------------------------------------
    def _render(self, rays_o, rays_d, target_rgb, target_d, u, emd_w):
        dev = self._device
        rays_o, rays_d = L.f32c(rays_o, dev), L.f32c(rays_d, dev)
        R = rays_o.shape[0]
        # targets may stay row-strided views (the loop's column slices of its (R,10) batch tensor): the kernels take a row stride
        if target_d is not None:
            target_d = target_d if _row_view(target_d, dev, R, 1) else L.f32c(target_d, dev).reshape(R)
        if target_rgb is not None:
            target_rgb = target_rgb if _row_view(target_rgb, dev, R, 3) else L.f32c(target_rgb, dev)
        tr = self.config["training"]
        S = (tr["n_samples_d"] + tr["n_range_d"]) if target_d is not None else tr["n_samples"]
        if tr["perturb"] > 0.0:
            u = torch.rand(R, S, device=dev, dtype=torch.float32) if u is None else L.f32c(u, dev)
        else:
            u = None
        return _RenderFn.apply(rays_o, rays_d, target_rgb, target_d, u, emd_w, self, self.embed_fn.params,
                               *self.decoder.ordered_params())

    def render_rays(self, rays_o, rays_d, target_d=None, u=None):
        rgb, depth, aux, z, raw = self._render(rays_o, rays_d, None, target_d, u, 0.0)[:5]
        return {"rgb": rgb, "depth": depth, "disp_map": aux[:, 1], "acc_map": aux[:, 2], "depth_var": aux[:, 0],
                "z_vals": z, "raw": raw}

    def sdf2weights(self, sdf, z_vals, args=None):
        raw = torch.zeros(*sdf.shape, L.MF_RAW_DIM, device=self._device, dtype=torch.float32)
        raw[..., 3] = sdf
        return self._raw2outputs(raw, z_vals, args)[3]

    def raw2outputs(self, raw, z_vals):
        return self._raw2outputs(raw, z_vals, None)

    def _raw2outputs(self, raw, z_vals, args):
        """(no autograd) -> rgb_map, disp_map, acc_map, weights, depth_map, depth_var."""
        dev = self._device
        raw10 = raw
        if raw.shape[-1] != L.MF_RAW_DIM:                    # the reference passes (R,S,4+) slices
            raw10 = torch.zeros(*raw.shape[:-1], L.MF_RAW_DIM, device=dev, dtype=torch.float32)
            raw10[..., :raw.shape[-1]] = raw
        raw10, z = L.f32c(raw10.detach(), dev), L.f32c(z_vals.detach(), dev)
        R, S = z.shape
        cfg, _ = self._render_cfg(True, 0.0, dev)
        if args is not None:
            cfg.trunc, cfg.sc_factor = float(args["training"]["trunc"]), float(args["data"]["sc_factor"])
        rgb = torch.empty(R, 3, device=dev); depth = torch.empty(R, device=dev); aux = torch.empty(R, 3, device=dev)
        w = torch.zeros(R, S, device=dev)
        L.call("mf_render_loss_fwd", L.ptr(raw10), L.ptr(z), None, None, None, C.byref(cfg), L.ptr(rgb), L.ptr(depth), L.ptr(aux),
               L.ptr(w), None, None, None, R, S, L.stream())
        return rgb, aux[:, 1], aux[:, 2], w, depth, aux[:, 0]

    # ---- training forward: model/scene_rep.py:190-238 --------------------------------------------
    def forward(self, rays_o, rays_d, target_rgb, target_d, EMD_w=0.01, u=None):
        if not self.training:
            return self.render_rays(rays_o, rays_d, target_d=target_d, u=u)
        out = self._render(rays_o, rays_d, target_rgb, target_d, u, EMD_w)
        return {"rgb": out[0], "depth": out[1], "rgb_loss": out[6], "depth_loss": out[7], "sdf_loss": out[8],
                "fs_loss": out[9], "psnr": out[10]}
