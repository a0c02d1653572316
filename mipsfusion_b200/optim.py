"""Fused dense Adam (mf_adam_step): same update as ``torch.optim.Adam`` configured at reference
mipsfusion.py:580-584 / InactiveMap.py:53-57 (betas, per-group lr / eps / L2 weight_decay), one kernel
per parameter tensor, with the reference's ``zero_grad()`` optionally folded into the same pass.  The optimiser built by
``create_map_optimizer`` steps the hash grid and the (flat) decoder in ONE launch while the flat layout holds."""
import contextlib
import ctypes as C

import torch
import torch.optim.optimizer as _torch_opt

from . import _lib as L


_NULL_CTX = contextlib.nullcontext()


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _patch_step_function(self):
        """torch.optim.Optimizer wraps ``step`` in a profiler record_function + hook dispatcher (~10 us of host time per call,
        which the host-bound drop-in loop feels).  ``step`` below goes through that wrapper only when a step hook is registered."""
        self._zero_grad_profile_name = f"Optimizer.zero_grad#{self.__class__.__name__}.zero_grad"

    def step(self, closure=None, zero_grad=False):
        if self._optimizer_step_pre_hooks or self._optimizer_step_post_hooks or _torch_opt._global_optimizer_pre_hooks \
                or _torch_opt._global_optimizer_post_hooks:
            hooked = self.__dict__.get("_hooked_step")
            if hooked is None:
                hooked = self.__dict__["_hooked_step"] = torch.optim.Optimizer.profile_hook_step(FusedAdam._step)
            return hooked(self, closure, zero_grad=zero_grad)
        return self._step(closure, zero_grad=zero_grad)

    @torch.no_grad()
    def _step(self, closure=None, zero_grad=False):
        """One Adam update.  zero_grad=True also clears every gradient in the same kernel pass
        (gradient tensors are kept, as with ``zero_grad(set_to_none=False)``)."""
        if zero_grad and self.__dict__.get("_map_model") is not None and self._step_map_model():
            return None
        st = None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            ps = [p for p in group["params"] if p.grad is not None and p.numel() > 0]
            if not ps:
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise L.MipsFusionB200Error("FusedAdam needs CUDA parameters (no CPU fallback)")
            step = None
            for p in ps:
                state = self.state[p]
                if not state:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                state["step"] += 1
                if step is None:
                    step = state["step"]
                if state["step"] != step or p.device != dev or not (p.is_contiguous() and p.grad.is_contiguous()):
                    raise L.MipsFusionB200Error("FusedAdam: a parameter group must be contiguous, on one device and in step")
            n = len(ps)
            arr = C.c_void_p * n
            P = arr(*[p.data_ptr() for p in ps]); G = arr(*[p.grad.data_ptr() for p in ps])
            M = arr(*[self.state[p]["exp_avg"].data_ptr() for p in ps]); V = arr(*[self.state[p]["exp_avg_sq"].data_ptr() for p in ps])
            N = (C.c_int64 * n)(*[p.numel() for p in ps])
            with torch.cuda.device(dev):
                if st is None:
                    st = L.stream()
                L.call("mf_adam_step_multi", n, P, G, M, V, N, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                       float(group["weight_decay"]), int(step), 1 if zero_grad else 0, st)
            for p in ps:
                torch.autograd.graph.increment_version(p)      # the kernel wrote p behind autograd's back
        return None


    # ---- one-launch step for the optimiser create_map_optimizer builds ----------------------------------------------
    def _step_map_model(self):
        """Decoder (as one flat tensor) and hash grid in ONE kernel (mf_adam_step_pair, zero_grad folded in), then the
        kernel-layout weight image is refreshed.  Needs: the decoder's parameters and gradients are the views of their flat
        buffers (MLP_reg.flatten_parameters / flat_grad_target), dense fp32 grid gradient.  -> False: use the generic path."""
        model = self._map_model
        dec, grid = model.decoder, model.embed_fn.params
        flat = dec.flat_storage()
        if flat is None or len(self.param_groups) != 2:
            return False
        ps = dec.ordered_params()
        gflat = dec.__dict__.get("_flat_grad")
        gg = grid.grad
        if gflat is None or gg is None or not (grid.is_contiguous() and gg.is_contiguous()) or gg.dtype != torch.float32 \
                or grid.numel() % 4 or (grid.data_ptr() | gg.data_ptr()) % 16:
            return False
        o, base = 0, gflat.data_ptr()
        for p in ps:                                        # every decoder gradient is its view of the flat gradient buffer
            g = p.grad
            if g is None or g.data_ptr() != base + 4 * o:
                return False
            o += p.numel()
        g_dec, g_grid = self.param_groups
        if len(g_grid["params"]) != 1 or g_grid["params"][0] is not grid or len(g_dec["params"]) != len(ps) \
                or any(a is not b for a, b in zip(g_dec["params"], ps)):
            return False
        fs = self.__dict__.get("_flat_state")
        if fs is None or fs[0] is not flat:
            # moments of the decoder as flat buffers; the per-parameter state entries torch expects are views of them
            m, v = torch.zeros_like(flat), torch.zeros_like(flat)
            o = 0
            for p in ps:
                n, state = p.numel(), self.state[p]
                if state:
                    m[o:o + n].copy_(state["exp_avg"].reshape(-1)); v[o:o + n].copy_(state["exp_avg_sq"].reshape(-1))
                else:
                    state["step"] = 0
                state["exp_avg"], state["exp_avg_sq"] = m[o:o + n].view(p.shape), v[o:o + n].view(p.shape)
                o += n
            fs = (flat, m, v)
            self.__dict__["_flat_state"] = fs
        b1, b2 = g_grid["betas"]
        if tuple(g_dec["betas"]) != (b1, b2):
            return False
        sg = self.state[grid]
        if not sg:
            sg["step"] = 0
            sg["exp_avg"] = torch.zeros_like(grid, memory_format=torch.contiguous_format)
            sg["exp_avg_sq"] = torch.zeros_like(grid, memory_format=torch.contiguous_format)
        step = sg["step"] + 1
        if any(self.state[p]["step"] + 1 != step for p in ps):
            return False
        sg["step"] = step
        for p in ps:
            self.state[p]["step"] = step
        dev = grid.device
        with (_NULL_CTX if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)):
            L.call("mf_adam_step_pair", grid.data_ptr(), gg.data_ptr(), sg["exp_avg"].data_ptr(), sg["exp_avg_sq"].data_ptr(), grid.numel(),
                   float(g_grid["lr"]), float(g_grid["eps"]), float(g_grid["weight_decay"]),
                   flat.data_ptr(), gflat.data_ptr(), fs[1].data_ptr(), fs[2].data_ptr(), flat.numel(),
                   float(g_dec["lr"]), float(g_dec["eps"]), float(g_dec["weight_decay"]), float(b1), float(b2), int(step), L.stream())
            torch.autograd.graph.increment_version(grid)
            for p in ps:
                torch.autograd.graph.increment_version(p)      # the kernel wrote p behind autograd's back
            dec.prepared()                                      # rebuild the kernel-layout image now (same stream, same buffer)
        return True


def create_map_optimizer(model, lr_decoder, lr_embed):
    """FusedAdam with the parameter groups of reference mipsfusion.py:580-584.  Also lets the model's backward kernels add
    straight into ``.grad`` (scene_rep._grad_targets): this optimiser's loop is backward -> step -> zero_grad."""
    model.accumulate_grads_in_place = True
    if model.embed_fn.params.is_cuda and model.decoder.__dict__.get("_ext_prep") is None:
        model.decoder.flatten_parameters()              # (a FusedMapper already keeps the decoder flat in its own buffer)
    opt = FusedAdam([{"params": list(model.decoder.ordered_params()), "weight_decay": 1e-6, "lr": lr_decoder},
                     {"params": list(model.embed_fn.parameters()), "eps": 1e-15, "lr": lr_embed}], betas=(0.9, 0.99))
    opt.__dict__["_map_model"] = model                  # enables the one-launch step while the flat layout holds
    return opt
