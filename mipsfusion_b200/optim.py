"""Fused dense Adam (mf_adam_step): same update as ``torch.optim.Adam`` configured at reference
mipsfusion.py:580-584 / InactiveMap.py:53-57 (betas, per-group lr / eps / L2 weight_decay), one kernel
per parameter tensor, with the reference's ``zero_grad()`` optionally folded into the same pass."""
import ctypes as C

import torch

from . import _lib as L


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, zero_grad=False):
        """One Adam update.  zero_grad=True also clears every gradient in the same kernel pass
        (gradient tensors are kept, as with ``zero_grad(set_to_none=False)``)."""
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


def create_map_optimizer(model, lr_decoder, lr_embed):
    """FusedAdam with the parameter groups of reference mipsfusion.py:580-584.  Also lets the model's backward kernels add
    straight into ``.grad`` (scene_rep._grad_targets): this optimiser's loop is backward -> step -> zero_grad."""
    model.accumulate_grads_in_place = True
    return FusedAdam([{"params": list(model.decoder.parameters()), "weight_decay": 1e-6, "lr": lr_decoder},
                      {"params": list(model.embed_fn.parameters()), "eps": 1e-15, "lr": lr_embed}], betas=(0.9, 0.99))
