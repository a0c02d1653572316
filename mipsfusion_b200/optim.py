"""Fused dense Adam (mf_adam_step): same update as ``torch.optim.Adam`` configured at reference
mipsfusion.py:580-584 / InactiveMap.py:53-57 (betas, per-group lr / eps / L2 weight_decay), one kernel
per parameter tensor, with the reference's ``zero_grad()`` optionally folded into the same pass."""
import torch

from . import _lib as L


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, zero_grad=False):
        """One Adam update.  zero_grad=True also clears every gradient in the same kernel pass
        (gradient tensors are kept, as with ``zero_grad(set_to_none=False)``)."""
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None or p.numel() == 0:
                    continue
                if not p.is_cuda:
                    raise L.MipsFusionB200Error("FusedAdam needs CUDA parameters (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad
                if not (p.is_contiguous() and g.is_contiguous()):
                    raise L.MipsFusionB200Error("FusedAdam needs contiguous parameters and gradients")
                with torch.cuda.device(p.device):
                    L.call("mf_adam_step", L.ptr(p), L.ptr(g), L.ptr(st["exp_avg"]), L.ptr(st["exp_avg_sq"]), p.numel(),
                           float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                           int(st["step"]), 1 if zero_grad else 0, L.stream())
                torch.autograd.graph.increment_version(p)      # the kernel wrote p behind autograd's back
        return None


def create_map_optimizer(model, lr_decoder, lr_embed):
    """FusedAdam with the parameter groups of reference mipsfusion.py:580-584."""
    return FusedAdam([{"params": list(model.decoder.parameters()), "weight_decay": 1e-6, "lr": lr_decoder},
                      {"params": list(model.embed_fn.parameters()), "eps": 1e-15, "lr": lr_embed}], betas=(0.9, 0.99))
