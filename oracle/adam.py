"""Oracle: the dense Adam update used for mapping.

TEST INFRASTRUCTURE ONLY.  torch.optim.Adam semantics (torch 1.13 / 2.x, amsgrad
off, maximize off) with the reference's hyper-parameters (mipsfusion.py:580-584:
betas (0.9, 0.99); decoder lr 1e-2, eps 1e-8, L2 weight_decay 1e-6; grid lr 1e-2,
eps 1e-15).  Restated explicitly so the fused kernel can be checked element-wise.
"""
import math
import torch


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.99, eps=1e-8, weight_decay=0.0):
    """In-place single-tensor Adam step, identical op order to torch._single_tensor_adam."""
    if weight_decay != 0:
        g = g.add(p, alpha=weight_decay)
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-step_size)
    return p, m, v


def make_optimizer(field, lr_decoder=1e-2, lr_embed=1e-2):
    """mipsfusion.py:580-584 on an OracleField."""
    return torch.optim.Adam([
        {"params": list(field.w.values()), "weight_decay": 1e-6, "lr": lr_decoder},
        {"params": [field.grid], "eps": 1e-15, "lr": lr_embed}], betas=(0.9, 0.99))
