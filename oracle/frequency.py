"""Oracle: tiny-cuda-nn 1.7 "Frequency" encoding.

TEST INFRASTRUCTURE ONLY.  **parity unpinned** (tinycudann==1.7 is un-vendored;
reference call sites model/encodings.py:29-39, model/scene_rep.py:37,123).
Published algorithm (include/tiny-cuda-nn/encodings/frequency.h):
  j = d*2K + 2k + s ;  out[j] = sin( scalbnf(x_d, k) * PI + s * PI/2 )
tcnn evaluates with the fast intrinsic __sinf; the oracle uses torch.sin on the
same fp32 argument, which therefore *is* the definition.
"""
import math
import torch

PI_F = torch.tensor(math.pi, dtype=torch.float32)          # (float)PI
HALF_PI_F = torch.tensor(math.pi / 2, dtype=torch.float32)


def frequency_args(x, n_frequencies=8):
    """fp32 sine arguments, (N, D*2K).  x: (N, D) fp32."""
    x = x.to(torch.float32)
    N, D = x.shape
    k = torch.arange(n_frequencies, dtype=torch.float32)
    xs = x[:, :, None] * torch.exp2(k)[None, None, :]            # scalbnf (exact)
    base = xs * PI_F                                             # one fp32 rounding
    arg = torch.stack([base, base + HALF_PI_F], -1)              # s = 0, 1
    return arg.reshape(N, D * n_frequencies * 2)


def frequency_encode(x, n_frequencies=8):
    return torch.sin(frequency_args(x, n_frequencies))
