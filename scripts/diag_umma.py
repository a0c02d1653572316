"""Diagnostics for the tcgen05 data path (run on the GPU box): prints how the single-layer tensor-core
matmul deviates from fp64 for structured inputs, to localise layout / descriptor mistakes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mipsfusion_b200 import _lib as L


def run(x, w, passes):
    K = x.shape[1]
    xc, wc = x.cuda().contiguous(), w.cuda().contiguous()
    out = torch.full((128, 128), float("nan"), device="cuda")
    L.call("mf_debug_umma_linear", L.ptr(xc), L.ptr(wc), L.ptr(out), K, passes, L.stream())
    torch.cuda.synchronize()
    return out.cpu()


def main():
    g = torch.Generator().manual_seed(0)
    print("timeout flag before:", L.lib().mf_tc_check_error())
    for K in (16, 64, 96, 128):
        x = torch.randn(128, K, generator=g); w = torch.randn(128, K, generator=g)
        ref = (x.double() @ w.double().T)
        for passes in (1, 3):
            o = run(x, w, passes)
            err = (o.double() - ref).abs().max().item() / ref.abs().max().item()
            print(f"K={K:3d} passes={passes}: rel err {err:.3e}  nan={int(torch.isnan(o).sum())}  timeout={L.lib().mf_tc_check_error()}")
    # structured probes: identity-like inputs expose permutations
    K = 64
    x = torch.zeros(128, K); w = torch.zeros(128, K)
    for r in range(128):
        x[r, r % K] = 1.0 + r / 256.0
    for n in range(128):
        w[n, n % K] = 1.0
    o = run(x, w, 1)
    ref = x @ w.T
    bad = (o - ref).abs() > 1e-2
    print("probe: mismatches", int(bad.sum()), "of", bad.numel())
    if bad.any():
        idx = bad.nonzero()[:10]
        for r, c in idx.tolist():
            print(f"   out[{r},{c}] = {o[r, c]:.4f}  expected {ref[r, c]:.4f}")
        print("   row 0 nonzeros:", o[0].nonzero().flatten().tolist()[:16], " expected:", ref[0].nonzero().flatten().tolist()[:16])
        print("   row 9 nonzeros:", o[9].nonzero().flatten().tolist()[:16], " expected:", ref[9].nonzero().flatten().tolist()[:16])


if __name__ == "__main__":
    main()
