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


def probes2():
    g = torch.Generator().manual_seed(1)
    for Kf in (64, 96, 128):
        dz = torch.randn(128, 128, generator=g); w = torch.randn(128, Kf, generator=g)
        dzc, wc = dz.cuda(), w.cuda()
        out = torch.full((128, Kf), float("nan"), device="cuda")
        L.call("mf_debug_umma_dgrad", L.ptr(dzc), L.ptr(wc), L.ptr(out), Kf, L.stream()); torch.cuda.synchronize()
        ref = dz.double() @ w.double()
        o = out.cpu()
        print(f"dgrad Kf={Kf:3d}: rel err {(o.double() - ref).abs().max().item() / ref.abs().max().item():.3e} nan={int(torch.isnan(o).sum())} timeout={L.lib().mf_tc_check_error()}")
        x = torch.randn(128, Kf, generator=g); xc = x.cuda()
        for passes in (1, 2):
            out = torch.full((128, Kf), float("nan"), device="cuda")
            L.call("mf_debug_umma_wgrad", L.ptr(dzc), L.ptr(xc), L.ptr(out), Kf, passes, L.stream()); torch.cuda.synchronize()
            ref = dz.double().T @ x.double()
            o = out.cpu()
            print(f"wgrad Kf={Kf:3d} passes={passes}: rel err {(o.double() - ref).abs().max().item() / ref.abs().max().item():.3e} nan={int(torch.isnan(o).sum())} timeout={L.lib().mf_tc_check_error()}")
    # structured probe for wgrad layout: dz = one-hot rows, x = one-hot rows
    Kf = 128
    dz = torch.zeros(128, 128); x = torch.zeros(128, Kf)
    for p in range(128):
        dz[p, (3 * p) % 128] = 1.0; x[p, (5 * p + 1) % Kf] = 1.0 + p / 256.0
    out = torch.zeros(128, Kf, device="cuda")
    dzc, xc = dz.cuda(), x.cuda()
    L.call("mf_debug_umma_wgrad", L.ptr(dzc), L.ptr(xc), L.ptr(out), Kf, 1, L.stream()); torch.cuda.synchronize()
    ref = dz.T @ x
    bad = (out.cpu() - ref).abs() > 1e-2
    print("wgrad probe mismatches:", int(bad.sum()))
    if bad.any():
        o = out.cpu()
        for r, c in bad.nonzero()[:8].tolist():
            print(f"   out[{r},{c}] = {o[r, c]:.4f} expected {ref[r, c]:.4f}")
    dz = torch.zeros(128, 128); w = torch.zeros(128, 128)
    for p in range(128):
        dz[p, (3 * p) % 128] = 1.0
    for n in range(128):
        w[n, (7 * n + 2) % 128] = 1.0 + n / 256.0
    out = torch.zeros(128, 128, device="cuda")
    dzc, wc = dz.cuda(), w.cuda()
    L.call("mf_debug_umma_dgrad", L.ptr(dzc), L.ptr(wc), L.ptr(out), 128, L.stream()); torch.cuda.synchronize()
    ref = dz @ w
    bad = (out.cpu() - ref).abs() > 1e-2
    print("dgrad probe mismatches:", int(bad.sum()))
    if bad.any():
        o = out.cpu()
        for r, c in bad.nonzero()[:8].tolist():
            print(f"   out[{r},{c}] = {o[r, c]:.4f} expected {ref[r, c]:.4f}")
        print("   row 1 nonzeros:", o[1].nonzero().flatten().tolist()[:8], "expected", ref[1].nonzero().flatten().tolist()[:8])


if __name__ == "__main__":
    main()
    probes2()
