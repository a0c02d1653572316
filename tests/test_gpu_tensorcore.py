"""GPU parity of the tcgen05 (tensor-core, bf16x3 split) decoder against fp64, the fp32 CUDA-core decoder
and the oracle."""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def restore_impl():
    from mipsfusion_b200 import _lib as L
    yield
    L.call("mf_set_decoder_impl", 0)


@pytest.mark.parametrize("K", [16, 64, 96, 128])
def test_umma_single_layer(K):
    from mipsfusion_b200 import _lib as L
    g = torch.Generator().manual_seed(K)
    x = torch.randn(128, K, generator=g); w = torch.randn(128, K, generator=g)
    xc, wc = x.cuda(), w.cuda()
    ref = x.double() @ w.double().T
    for passes, tol in ((3, 2e-5), (1, 2e-2)):
        out = torch.zeros(128, 128, device="cuda")
        L.call("mf_debug_umma_linear", L.ptr(xc), L.ptr(wc), L.ptr(out), K, passes, L.stream())
        torch.cuda.synchronize()
        assert L.lib().mf_tc_check_error() == 0, "tcgen05 completion wait timed out"
        assert H.rel_err(out, ref) < tol, (K, passes)


def test_tensorcore_vs_cuda_core_vs_oracle():
    from mipsfusion_b200 import _lib as L
    cfg = H.make_config(16)
    of = H.oracle_field(cfg, grid_scale=0.4, seed=9)
    model = H.cuda_model(cfg, H.state_of(of), train=False)
    g = torch.Generator().manual_seed(1)
    for n in (1, 127, 128, 129, 5000):
        pts = torch.rand(n, 3, generator=g) * torch.tensor([3.5, 6.5, 4.2]) + torch.tensor([-0.6, 0.5, -1.15])
        with torch.no_grad():
            ref = of.run_network(pts)
            L.call("mf_set_decoder_impl", 0); tc = model.run_network(pts.cuda()).cpu()
            L.call("mf_set_decoder_impl", 1); cc = model.run_network(pts.cuda()).cpu()
        assert L.lib().mf_tc_check_error() == 0
        assert H.rel_err(cc, ref) < 1e-5, n
        assert H.rel_err(tc, ref) < 1e-4, n          # bf16x3: ~2^-16 per operand
        assert H.rel_err(tc, cc) < 1e-4, n
