"""GPU parity: hash-grid / frequency encodings and the decoder, through the drop-in modules (C-ABI)."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import hashgrid as hg
from oracle.decoder import mlp_reg
from oracle.frequency import frequency_encode

pytestmark = pytest.mark.gpu


def edge_points(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, 3, generator=g) * 1.8 - 0.4).float()
    x[0] = 0.0; x[1] = 1.0; x[2] = -1e-7; x[3] = 0.5; x[4] = -7.25; x[5] = 1.0 - 2 ** -24; x[6] = 33.0
    x[7] = torch.tensor([0.0, 1.0, 0.999999])
    return x


@pytest.mark.parametrize("T", [19, 16, 10])
def test_hashgrid_indices_bit_exact_and_values(T):
    import mipsfusion_b200 as mf
    enc, dim = mf.get_encoder("HashGrid", log2_hashmap_size=T, desired_resolution=256)
    assert dim == 32
    t = hg.level_table(T)
    g = torch.Generator().manual_seed(T)
    params = torch.rand(t["n_params"], generator=g) * 2 - 1
    enc.params.data.copy_(params)
    enc = enc.cuda()
    x = edge_points(20000, T)
    idx_o, _, _ = hg.grid_corners(x, t)
    idx, out = enc.indices(x.cuda())
    assert torch.equal(idx.cpu(), idx_o)                                   # integer work: bit exact
    out_o = hg.hashgrid_encode(x, params, t)
    np.testing.assert_allclose(out.cpu().numpy(), out_o.numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(enc(x.cuda()).detach().cpu().numpy(), out_o.numpy(), rtol=0, atol=2e-6)


def test_hashgrid_backward():
    import mipsfusion_b200 as mf
    T = 12
    enc, _ = mf.get_encoder("HashGrid", log2_hashmap_size=T, desired_resolution=256)
    t = hg.level_table(T)
    g = torch.Generator().manual_seed(3)
    params = torch.rand(t["n_params"], generator=g) * 2 - 1
    enc.params.data.copy_(params)
    enc = enc.cuda()
    x = torch.rand(3000, 3, generator=g)
    dy = torch.randn(3000, 32, generator=g)
    xo = x.clone().requires_grad_(True); po = params.clone().requires_grad_(True)
    hg.hashgrid_encode(xo, po, t).backward(dy)
    xc = x.cuda().requires_grad_(True)
    enc(xc).backward(dy.cuda())
    assert H.rel_err(enc.params.grad.cpu(), po.grad) < 1e-5
    assert H.rel_err(xc.grad.cpu(), xo.grad) < 1e-4
    # empty input
    assert enc(torch.zeros(0, 3, device="cuda")).shape == (0, 32)


def test_frequency_forward_backward():
    import mipsfusion_b200 as mf
    enc, dim = mf.get_encoder("Frequency", n_bins=8)
    assert dim == 48 and enc.params.numel() == 0
    g = torch.Generator().manual_seed(1)
    x = torch.rand(5000, 3, generator=g) * 1.4 - 0.2
    y_o = frequency_encode(x, 8)
    xc = x.cuda().requires_grad_(True)
    y = enc(xc)
    np.testing.assert_allclose(y.detach().cpu().numpy(), y_o.numpy(), rtol=0, atol=2e-6)
    dy = torch.randn(5000, 48, generator=g)
    xo = x.clone().requires_grad_(True)
    frequency_encode(xo, 8).backward(dy)
    y.backward(dy.cuda())
    assert H.rel_err(xc.grad.cpu(), xo.grad) < 1e-4


def test_decoder_golden_and_backward(golden):
    import mipsfusion_b200 as mf
    fx = golden("decoder")
    dec = mf.MLP_reg(None, input_ch=32, input_ch_pos=48)
    dec.load_state_dict({k[2:]: H.T(v) for k, v in fx.items() if k.startswith("w:")})
    dec = dec.cuda()
    e, ep, p = (H.T(fx[k]) for k in ("embed", "embed_pos", "pts"))
    ec, epc, pc = (t.cuda().requires_grad_(True) for t in (e, ep, p))
    out = dec(ec, epc, pc)
    # vector produced by the reference's own MLP_reg.forward
    assert H.rel_err(out.detach().cpu(), fx["out"]) < 1e-5
    g = torch.Generator().manual_seed(0)
    d_out = torch.randn(out.shape, generator=g)
    out.backward(d_out.cuda())
    w = {k[2:]: H.T(v).clone().requires_grad_(True) for k, v in fx.items() if k.startswith("w:")}
    eo, epo, po = (t.clone().requires_grad_(True) for t in (e, ep, p))
    mlp_reg(w, eo, epo, po).backward(d_out)
    for name, prm in dec.named_parameters():
        assert H.rel_err(prm.grad.cpu(), w[name].grad) < 1e-4, name
    assert H.rel_err(ec.grad.cpu(), eo.grad) < 1e-4
    assert H.rel_err(epc.grad.cpu(), epo.grad) < 1e-4
    assert H.rel_err(pc.grad.cpu(), po.grad) < 1e-4
