"""Oracle self-consistency for the un-vendored tcnn encodings (parity unpinned):
torch int64 emulation vs the plain-C uint32 restatement, the SURVEY Appendix A
level table, and structural properties."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import hashgrid as hg
from oracle.frequency import frequency_encode

ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


@pytest.fixture(scope="module")
def clib():
    so = os.path.join(ORACLE_DIR, "libhashgrid_ref.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ctypes.CDLL(so)


def c_table(clib, T, L=16):
    sc = np.zeros(L, np.float32); rs = np.zeros(L, np.uint32); sz = np.zeros(L, np.uint32); of = np.zeros(L + 1, np.uint32)
    clib.hg_level_table(T, L, 16, ctypes.c_double(hg.per_level_scale_of()), sc.ctypes, rs.ctypes, sz.ctypes, of.ctypes)
    return sc, rs, sz, of


def test_level_table_appendix_a():
    t = hg.level_table(19)
    assert t["n_params"] == 9014144 and int(t["offset"][-1]) == 4507072
    assert list(t["resolution"]) == [16, 20, 24, 28, 34, 41, 49, 59, 71, 85, 102, 123, 148, 177, 213, 256]
    assert list(t["size"][:9]) == [4096, 8000, 13824, 21952, 39304, 68928, 117656, 205384, 357912]
    assert all(int(s) == 524288 for s in t["size"][9:])
    assert abs(float(t["scale"][15]) - 254.9998) < 1e-3 and float(t["scale"][15]) < 255.0
    t16 = hg.level_table(16)
    assert t16["n_params"] == 1616144 and list(t16["size"][4:6]) == [39304, 65536]


@pytest.mark.parametrize("T", [19, 16, 10])
def test_torch_vs_c(clib, T):
    sc, rs, sz, of = c_table(clib, T)
    t = hg.level_table(T)
    assert (sc == t["scale"]).all() and (rs == t["resolution"]).all() and (sz == t["size"]).all() and (of == t["offset"]).all()
    g = torch.Generator().manual_seed(T)
    x = (torch.rand(3000, 3, generator=g) * 1.8 - 0.4).float()
    x[0] = 0.0; x[1] = 1.0; x[2] = -1e-7; x[3] = 0.5; x[4] = -7.25; x[5] = 1.0 - 2 ** -24; x[6] = 33.0
    params = (torch.rand(t["n_params"], generator=g) * 2 - 1)
    idx, w, _ = hg.grid_corners(x, t)
    out = hg.hashgrid_encode(x, params, t)
    N, L = x.shape[0], 16
    cidx = np.zeros((N, L, 8), np.uint32); cw = np.zeros((N, L, 8), np.float32); cout = np.zeros((N, L * 2), np.float32)
    xn, pn = x.numpy().copy(), params.numpy().copy()
    clib.hg_eval(xn.ctypes, ctypes.c_int64(N), L, 2, sc.ctypes, rs.ctypes, sz.ctypes, of.ctypes, pn.ctypes,
                 cidx.ctypes, cw.ctypes, cout.ctypes)
    assert np.array_equal(cidx.astype(np.int64), idx.numpy())
    assert np.array_equal(cw, w.numpy())
    np.testing.assert_allclose(cout, out.numpy(), rtol=0, atol=1e-6)
    assert (idx.numpy() < t["size"][None, :, None]).all() and (idx.numpy() >= 0).all()


def test_weights_partition_of_unity_and_grad():
    t = hg.level_table(10)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(200, 3, generator=g).requires_grad_(True)
    _, w, _ = hg.grid_corners(x.detach(), t)
    np.testing.assert_allclose(w.sum(-1).numpy(), 1.0, atol=1e-6)
    params = torch.randn(t["n_params"], generator=g).requires_grad_(True)
    out = hg.hashgrid_encode(x, params, t)
    out.square().sum().backward()
    assert params.grad.abs().sum() > 0 and x.grad.abs().sum() > 0
    # finite-difference check of d/dx (kernel_grid_backward_input) on the coarsest level, away from cell faces
    xd = x.detach().clone()
    eps = 1e-3
    base = hg.hashgrid_encode(xd, params.detach(), t)[:, :2]
    xp = xd.clone(); xp[:, 0] += eps
    fd = (hg.hashgrid_encode(xp, params.detach(), t)[:, :2] - base) / eps
    x2 = xd.clone().requires_grad_(True)
    o2 = hg.hashgrid_encode(x2, params.detach(), t)[:, 0].sum()
    o2.backward()
    frac = (xd[:, 0] * 15 + 0.5) % 1.0
    keep = (frac < 0.95)
    np.testing.assert_allclose(x2.grad[keep, 0].numpy(), fd[keep, 0].numpy(), rtol=5e-2, atol=5e-2)


def test_frequency_layout():
    x = torch.tensor([[0.25, 0.5, 0.125]])
    y = frequency_encode(x, 8)
    assert y.shape == (1, 48)
    # j = d*16 + 2k + s : sin(2^k pi x_d + s pi/2)
    np.testing.assert_allclose(float(y[0, 0]), np.sin(np.pi * 0.25), rtol=1e-6)
    np.testing.assert_allclose(float(y[0, 1]), np.cos(np.pi * 0.25), rtol=1e-6)
    np.testing.assert_allclose(float(y[0, 16 + 2]), np.sin(2 * np.pi * 0.5), atol=1e-6)
    np.testing.assert_allclose(float(y[0, 32 + 4 + 1]), np.cos(4 * np.pi * 0.125), atol=1e-6)
