"""Overlap SDF difference of the inactive-map global BA: a composition of run_network on pose-dependent
points.  Vectors: the reference's own InactiveMap.get_SDF_dif / get_SDF_dif2 with two reference models
(tests/golden/overlap.npz).  The oracle composition is checked on the CPU, the same composition over the CUDA models on the
GPU (loss and the gradients w.r.t. the two first-keyframe poses, which flow through the points: fp32 backward route)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as H
from oracle import overlap as oov

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "overlap.npz")


def _state(fx, i):
    pre = f"w{i}:"
    return {k[len(pre):]: torch.from_numpy(v) for k, v in fx.items() if k.startswith(pre)}


def _run(models, fx, dev):
    t = lambda k: torch.from_numpy(fx[k]).to(dev)
    trunc = float(fx["trunc"])
    f1 = t("first1").requires_grad_(True); f2 = t("first2").requires_grad_(True)
    loss = oov.get_sdf_dif(models, t("rays"), t("ovlp"), 0, 1, f1, f2, trunc)
    loss.backward()
    a = (float(loss), f1.grad.cpu().numpy().copy(), f2.grad.cpu().numpy().copy())
    f1.grad = None; f2.grad = None
    rays = t("rays")
    loss2 = oov.get_sdf_dif2(models, rays[:, 6:7], rays[:, :3], t("mask2"), t("pose2"), 1, 0, f1, f2, trunc)
    loss2.backward()
    return a, (float(loss2), f1.grad.cpu().numpy().copy(), f2.grad.cpu().numpy().copy())


def _check(got, fx, tol, gtol=None, one_flip=None):
    """one_flip: allow at most ONE of the four pose-gradient matrices to miss gtol, and only with the signature of a single
    sample whose grid cell flipped: the error matrix is rank one (every sample adds an outer product row x [point, 1] to
    d loss / d pose) and stays below `one_flip`."""
    (l1, g11, g12), (l2, g21, g22) = got
    np.testing.assert_allclose(l1, float(fx["loss"]), rtol=tol)
    np.testing.assert_allclose(l2, float(fx["loss2"]), rtol=tol)
    flipped = []
    for g, k in ((g11, "g_first1"), (g12, "g_first2"), (g21, "g2_first1"), (g22, "g2_first2")):
        err = H.rel_err(g, fx[k])
        if err < (gtol or tol):
            continue
        assert one_flip is not None and err < one_flip, (k, err)
        sv = np.linalg.svd(np.asarray(g, np.float64) - np.asarray(fx[k], np.float64), compute_uv=False)
        assert sv[1] < 2e-2 * sv[0], (k, err, sv)           # rank one: one sample
        flipped.append(k)
    assert len(flipped) <= 1, flipped


def test_oracle_overlap_matches_reference():
    fx = np.load(GOLD)
    cfg = H.make_config(int(fx["hash_size"]))
    models = [H.oracle_field(cfg, _state(fx, i)) for i in range(2)]
    _check(_run(models, fx, "cpu"), fx, 2e-5)


@pytest.mark.gpu
def test_gpu_overlap_matches_reference():
    fx = np.load(GOLD)
    cfg = H.make_config(int(fx["hash_size"]))
    models = [H.cuda_model(cfg, _state(fx, i), train=False) for i in range(2)]
    # Measured on B200 (scripts/exp_overlap.py): both losses agree to 1e-7, three of the four 4x4 pose-gradient matrices to
    # 1e-5.  The fourth differs by 3.2e-2 (max-norm) with a RANK-ONE error pattern = one sample: torch's matrix inverse gives
    # a last-ulp different local pose on the GPU, one sample of this fixture sits on a cell boundary of a fine grid level, its
    # floor() flips, and the trilinear interpolant's gradient is discontinuous there (the value is continuous: the loss is
    # unaffected).  With 96 points x 2 models x 16 levels x 3 axes about one such event is expected.  Hence: losses at 1e-3,
    # gradients at 1e-4, and at most one matrix may differ -- by a rank-one matrix (= one sample) below 5e-2; _check asserts
    # exactly that structure.
    _check(_run(models, fx, "cuda"), fx, 1e-3, gtol=1e-4, one_flip=5e-2)
