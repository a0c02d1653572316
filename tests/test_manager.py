"""Row N3: the containment tests of the submap manager (Manager.py:159-244, geometry_helper.py:193-203).
CPU: the oracle restatement against vectors produced by the reference's OWN Manager methods (tests/golden/make_golden.py
`manager`).  GPU: mf_containment through mipsfusion_b200.manager against the same vectors -- integer counts, bit exact."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as H  # noqa: F401  (path set-up)
from oracle import manager as oman

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "manager.npz")


def _fixture():
    from mipsfusion_b200 import synth
    g = np.load(GOLD)
    dirs = synth.camera_rays()[::2, ::2].contiguous()
    t = {k: torch.from_numpy(g[k]) for k in ("depth", "c2w", "centers", "lens", "min_len", "pts_sub")}
    return g, dirs, t


def test_oracle_matches_reference_manager():
    g, dirs, t = _fixture()
    Hh, Ww = dirs.shape[0], dirs.shape[1]
    scores = oman.containing_scores(Hh, Ww, t["depth"], dirs, t["c2w"], t["centers"], t["lens"])
    assert np.array_equal(scores.numpy(), g["scores"])
    assert int(torch.argsort(scores, descending=True)[0]) == int(g["top_id"])
    lens = torch.where(t["lens"] < t["min_len"], t["min_len"], t["lens"])
    for key, (rh, rw) in (("ratios", (150, 200)), ("ratios_small", (30, 40))):
        r = [float(oman.compute_containing_ratio(Hh, Ww, t["depth"], dirs, t["c2w"], t["centers"][j], lens[j], rh, rw)) for j in range(len(lens))]
        np.testing.assert_array_equal(np.array(r, dtype=np.float64), g[key])
    m = oman.pts_in_bbox(t["pts_sub"], t["centers"] - 0.5 * t["lens"], t["centers"] + 0.5 * t["lens"])
    assert np.array_equal(m.numpy(), g["mask_sub"])


@pytest.mark.gpu
def test_gpu_containment_bit_exact_against_reference():
    import mipsfusion_b200 as mf
    g, dirs, t = _fixture()
    Hh, Ww = dirs.shape[0], dirs.shape[1]
    depth_d, dirs_d, c2w_d = t["depth"].cuda(), dirs.cuda(), t["c2w"].cuda()
    sc = mf.SubmapContainment(Hh, Ww, min_cr_localMLP_len=t["min_len"])
    scores = sc.containing_scores(depth_d, dirs_d, c2w_d, t["centers"], t["lens"])
    assert np.array_equal(scores.cpu().numpy(), g["scores"])                       # 90,000 (direction, depth) pairs x 6 boxes
    top = sc.find_highest_containing_ratio(depth_d, dirs_d, c2w_d, torch.arange(6), t["centers"], t["lens"])
    assert int(top) == int(g["top_id"])
    for key, (rh, rw) in (("ratios", (150, 200)), ("ratios_small", (30, 40))):
        r = [float(sc.compute_containing_ratio(depth_d, dirs_d, c2w_d, t["centers"][j], t["lens"][j], rh, rw)) for j in range(6)]
        np.testing.assert_array_equal(np.array(r, dtype=np.float64), g[key])       # integer counts, one fp32 division
    m = mf.pts_in_bbox(t["pts_sub"].cuda(), (t["centers"] - 0.5 * t["lens"]).cuda(), (t["centers"] + 0.5 * t["lens"]).cuda())
    assert np.array_equal(m.cpu().numpy(), g["mask_sub"])
    # edge cases: no points, a box nothing falls into, a point exactly on a face (strict inequalities)
    assert mf.pts_in_bbox(torch.zeros(0, 3).cuda(), torch.zeros(1, 3).cuda(), torch.ones(1, 3).cuda()).shape == (0, 1)
    p = torch.tensor([[0.5, 0.5, 0.5], [1.0, 0.5, 0.5], [0.0, 0.5, 0.5], [2.0, 2.0, 2.0]]).cuda()
    assert mf.pts_in_bbox(p, torch.zeros(1, 3).cuda(), torch.ones(1, 3).cuda())[:, 0].tolist() == [True, False, False, False]
