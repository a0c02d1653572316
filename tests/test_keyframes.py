"""Keyframe ray store ("next" row N1): oracle and CUDA path against vectors produced by the reference's own
KeyframeSet (tests/golden/keyframes.npz, random.sample draws recorded)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as H
from oracle import keyframes as okf
from oracle import sampling as osamp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keyframes.npz")
CASES = ["one", "two", "many", "five"]


def _draws(fx, name):
    d = [torch.from_numpy(fx[f"{name}:draw{i}"]) for i in range(int(fx[f"{name}:n_draws"]))]
    n_rel = len(fx[f"{name}:related"])
    # the reference draws first, then (more than two keyframes) last, then other (model/keyframeSet.py:393,403,419)
    if n_rel == 1:
        return dict(idx_first=d[0])
    if n_rel == 2:
        return dict(idx_first=d[0], idx_other=d[1])
    return dict(idx_first=d[0], idx_last=d[1], idx_other=d[2])


def test_oracle_store_and_sampling_match_reference():
    fx = np.load(GOLD)
    Hh, W, nh, nw = int(fx["H"]), int(fx["W"]), int(fx["nh"]), int(fx["nw"])
    rows, cols = osamp.sample_pixels_uniformly(Hh, W, nh, nw)
    for k in range(fx["store"].shape[0]):
        got = okf.store_keyframe(torch.from_numpy(fx["direction"][k]), torch.from_numpy(fx["rgb"][k]), torch.from_numpy(fx["depth"][k]), rows, cols, W)
        np.testing.assert_array_equal(got.numpy(), fx["store"][k])
    store = torch.from_numpy(fx["store"])
    for name in CASES:
        related = torch.from_numpy(fx[f"{name}:related"])
        dr = _draws(fx, name)
        counts = okf.split_counts(int(fx[f"{name}:pix"]), len(related))
        assert counts[0] == len(dr["idx_first"]) and counts[1] == len(dr.get("idx_other", [])) and counts[2] == len(dr.get("idx_last", []))
        rays, kf_ids, kf_indices = okf.sample_rays_in_submap(store, int(fx[f"{name}:first"]), related, int(fx[f"{name}:pix"]), **dr)
        np.testing.assert_array_equal(rays.numpy(), fx[f"{name}:rays"])
        np.testing.assert_array_equal(kf_ids.numpy(), fx[f"{name}:kf_ids"])
        np.testing.assert_array_equal(kf_indices.numpy(), fx[f"{name}:kf_indices"])
    rays, kf_ids, kf_indices = okf.sample_rays_in_given_kf(store, torch.from_numpy(fx["given:ids"]), torch.from_numpy(fx["given:draw"]))
    np.testing.assert_array_equal(rays.numpy(), fx["given:rays"])
    np.testing.assert_array_equal(kf_ids.numpy(), fx["given:kf_ids"])
    np.testing.assert_array_equal(kf_indices.numpy(), fx["given:kf_indices"])


@pytest.mark.gpu
def test_gpu_store_bit_exact_against_reference():
    import mipsfusion_b200 as mf
    fx = np.load(GOLD)
    Hh, W, nh, nw = int(fx["H"]), int(fx["W"]), int(fx["nh"]), int(fx["nw"])
    cfg = {"sampling": {"kf_n_rays_h": nh, "kf_n_rays_w": nw}}
    st = mf.KeyframeRayStore(cfg, Hh, W, fx["store"].shape[0], "cuda")
    for k in range(fx["store"].shape[0]):
        st.add_keyframe({"direction": torch.from_numpy(fx["direction"][k])[None], "rgb": torch.from_numpy(fx["rgb"][k])[None],
                         "depth": torch.from_numpy(fx["depth"][k])[None], "frame_id": 5 * k})
    np.testing.assert_array_equal(st.rays.cpu().numpy(), fx["store"])
    assert st.frame_ids == [5 * k for k in range(len(st))]
    for name in CASES:
        rays, kf_ids, kf_indices = st.sample_rays_in_submap(int(fx[f"{name}:first"]), fx[f"{name}:related"], int(fx[f"{name}:pix"]), **_draws(fx, name))
        np.testing.assert_array_equal(rays.cpu().numpy(), fx[f"{name}:rays"])
        np.testing.assert_array_equal(kf_ids.cpu().numpy(), fx[f"{name}:kf_ids"])
        np.testing.assert_array_equal(kf_indices.cpu().numpy(), fx[f"{name}:kf_indices"])
    rays, kf_ids, kf_indices = st.sample_rays_in_given_kf(fx["given:ids"], 25, idx=torch.from_numpy(fx["given:draw"]))
    np.testing.assert_array_equal(rays.cpu().numpy(), fx["given:rays"])
    np.testing.assert_array_equal(kf_ids.cpu().numpy(), fx["given:kf_ids"])
    np.testing.assert_array_equal(kf_indices.cpu().numpy(), fx["given:kf_indices"])


@pytest.mark.gpu
def test_gpu_device_draws_and_store_fed_mapping_step():
    """Draws made on the device are k distinct in-range indices in key order; a mapping step fed from the store equals the
    step on the same rays passed explicitly."""
    import mipsfusion_b200 as mf
    from mipsfusion_b200.mapper import FusedMapper
    n, k = 30000, 2600
    keys = torch.rand(n, generator=torch.Generator().manual_seed(3))
    idx = mf.sample_without_replacement(n, k, torch.device("cuda"), keys=keys.cuda()).cpu()
    assert idx.shape[0] == k and len(set(idx.tolist())) == k and int(idx.min()) >= 0 and int(idx.max()) < n
    order = torch.argsort(keys, descending=True, stable=True)[:k]
    np.testing.assert_array_equal(idx.numpy(), order.numpy())
    for edge_k in (0, 4096, n):
        assert mf.sample_without_replacement(n, edge_k, torch.device("cuda")).shape[0] == edge_k
    # store-fed step == explicit step (3 keyframes of one synthetic frame + current-frame rays)
    from mipsfusion_b200 import synth
    cfg = H.make_config(12, n_samples_d=32, n_range_d=11)
    cfg["training"]["perturb"] = 0
    cfg["sampling"] = {"kf_n_rays_h": 30, "kf_n_rays_w": 40}
    of = H.oracle_field(cfg, seed=2)
    dirs = synth.camera_rays()
    poses = synth.trajectory(4)[:3]
    st = mf.KeyframeRayStore(cfg, dirs.shape[0], dirs.shape[1], 4, "cuda")
    for kf, c2w in enumerate(poses):
        fr = synth.render_frame(c2w, dirs, seed=kf)
        st.add_keyframe({"direction": dirs, "rgb": fr["rgb"], "depth": fr["depth"], "frame_id": kf})
    g = torch.Generator().manual_seed(9)
    draws = dict(idx_first=torch.randperm(1200, generator=g)[:85], idx_last=torch.randperm(1200, generator=g)[:85],
                 idx_other=torch.randperm(1200, generator=g)[:86])
    cur = st.rays[2, :40].clone()
    poses_all = torch.stack(list(poses)).cuda()
    m1 = FusedMapper(H.cuda_model(cfg, H.state_of(of)))
    a = m1.step_from_store(st, 0, [0, 1, 2], poses_all, 256, cur_rays7=cur, **draws).cpu().numpy().copy()
    rays, _, kidx = st.sample_rays_in_submap(0, [0, 1, 2], 256, **draws)
    rays = torch.cat([rays, cur], 0).cpu(); kidx = torch.cat([kidx.cpu(), -torch.ones(40, dtype=torch.int64)])
    P = torch.stack(list(poses))
    rays_d = torch.sum(rays[:, None, :3] * P[kidx, :3, :3], -1)
    rays_o = P[kidx, :3, 3]
    m2 = FusedMapper(H.cuda_model(cfg, H.state_of(of)))
    b = m2.step(rays_o.cuda().contiguous(), rays_d.cuda().contiguous(), rays[:, 3:6].cuda().contiguous(), rays[:, 6].cuda().contiguous()).cpu().numpy().copy()
    np.testing.assert_array_equal(a, b)


def test_feistel_sampler_oracle_properties():
    """The device sampler's permutation (restated in numpy): k distinct in-range values, a full permutation at k = n."""
    for n, k in [(30000, 2600), (7, 7), (1, 1), (180000, 3000), (1200, 85), (4097, 4097)]:
        a = okf.feistel_sample(n, k, 4242 + n)
        assert len(set(a.tolist())) == k and a.min() >= 0 and a.max() < n
    assert sorted(okf.feistel_sample(3000, 3000, 1).tolist()) == list(range(3000))
    assert not np.array_equal(okf.feistel_sample(30000, 64, 1), okf.feistel_sample(30000, 64, 2))


@pytest.mark.gpu
def test_gpu_feistel_sampler_and_fused_sampling_bit_exact():
    import mipsfusion_b200 as mf
    dev = torch.device("cuda")
    for n, k, seed in [(30000, 2600, 5), (7, 7, 0), (1, 1, 9), (180000, 3000, 2 ** 31 + 5), (1200, 85, 77)]:
        got = mf.sample_without_replacement(n, k, dev, seed=seed).cpu().numpy()
        np.testing.assert_array_equal(got, okf.feistel_sample(n, k, seed & 0xffffffff))
    fx = np.load(GOLD)
    cfg = {"sampling": {"kf_n_rays_h": int(fx["nh"]), "kf_n_rays_w": int(fx["nw"])}}
    st = mf.KeyframeRayStore(cfg, int(fx["H"]), int(fx["W"]), fx["store"].shape[0], "cuda")
    st.rays.copy_(torch.from_numpy(fx["store"])); st.frame_ids = list(range(fx["store"].shape[0]))
    store = torch.from_numpy(fx["store"])
    for first, related, pix in [(2, [2], 20), (1, [1, 4], 30), (0, [0, 2, 3, 5], 37), (1, [1, 0, 2, 3, 4, 5], 40)]:
        rays, kf_ids, kf_indices, draws = st.sample_rays_in_submap(first, related, pix, seed=100 + pix, return_draws=True)
        nf, no, nl = st.split_counts(pix, len(related))
        nr = st.num_rays_to_save
        n_other_kf = len(related) - 2 if len(related) > 2 else len(related) - 1
        exp = dict(idx_first=torch.from_numpy(okf.feistel_sample(nr, nf, 100 + pix)))
        if no:
            exp["idx_other"] = torch.from_numpy(okf.feistel_sample(n_other_kf * nr, no, 101 + pix))
        if nl:
            exp["idx_last"] = torch.from_numpy(okf.feistel_sample(nr, nl, 102 + pix))
        e_rays, e_ids, e_idx = okf.sample_rays_in_submap(store, first, torch.tensor(related), pix, **exp)
        np.testing.assert_array_equal(draws.cpu().numpy(), torch.cat([exp["idx_first"], exp.get("idx_other", torch.empty(0, dtype=torch.int64)),
                                                                       exp.get("idx_last", torch.empty(0, dtype=torch.int64))]).numpy())
        np.testing.assert_array_equal(rays.cpu().numpy(), e_rays.numpy())
        np.testing.assert_array_equal(kf_ids.cpu().numpy(), e_ids.numpy())
        np.testing.assert_array_equal(kf_indices.cpu().numpy(), e_idx.numpy())


def test_split_counts_host_logic_matches_oracle_and_reference_formula():
    """first / other / last ray counts of sample_rays_in_submap (model/keyframeSet.py:392,402,409,413): host mirror == oracle,
    they add up to the request and follow the reference's max(n // k, n // 10 | n // 5) rule."""
    from mipsfusion_b200.keyframe_store import KeyframeRayStore
    for pix in (1, 9, 10, 37, 64, 2048, 2600, 4095):
        for k in (1, 2, 3, 4, 7, 12, 50):
            a = KeyframeRayStore.split_counts(pix, k)
            assert a == okf.split_counts(pix, k)
            assert sum(a) == pix and min(a) >= 0
            assert a[0] == max(pix // k, pix // 10)
            if k > 2:
                assert a[2] == max(pix // k, pix // 5)
