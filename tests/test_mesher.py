"""N2 (cont.): the Mesher's visibility filter -- oracle vs the reference's own methods (golden, CPU) and the CUDA kernels vs both
(GPU), bit for bit on the booleans."""
import numpy as np
import pytest
import torch

from oracle import mesher as om


def _inputs(g):
    W, H, fx, fy, cx, cy = g["cam"]
    K = torch.tensor([[fx, 0., cx], [0., fy, cy], [0., 0., 1.]], dtype=torch.float32)
    return int(W), int(H), K, torch.from_numpy(g["points"]), torch.from_numpy(g["poses"]), torch.from_numpy(g["rays_depth"]), torch.from_numpy(g["kf_ids"])


def test_oracle_matches_reference_mesher(golden):
    g = golden("mesher")
    W, H, K, pts, poses, depth, ids = _inputs(g)
    seen = om.point_mask(pts, depth[ids].max(1).values, poses, K, W, H)
    assert np.array_equal(seen.numpy(), g["seen"]) and 0.2 < g["seen"].mean() < 0.8
    assert np.array_equal(om.get_face_mask(g["seen"], g["faces"]), g["face_seen"])
    assert g["face_seen"].mean() < 0.9                                     # some faces are dropped


@pytest.mark.gpu
def test_gpu_visibility_bit_exact_against_reference(golden):
    import mipsfusion_b200 as mf
    g = golden("mesher")
    W, H, K, pts, poses, depth, ids = _inputs(g)
    vis = mf.MeshVisibility(K, W, H, depth[..., None])                      # rays (num_kf, n_rays, 1): last channel = depth
    seen = vis.point_mask(pts.cuda(), ids, poses)                           # poses on the CPU: same inverse as the fixture
    assert seen.dtype == torch.bool and np.array_equal(seen.cpu().numpy(), g["seen"])
    keep = vis.get_face_mask(seen, g["faces"])
    assert np.array_equal(keep.cpu().numpy(), g["face_seen"])
    # a larger cloud against the oracle, including points on a camera's z = 0 plane and no keyframes at all
    gen = torch.Generator().manual_seed(1)
    big = torch.tensor([1.0, 3.5, 1.0]) + 6.0 * (torch.rand(300000, 3, generator=gen) - 0.5)
    big[:100] = poses[0, :3, 3]                                              # camera centres: z = 0 exactly
    ref = om.point_mask(big, depth[ids].max(1).values, poses, K, W, H)
    out = vis.point_mask(big.cuda(), ids, poses)
    assert np.array_equal(out.cpu().numpy(), ref.numpy())
    none = vis.point_mask(big[:10].cuda(), ids[:0], poses[:0])
    assert not bool(none.any())
