"""Containment tests of the submap manager on the device: the point-in-box part of reference
Manager.find_highest_containing_ratio / compute_containing_ratio (Manager.py:159-244) and ``pts_in_bbox``
(helper_functions/geometry_helper.py:193-203).  The reference back-projects a pixel lattice of the depth image to world
points with torch on the CPU and tests them against the submaps' axis-aligned boxes; here one kernel does both and returns
integer counts, so the decisions (which submap contains the frame best, containing ratio against the thresholds of
mipsfusion.py's submap switching) are bit-identical.  The surrounding bookkeeping (``kfSet.localMLP_info``, thresholds,
loop-closure state) stays in the reference's Manager; this class takes the boxes as tensors."""
import torch

from . import _lib as L
from .sampling_helper import sample_pixels_uniformly


def pts_in_bbox(pts, xyz_min, xyz_max):
    """geometry_helper.pts_in_bbox: pts (n,3), xyz_min / xyz_max (m,3) -> bool (n,m), strict inequalities."""
    dev = pts.device
    if dev.type != "cuda":
        raise L.MipsFusionB200Error("mipsfusion_b200 kernels need CUDA tensors (no CPU fallback)")
    p = L.f32c(pts.reshape(-1, 3), dev)
    lo, hi = L.f32c(xyz_min.reshape(-1, 3), dev), L.f32c(xyz_max.reshape(-1, 3), dev)
    n, m = p.shape[0], lo.shape[0]
    mask = torch.empty(n, m, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        L.call("mf_containment", None, None, None, L.ptr(p), L.ptr(lo), L.ptr(hi), m, n, 0, L.ptr(mask), None, L.stream())
    return mask.bool()


class SubmapContainment:
    """``H``, ``W``: image size the pixel lattice is laid over (``dataset.H``, ``dataset.W`` of the reference)."""

    def __init__(self, H, W, min_cr_localMLP_len=None):
        self.H, self.W = int(H), int(W)
        self.min_cr_localMLP_len = None if min_cr_localMLP_len is None else torch.as_tensor(min_cr_localMLP_len, dtype=torch.float32)

    def _counts(self, depth_img, rays_d, pose_world, xyz_min, xyz_max, rays_h, rays_w, cross=False):
        dev = depth_img.device
        if dev.type != "cuda":
            raise L.MipsFusionB200Error("mipsfusion_b200 kernels need CUDA tensors (no CPU fallback)")
        rows, cols = sample_pixels_uniformly(self.H, self.W, rays_h, rays_w)          # sampling_helper.py:28-46
        rows, cols = rows.to(dev), cols.to(dev)
        target_d = L.f32c(depth_img[rows, cols], dev)
        dirs = L.f32c(rays_d[rows, cols], dev)
        pose = L.f32c(torch.as_tensor(pose_world), dev).reshape(16)
        lo, hi = L.f32c(xyz_min.reshape(-1, 3), dev), L.f32c(xyz_max.reshape(-1, 3), dev)
        k = lo.shape[0]
        counts = torch.empty(2 * k + 1, device=dev, dtype=torch.int64)
        with torch.cuda.device(dev):
            L.call("mf_containment", L.ptr(dirs), L.ptr(target_d), L.ptr(pose), None, L.ptr(lo), L.ptr(hi), k, dirs.shape[0], int(cross),
                   None, L.ptr(counts), L.stream())
        return counts, k

    def containing_scores(self, depth_img, rays_d, pose_world, centers, lens, rays_h=15, rays_w=20):
        """Step 2 of Manager.find_highest_containing_ratio (:174-182): number of sampled points inside each of the k boxes
        ``centers +- 0.5 lens`` ((k,3) each).  As in the reference, the points are ALL pairs (ray direction i, depth j) of the
        lattice -- its expression ``rays_d[..., None, :] * target_d[..., :, None]`` broadcasts (P,1,3) * (P,1) to (P,P,3) -- so
        the scores count P * P points (90,000 for the default 15 x 20 lattice).  -> int64 (k,) on the device."""
        c, l = torch.as_tensor(centers, dtype=torch.float32), torch.as_tensor(lens, dtype=torch.float32)
        counts, k = self._counts(depth_img, rays_d, pose_world, c - 0.5 * l, c + 0.5 * l, rays_h, rays_w, cross=True)
        return counts[:k]

    def find_highest_containing_ratio(self, depth_img, rays_d, pose_world, localMLP_Ids, centers, lens, rays_h=15, rays_w=20):
        """Manager.find_highest_containing_ratio (:159-190): the id (element of ``localMLP_Ids``) of the box that contains the most
        lattice points; ties resolve as ``torch.argsort(score, descending=True)[0]`` does on the same device."""
        score = self.containing_scores(depth_img, rays_d, pose_world, centers, lens, rays_h, rays_w)
        top = torch.argsort(score, descending=True)
        return torch.as_tensor(localMLP_Ids).to(top.device)[top][0]

    def compute_containing_ratio(self, depth_img, rays_d, pose_world, center, length, rays_h=150, rays_w=200, clamp_len=True):
        """Manager.compute_containing_ratio (:199-244): share of the lattice's valid-depth surface points inside the box.
        ``clamp_len``: apply the reference's lower bound ``min_cr_localMLP_len`` to ``length`` (its default branch, :228-230).
        -> 0-dim float32 tensor on the device (integer counts divided once, as the reference)."""
        center, length = torch.as_tensor(center, dtype=torch.float32), torch.as_tensor(length, dtype=torch.float32)
        if clamp_len and self.min_cr_localMLP_len is not None:
            m = self.min_cr_localMLP_len.to(length)
            length = torch.where(length < m, m, length)
        counts, _ = self._counts(depth_img, rays_d, pose_world, (center - 0.5 * length)[None], (center + 0.5 * length)[None], rays_h, rays_w)
        return counts[1] / counts[2]
