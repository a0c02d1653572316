"""Mesh visibility filter of the reference's Mesher on the device ("seen-mask"): ``point_mask``
(model/Mesher.py:247-281) and ``get_face_mask`` (:221-231).  The reference builds a (k, n, 3) tensor of the vertices in every
keyframe's camera frame and loops over the keyframes in Python; here one kernel visits each vertex once, walks the keyframes
and stops at the first one that sees it.  Arithmetic is fp32 in the reference's operation order, so the masks equal the
reference's bit for bit (``tests/golden/mesher.npz`` comes from the reference's own methods).  No CPU fallback."""
import torch

from . import _lib as L


class MeshVisibility:
    """``K`` (3,3) camera matrix, ``W`` / ``H`` image size (``config['cam']``), ``kf_rays``: the keyframe store's ``rays`` tensor
    (num_kf, n_rays, >=1) whose last channel is depth (``kfSet.rays``, model/keyframeSet.py) -- only each keyframe's maximum is read."""

    def __init__(self, K, W, H, kf_rays, edge=20):
        self.K = torch.as_tensor(K, dtype=torch.float32)
        self.W, self.H, self.edge = int(W), int(H), int(edge)
        self.kf_rays = kf_rays

    def point_mask(self, points, kf_Ids, kf_pose_c2w):
        """Mesher.point_mask: points (n,3) world, kf_Ids (k,), kf_pose_c2w (k,4,4) -> bool (n,) on the points' CUDA device."""
        dev = points.device if points.is_cuda else torch.device("cuda")
        pts = L.f32c(points.reshape(-1, 3), dev)
        ids = torch.as_tensor(kf_Ids).to(torch.int64).cpu()
        if ids.numel() == 0:                                                               # no keyframe: nothing is seen (Mesher.py:249)
            return torch.zeros(pts.shape[0], device=dev, dtype=torch.bool)
        w2c = torch.as_tensor(kf_pose_c2w).to(torch.float32).inverse()[:, :3, :4]          # Mesher.py:252 (on the poses' own device)
        w2c = L.f32c(w2c.reshape(-1, 12), dev)
        depth = self.kf_rays[ids.to(self.kf_rays.device)][..., -1]
        max_depth = L.f32c(depth.reshape(depth.shape[0], -1).max(dim=1).values, dev)          # Mesher.py:273
        seen = torch.empty(pts.shape[0], device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            L.call("mf_mesh_seen_mask", L.ptr(pts), pts.shape[0], L.ptr(w2c), L.ptr(max_depth), int(ids.numel()), L.ptr(L.f32c(self.K.reshape(9), dev)),
                   self.W, self.H, self.edge, L.ptr(seen), L.stream())
        return seen.bool()

    @staticmethod
    def get_face_mask(vert_mask, faces):
        """Mesher.get_face_mask: vert_mask (V,) bool, faces (F,3) -> bool (F,): False only where all three vertices are unseen."""
        vm = torch.as_tensor(vert_mask)
        dev = vm.device if vm.is_cuda else torch.device("cuda")
        vm = vm.to(dev).to(torch.uint8).contiguous()
        import numpy as np
        if not isinstance(faces, torch.Tensor):
            faces = torch.from_numpy(np.ascontiguousarray(np.asarray(faces).astype(np.int64)))
        f = faces.to(dev).to(torch.int64).contiguous()
        keep = torch.empty(f.shape[0], device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            L.call("mf_mesh_face_mask", L.ptr(vm), L.ptr(f), f.shape[0], L.ptr(keep), L.stream())
        return keep.bool()
