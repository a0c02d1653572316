"""Joint multi-submap SDF / colour query with entropy- and distance-weighted blending: the query
+ blend part of reference model/Mesher.py:464-528 (geometry) and :606-663 (colour), fused into
CUDA passes over a grid generated on the fly (no host arrays, no per-chunk .cpu().numpy()).

The open3d bounding geometry and mesh clean-up stay where they are in the reference (out of scope);
``query`` hands back exactly the arrays they consume: the blended TSDF grid (-1 where no submap sees the
point), the validity mask, and per-submap containment masks.  ``extract_mesh`` goes on to the surface
on the device (marching cubes of ``marching_cubes.py``, then the blended vertex colours), so that the
grid never visits the host."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import dist as D


def get_grid_uniform(xyz_min, xyz_max, padding=0.05, voxel_size=0.05):
    """Per-axis coordinates of reference Mesher.get_grid_uniform (model/Mesher.py:43-55)."""
    res = [int(((xyz_max[k] + padding) - (xyz_min[k] - padding)) // voxel_size) for k in range(3)]
    return [np.linspace(xyz_min[k] - padding, xyz_max[k] + padding, res[k]) for k in range(3)]


class JointSubmapQuery:
    """Holds the per-submap descriptors (model, first-keyframe pose, AABB, centroid)."""

    def __init__(self, model_list, first_kf_poses, aabb_mins, aabb_maxs, centroids, device=None):
        self.models = list(model_list)
        self.M = len(self.models)
        self.device = torch.device(device) if device is not None else self.models[0]._device
        self.poses = [torch.as_tensor(p, dtype=torch.float32).cpu() for p in first_kf_poses]
        self.aabb_min = [np.asarray(a, dtype=np.float64) for a in aabb_mins]
        self.aabb_max = [np.asarray(a, dtype=np.float64) for a in aabb_maxs]
        self.centroids = [np.asarray(c, dtype=np.float32).reshape(3) for c in centroids]

    def _submaps(self):
        arr = (L.Submap * self.M)()
        keep = []
        for i, model in enumerate(self.models):
            f = model._field()
            keep.append(f._keepalive)
            arr[i].field = f
            w2l = torch.inverse(self.poses[i])[:3, :4].contiguous().reshape(-1).tolist()   # geometry_helper.py:94
            for k in range(12):
                arr[i].w2l[k] = w2l[k]
            for k in range(3):
                arr[i].aabb_min[k], arr[i].aabb_max[k] = float(self.aabb_min[i][k]), float(self.aabb_max[i][k])
                arr[i].centroid[k] = float(self.centroids[i][k])
        return arr, keep

    def _point_set(self, points, axes):
        ps = L.PointSet()
        keep = []
        if points is not None:
            p = torch.as_tensor(points, dtype=torch.float64).to(self.device).contiguous()
            keep.append(p)
            ps.pts, total = p.data_ptr(), p.shape[0]
        else:
            ax = [torch.as_tensor(a, dtype=torch.float64).to(self.device).contiguous() for a in axes]
            keep += ax
            ps.ax, ps.ay, ps.az = ax[0].data_ptr(), ax[1].data_ptr(), ax[2].data_ptr()
            ps.nx, ps.ny, ps.nz = ax[0].numel(), ax[1].numel(), ax[2].numel()
            total = ps.nx * ps.ny * ps.nz
        return ps, keep, total

    @torch.no_grad()
    def query(self, points=None, axes=None, vis=None, color=False, want_contain=False, group=None, shard="points"):
        """points (G,3) world coordinates or axes=[x,y,z] (np.linspace arrays).  vis: optional (G,M) bool.
        Multi-GPU (``group``): shard="points" splits the point index range across ranks (no data-path
        collective besides a max over M floats; results stay sharded, see ``range``); shard="submaps"
        gives rank r the submaps m = r mod G and all-reduces the partial sums.
        -> dict(sdf | rgb, mask, contain?, range=(g_begin, g_count))"""
        dev = self.device
        ps, keep_ps, total = self._point_set(points, axes)
        subs, keep_sub = self._submaps()
        ws, rk = D.world(group)
        g_begin, g_count = 0, total
        m_sel = list(range(self.M))
        if ws > 1 and shard == "points":
            g_begin, g_count, _ = D.shard_range(total, ws, rk)
        elif ws > 1:
            m_sel = D.round_robin(self.M, ws, rk)
        K = 4 if color else 2
        st = L.stream()
        max_dist = torch.zeros(self.M, device=dev, dtype=torch.float32)
        acc = torch.zeros(max(g_count, 1), K, device=dev, dtype=torch.float32)
        mask_any = torch.zeros(max(g_count, 1), device=dev, dtype=torch.uint8)
        contain = torch.zeros(max(g_count, 1), self.M, device=dev, dtype=torch.uint8) if want_contain else None
        vis_t = None
        if vis is not None:
            vis_t = torch.as_tensor(vis).to(dev).to(torch.uint8)[g_begin:g_begin + g_count].contiguous()
        scratch = torch.empty(int(L.lib().mf_joint_query_scratch_size(max(g_count, 1))), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            L.call("mf_joint_query_maxdist", C.byref(ps), subs, self.M, g_begin, g_count, L.ptr(max_dist), st)
            if ws > 1 and shard == "points":
                D.allreduce_max_(max_dist, group)
            # contiguous runs of selected submaps
            for m in m_sel:
                L.call("mf_joint_query_accumulate", C.byref(ps), subs, self.M, m, 1, L.ptr(max_dist), L.ptr(vis_t), int(color),
                       g_begin, g_count, L.ptr(acc), L.ptr(mask_any), L.ptr(contain), L.ptr(scratch), st)
            if ws > 1 and shard == "submaps":
                D.allreduce_sum_(acc, group)
                mask_any = D.allreduce_max_(mask_any.to(torch.int32), group).to(torch.uint8)
                if contain is not None:
                    contain = D.allreduce_max_(contain.to(torch.int32), group).to(torch.uint8)
            out = torch.empty(max(g_count, 1), K - 1, device=dev, dtype=torch.float32)
            L.call("mf_joint_query_finalize", L.ptr(acc), L.ptr(mask_any), int(color), g_count, L.ptr(out), st)
        res = {"mask": mask_any[:g_count].bool(), "range": (g_begin, g_count), "max_dist": max_dist}
        res["rgb" if color else "sdf"] = out[:g_count] if color else out[:g_count, 0]
        if contain is not None:
            res["contain"] = contain[:g_count].bool()
        return res

    @torch.no_grad()
    def extract_mesh(self, axes, isolevel=0.0, truncation=3.0, color=True, vis=None):
        """Steps 4-5 (+ the colour query of step 9) of reference Mesher.extract_mesh_jointly (model/Mesher.py:464-540,606-663)
        without leaving the device: blended SDF over the grid ``axes`` -> volume (nx, ny, nz) as Mesher.py:533 reshapes it ->
        marching cubes -> vertices in world coordinates (voxel index * spacing + origin, Mesher.py:535-543) -> blended colours.
        Grid points no submap contains are excluded from the surface the way the reference's ``mask=final_mask_mc`` does: they are
        handed to marching cubes as -inf, which its voxel test rejects (marching_cubes.cpp:83).
        The reference calls skimage's Lewiner marching cubes here (unpinned third-party package, absent offline); this route
        uses the marching cubes the reference ships in-tree (NumpyMarchingCubes, utils/utils.py:78), bit-identical to it.
        -> dict(vertices (V,3) float32 world, faces (F,3) int64, colors (V,3) float32 | None, sdf_volume (nx,ny,nz))"""
        from .marching_cubes import marching_cubes_device
        nx, ny, nz = (len(a) for a in axes)
        r = self.query(axes=axes, vis=vis)
        sdf = r["sdf"].masked_fill(~r["mask"], float("-inf"))
        vol = sdf.reshape(ny, nx, nz).transpose(0, 1).contiguous()
        verts, faces = marching_cubes_device(vol, isolevel, truncation)
        origin = torch.tensor([axes[0][0], axes[1][0], axes[2][0]], dtype=torch.float64, device=verts.device)
        spacing = torch.tensor([axes[k][2] - axes[k][1] for k in range(3)], dtype=torch.float64, device=verts.device)
        world = verts.to(torch.float64) * spacing + origin
        colors = None
        if color and world.shape[0] > 0:
            colors = self.query(points=world, color=True)["rgb"]
        return {"vertices": world.to(torch.float32), "faces": faces, "colors": colors, "sdf_volume": vol}
