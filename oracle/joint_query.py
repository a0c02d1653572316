"""Oracle: joint multi-submap SDF / colour query and weight blend (Mesher path).

TEST INFRASTRUCTURE ONLY.  Restates reference model/Mesher.py:464-528 (geometry)
and :606-663 (colour), vis/math_helper.py:47-96 (gaussian distance weight,
compute_weights) and helper_functions/geometry_helper.py:93-99 (world->local).
open3d's AxisAlignedBoundingBox.get_point_indices_within_bounding_box is the
inclusive test min <= p <= max on float64 coordinates.  The obbox / keyframe
visibility mask (open3d + kfSet, out of scope) is an explicit optional input.
"""
import numpy as np
import torch


def grid_points(origin, spacing, dims):
    """x-fastest?  No: reference get_grid_uniform uses np.meshgrid(x, y, z) (indexing 'xy')
    flattened -> index = (iy * nx + ix) * nz + iz  (Mesher.py:43-55)."""
    nx, ny, nz = dims
    x = origin[0] + spacing * np.arange(nx, dtype=np.float64)
    y = origin[1] + spacing * np.arange(ny, dtype=np.float64)
    z = origin[2] + spacing * np.arange(nz, dtype=np.float64)
    xx, yy, zz = np.meshgrid(x, y, z)
    return np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T.astype(np.float32)


def pdf_gauss(x, mu=0.0, sigma=1.0):                               # vis/math_helper.py:47-51
    k1 = 1 / (sigma * np.sqrt(2 * np.pi))
    m1 = (x - mu) / sigma
    return k1 * np.exp(-0.5 * m1 ** 2)


def compute_weights(entropy, dist_w, mask):                        # vis/math_helper.py:79-96
    mask2 = (mask.astype(np.float32).sum(-1) != 0)[..., None]
    ent_inv = 1.0 * np.exp(-10.0 * entropy)
    vals = (ent_inv * mask.astype(np.float32)) * (dist_w * mask.astype(np.float32))
    norms = vals.sum(-1, keepdims=True)
    ok = mask2 & (norms > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(ok, vals / norms, np.zeros_like(vals))


def joint_query(points, fields, first_kf_poses, aabb_min, aabb_max, centroids, vis_masks=None, color=False):
    """points (G,3) world, float64 as the reference builds them (np.meshgrid of linspace axes, Mesher.py:43-55) or float32;
    containment is tested on the coordinates as given (open3d works on float64, Mesher.py:470), the network sees them
    rounded to float32 (Mesher.py:477).  per submap i: fields[i] (OracleField), first_kf_poses[i] (4,4) fp32
    c2w of the submap frame, aabb (3,) fp64, centroid (3,) fp32.  Returns dict with blended sdf
    (or rgb), per-submap containment masks (the integer 'submap assignment') and weights."""
    G, M = points.shape[0], len(fields)
    contain = np.zeros((G, M), dtype=bool)
    mask = np.zeros((G, M), dtype=bool)
    entropy = np.zeros((G, M), np.float32)
    tsdf = np.full((G, M), -1, np.float32)
    rgb = np.zeros((G, M, 3), np.float32)
    dist_w = np.zeros((G, M), np.float32)
    p64 = np.asarray(points, dtype=np.float64)
    points = np.asarray(points).astype(np.float32)
    for i in range(M):
        c1 = np.all((p64 >= aabb_min[i]) & (p64 <= aabb_max[i]), axis=-1)           # Mesher.py:470
        contain[:, i] = c1
        idx = np.where(c1)[0]
        if idx.size == 0:
            continue
        vp = torch.from_numpy(points[c1])
        w2c = torch.inverse(first_kf_poses[i])                                       # geometry_helper.py:93-99
        local = torch.sum(vp[:, None, :] * w2c[None, :3, :3], -1) + w2c[None, :3, 3]
        with torch.no_grad():
            out = fields[i].query_color_sdf(fields[i].normalize(local))             # Mesher.py:480-489
        out = out.numpy()
        tsdf[idx, i] = out[:, 3]
        entropy[idx, i] = out[:, 4]
        rgb[idx, i] = 1.0 / (1.0 + np.exp(-out[:, :3].astype(np.float32)))
        d = np.linalg.norm(points[c1] - centroids[i][None], axis=-1)                 # math_helper.py:58-60
        sigma = np.abs(d).max() / 3.0                                                # :66-72
        dist_w[idx, i] = pdf_gauss(np.abs(d), 0.0, sigma)
        mask[idx, i] = True if vis_masks is None else vis_masks[idx, i]              # Mesher.py:509-512
    entropy = np.clip(entropy, 0, 10000.0)                                           # :522
    W = compute_weights(entropy, dist_w, mask)
    any_mask = mask.any(-1)
    res = {"contain": contain, "weights": W, "mask": any_mask, "dist_w": dist_w, "entropy": entropy, "tsdf": tsdf}
    if color:
        res["rgb"] = (rgb * W[..., None]).sum(1)                                     # :660-663
    else:
        wsdf = (tsdf * W).sum(-1)
        out_sdf = np.full((G,), -1, np.float32)                                      # :461, :525-527
        out_sdf[any_mask] = wsdf[any_mask]
        res["sdf"] = out_sdf
    return res
