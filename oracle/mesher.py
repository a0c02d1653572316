"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the mesh visibility filter of the reference's Mesher: `point_mask`
(model/Mesher.py:247-281: a vertex is seen if it projects inside some keyframe's image, 20 px from the border, in front of the
camera and nearer than that keyframe's largest stored depth), `project_to_pixel` (helper_functions/geometry_helper.py:216-222)
and `get_face_mask` (model/Mesher.py:221-231 with vis/math_helper.py:17-21: a face is dropped only if ALL its vertices are
unseen).  Pinned against the reference's own methods by tests/golden/mesher.npz (tests/golden/make_mesher_golden.py).  Never
imported by the product."""
import numpy as np
import torch


def project_to_pixel(K, pts):
    """geometry_helper.py:216-222 (pts (n,3,1); negates x IN PLACE like the reference)."""
    pts[:, 0] *= -1
    uv = (K @ pts).squeeze(-1)
    z = uv[:, -1:] + 1e-5
    return (uv[:, :2] / z).float()


def point_mask(points, kf_max_depth, kf_pose_c2w, K, W, H, edge=20):
    """Mesher.py:247-281.  points (n,3); kf_max_depth (k,) = max of each selected keyframe's stored depths (:273);
    kf_pose_c2w (k,4,4) -> bool (n,)."""
    points = points.to(torch.float32)
    seen = torch.zeros_like(points[:, 0], dtype=torch.bool)
    w2c = kf_pose_c2w.inverse()
    rot, trans = w2c[:, :3, :3], w2c[:, :3, 3]
    rotated = torch.sum(points[None, :, None, :] * rot[:, None, :, :], -1)
    transed = rotated + trans[:, None, :]
    for i in range(kf_pose_c2w.shape[0]):
        cam = transed[i]
        uv = project_to_pixel(K, cam.unsqueeze(-1))
        m1 = (uv[:, 0] < W - edge) * (uv[:, 0] > edge) * (uv[:, 1] < H - edge) * (uv[:, 1] > edge)
        m1 = m1 & (cam[..., -1] < 0)
        cz = torch.abs(cam[:, -1])
        m2 = (cz > 0) * (cz < kf_max_depth[i])
        seen = torch.logical_or(seen, m1 & m2)
    return seen


def get_face_mask(vert_mask, faces):
    """Mesher.py:221-231: face kept unless every vertex is unseen (reduce_and = cumprod over the three flags)."""
    unseen = np.logical_not(np.asarray(vert_mask))[np.asarray(faces).astype(np.int64)]
    return ~(np.cumprod(unseen.astype(np.float32), axis=-1)[:, -1].astype(bool))
