"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the containment tests of the reference's submap manager
(Manager.find_highest_containing_ratio / compute_containing_ratio, Manager.py:159-244; pts_in_bbox,
helper_functions/geometry_helper.py:193-203).  Pinned against the reference's own methods by tests/golden/manager.npz
(tests/golden/make_golden.py).  Never imported by the product."""
import torch

from .sampling import sample_pixels_uniformly


def pts_in_bbox(pts, xyz_min, xyz_max):
    """geometry_helper.py:193-203."""
    return torch.stack([torch.logical_and((pts > xyz_min[i]).all(dim=-1), (pts < xyz_max[i]).all(dim=-1)) for i in range(xyz_min.shape[0])], dim=-1)


def lattice_points(H, W, depth_img, rays_d, pose_world, rays_h, rays_w):
    """Manager.py:163-174 / :203-214: lattice pixels -> world surface points (n,3), their depths (n,)."""
    ih, iw = sample_pixels_uniformly(H, W, rays_h, rays_w)
    target_d = depth_img[ih, iw]
    rays_d_cam = rays_d[ih, iw]
    n = rays_h * rays_w
    rays_o = pose_world[:3, -1].repeat(n, 1)
    rays_dw = torch.sum(rays_d_cam[..., None, :] * pose_world[None, :3, :3], -1)
    pts = rays_o[..., None, :] + rays_dw[..., None, :] * target_d[..., None, None]
    return pts.reshape(-1, 3), target_d


def containing_scores(H, W, depth_img, rays_d, pose_world, centers, lens, rays_h=15, rays_w=20):
    """Manager.py:163-182.  NB :174 multiplies (P,1,3) by target_d[..., :, None] = (P,1), which broadcasts to (P,P,3): every ray
    direction with every depth; restated as written."""
    ih, iw = sample_pixels_uniformly(H, W, rays_h, rays_w)
    target_d, rays_d_cam = depth_img[ih, iw], rays_d[ih, iw]
    n = rays_h * rays_w
    rays_o = pose_world[:3, -1].repeat(n, 1)
    rays_dw = torch.sum(rays_d_cam[..., None, :] * pose_world[None, :3, :3], -1)
    pts = rays_o[..., None, :] + rays_dw[..., None, :] * target_d[..., :, None]
    return torch.count_nonzero(pts_in_bbox(pts.reshape((-1, 3)), centers - 0.5 * lens, centers + 0.5 * lens), dim=0)


def compute_containing_ratio(H, W, depth_img, rays_d, pose_world, center, length, rays_h=150, rays_w=200):
    """Manager.py:216-244 (length already clamped by the caller)."""
    pts, target_d = lattice_points(H, W, depth_img, rays_d, pose_world, rays_h, rays_w)
    mask = pts_in_bbox(pts, (center - 0.5 * length)[None], (center + 0.5 * length)[None])
    depth_mask = torch.where(target_d[..., None] > 0., torch.ones_like(target_d[..., None]), torch.zeros_like(target_d[..., None]))
    mask = mask.to(depth_mask) * depth_mask
    return torch.count_nonzero(mask) / torch.count_nonzero(depth_mask)
