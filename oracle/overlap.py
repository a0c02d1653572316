"""ORACLE (test infrastructure only): the overlap SDF difference of the inactive-map global BA,
a composition of run_network with pose-dependent points.  Restates InactiveMap.infer_pts / get_SDF_dif / get_SDF_dif2
(InactiveMap.py:128-192) and compute_avg_SDF_difference / compute_avg_RGB_difference
(helper_functions/geometry_helper.py:225-237).  Pinned against the reference's own methods in tests/golden/overlap.npz.
`models` is any list of objects with a ``run_network(pts (n,3)) -> (n,10)`` method (the oracle field on the CPU, the CUDA
JointEncoding on the GPU); all tensors must live on the models' device."""
import torch


def compute_avg_sdf_difference(pred_sdf1, pred_sdf2, mask):                      # geometry_helper.py:225-229
    loss = torch.sum(torch.square(pred_sdf1 * mask - pred_sdf2 * mask))
    return loss / (torch.count_nonzero(mask) + 0.001)


def compute_avg_rgb_difference(pred_rgb1, pred_rgb2, mask):                      # geometry_helper.py:232-236
    losses = torch.where(mask.squeeze(-1) > 0., torch.sum(torch.abs(pred_rgb1 - pred_rgb2), dim=1), 0.)
    return torch.sum(torch.square(losses)) / (torch.count_nonzero(mask) + 0.001)


def infer_pts(local_poses, model, rays_d_cam, target_d, trunc_value):           # InactiveMap.py:128-139
    rays_d = torch.sum(rays_d_cam[..., None, None, :] * local_poses[..., None, :3, :3], -1)
    rays_o = local_poses[..., None, :3, -1].repeat(1, rays_d.shape[1], 1).reshape(-1, 3)
    rays_d = rays_d.reshape(-1, 3)
    pts_local = (rays_o[..., None, :] + rays_d[..., None, :] * target_d[..., :, None]).reshape(-1, 3)
    rgb_sdf = model.run_network(pts_local)
    return rgb_sdf[..., :3], rgb_sdf[..., 3:4] * trunc_value


def get_sdf_dif(models, rays, ovlp_kf_pose, id1, id2, first_kf_pose1, first_kf_pose2, trunc_value):   # InactiveMap.py:149-165
    rays_d_cam, target_d = rays[..., :3], rays[..., 6:7]
    depth_mask = torch.where(target_d > 0., torch.ones_like(target_d), torch.zeros_like(target_d))
    return get_sdf_dif2(models, target_d, rays_d_cam, depth_mask, ovlp_kf_pose, id1, id2, first_kf_pose1, first_kf_pose2, trunc_value)


def get_sdf_dif2(models, target_d, rays_d_cam, mask, ovlp_kf_pose, id1, id2, first_kf_pose1, first_kf_pose2, trunc_value):  # :177-192
    mask = mask.to(target_d)
    local_poses1 = first_kf_pose1.inverse() @ ovlp_kf_pose
    local_poses2 = first_kf_pose2.inverse() @ ovlp_kf_pose
    rgb1, sdf1 = infer_pts(local_poses1, models[int(id1)], rays_d_cam, target_d, trunc_value)
    rgb2, sdf2 = infer_pts(local_poses2, models[int(id2)], rays_d_cam, target_d, trunc_value)
    return compute_avg_sdf_difference(sdf1, sdf2, mask) + 0. * compute_avg_rgb_difference(rgb1, rgb2, mask)
