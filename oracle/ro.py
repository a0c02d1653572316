"""Oracle: RandomOptimizer particle pose-candidate scoring and update.

TEST INFRASTRUCTURE ONLY.  Restates reference RandomOptimizer.py:54-73, 81-85,
113-131, 154-157, 184-227 and pytorch3d.transforms.quaternion_to_matrix (the
reference's un-pinned dependency; published formula: two_s = 2/(q.q), real
part first, no unit-norm assumption).  The particle template is an explicit
input (the reference draws it with np.random, RandomOptimizer.py:29-32).
"""
import torch


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def pose_6D_to_7D(batch_pose):                                    # RandomOptimizer.py:54-60
    imag_sq_sum = batch_pose[:, 0] ** 2 + batch_pose[:, 1] ** 2 + batch_pose[:, 2] ** 2
    qw = torch.where(imag_sq_sum <= 1.0, torch.sqrt(1 - imag_sq_sum), 0.0).unsqueeze(1)
    return torch.cat([qw, batch_pose], dim=-1)


def get_abs_pose(rot, trans, pst7):                               # :69-73
    return rot @ quaternion_to_matrix(pst7[:, :4]), trans + pst7[:, 4:, None]


def get_fitness(field, abs_rot, abs_trans, target_d, rays_d_cam, trunc, sdf_weight=1000.0):   # :113-131
    cam = rays_d_cam * target_d
    valid = torch.where(target_d > 0.0, torch.ones_like(target_d), torch.zeros_like(target_d)).squeeze(-1)[None]
    world = torch.transpose(abs_rot @ torch.transpose(cam, 0, 1) + abs_trans, 1, 2)          # :81-85
    pred_sdf = field.run_network(world)[..., 3:4].squeeze(-1) * trunc
    mean_sdf = torch.mean(valid * torch.abs(pred_sdf), dim=-1)
    return mean_sdf * sdf_weight, mean_sdf


def ro_iteration(field, rot_cur, trans_cur, search_size, particles, target_d, rays_d_cam, trunc, rescale=0.5):
    """One pass of the loop body RandomOptimizer.py:184-224.  Returns new state and
    the integer decisions (better_mask, count, success_flag, argmin diagnostic)."""
    pst = particles * search_size
    pst7 = pose_6D_to_7D(pst)
    abs_rot, abs_trans = get_abs_pose(rot_cur, trans_cur, pst7)
    fit, mean_sdf_all = get_fitness(field, abs_rot, abs_trans, target_d, rays_d_cam, trunc)
    f0 = fit[0]
    better = torch.where(fit < f0, torch.ones_like(f0), torch.zeros_like(f0))
    weights = (f0 - fit) * better
    wsum = torch.sum(weights) + 0.00001
    count = int(torch.count_nonzero(better))
    success = count > 0
    if success:
        mean_sdf = torch.sum(weights * mean_sdf_all) / wsum
        mt = torch.sum(pst7 * weights[:, None], dim=0) / wsum
        mq = mt[:4] / (mt[:4].norm() + 1e-5)
        mt = torch.cat([mq, mt[4:]], dim=0)
        rot_cur = rot_cur @ quaternion_to_matrix(mt[:4])                                      # :140-146
        trans_cur = trans_cur + mt[4:][..., None]
    else:
        mean_sdf = mean_sdf_all[0]
        mt = torch.tensor([1.0, 0, 0, 0, 0, 0, 0])
    s = torch.abs(mt[1:]) + 0.0001                                                            # :154-157
    ss = rescale * mean_sdf * s / s.norm() + 0.0001
    search_size = (ss if success else ss * 2)[None]
    info = {"fitness": fit, "mean_sdf": mean_sdf_all, "better_mask": better.bool(), "count": count,
            "success": success, "argmin": int(torch.argmin(fit))}
    return rot_cur, trans_cur, search_size, info


def optimize(field, depth_img, rays_dir, row_idx, col_idx, initial_pose, particles, n_iter, trunc,
             init_scale=0.02, rescale=0.5):
    """RandomOptimizer.optimize, RandomOptimizer.py:165-227."""
    rot_cur, trans_cur = initial_pose[:3, :3], initial_pose[:3, 3:]
    search_size = init_scale
    infos = []
    with torch.no_grad():
        for i in range(n_iter):
            off = i % 5
            ih, iw = row_idx + off, col_idx + off
            target_d = depth_img[ih, iw].unsqueeze(-1)
            rays_d_cam = rays_dir[ih, iw, :]
            rot_cur, trans_cur, search_size, info = ro_iteration(
                field, rot_cur, trans_cur, search_size, particles, target_d, rays_d_cam, trunc, rescale)
            infos.append(info)
    T = torch.eye(4)
    T[:3, :3] = rot_cur
    T[:3, 3] = trans_cur.squeeze()
    return T, infos
