"""ORACLE (test infrastructure only): the gradient pose refinement ("GO") loop of tracking, restating reference
MIPSFusion.tracking_render (mipsfusion.py:501-556) with get_pose_param_optim (:235-241: quaternion + translation parameters,
torch.optim.Adam with lr_rot / lr_trans), matrix_from_tensor = qt_to_transform_matrix (geometry_helper.py:11-17) and
get_loss_from_ret (:141-152).  The pixel batch is an explicit input (the reference samples it once, :512-522); the model is the
oracle field (the reference's JointEncoding.forward restated in oracle/scene.py)."""
import torch

from .ro import quaternion_to_matrix
from .shims.pytorch3d.transforms import matrix_to_quaternion


def qt_to_transform_matrix(rot, trans):                           # geometry_helper.py:11-17
    bs = rot.shape[0]
    T = torch.eye(4).to(rot)[None, ...].repeat(bs, 1, 1)
    T[:, :3, :3] = quaternion_to_matrix(rot)
    T[:, :3, 3] = trans
    return T


def refine_pose(field, c2w_init, rays_d_cam, target_s, target_d, n_iter, lr_rot=1e-3, lr_trans=1e-3, wait_iters=100, best=True, u=None):
    """-> (c2w (4,4), per-iteration losses, per-iteration poses).  u: optional (n_iter, N, S) jitter draws."""
    poses = c2w_init[None, ...]
    cur_trans = torch.nn.parameter.Parameter(poses[:, :3, 3].clone())                         # :237
    cur_rot = torch.nn.parameter.Parameter(matrix_to_quaternion(poses[:, :3, :3]).clone())    # :238
    opt = torch.optim.Adam([{"params": cur_rot, "lr": lr_rot}, {"params": cur_trans, "lr": lr_trans}])
    best_loss, best_c2w, thresh = None, None, 0
    losses, seen = [], []
    c2w_est = None
    N = rays_d_cam.shape[0]
    for i in range(n_iter):
        opt.zero_grad()
        c2w_est = qt_to_transform_matrix(cur_rot, cur_trans)
        rays_o = c2w_est[..., :3, -1].repeat(N, 1)                                             # :531
        rays_d = torch.sum(rays_d_cam[..., None, :] * c2w_est[:, :3, :3], -1)                  # :532
        ret = field.forward(rays_o, rays_d, target_s, target_d, None if u is None else u[i], EMD_w=0.)
        loss = field.total_loss(ret)
        losses.append(float(loss)); seen.append(c2w_est.detach()[0].clone())
        if best_loss is None:
            best_loss, best_c2w = float(loss), c2w_est.detach()
        with torch.no_grad():
            c2w_est = qt_to_transform_matrix(cur_rot, cur_trans)
            if float(loss) < best_loss:
                best_loss, best_c2w, thresh = float(loss), c2w_est.detach(), 0
            else:
                thresh += 1
        if thresh > wait_iters:
            break
        loss.backward()
        opt.step()
    out = best_c2w[0] if best else c2w_est.detach()[0]
    return out.clone(), losses, seen
