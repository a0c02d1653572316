"""Full-image rendering: reference Logger.render_full_img (Logger.py:193-214) on the fused kernels.

The reference walks the image in batches of 10,000 rays, moving every batch host -> device and concatenating the results;
here the whole 620 x 460 image (285,200 rays x 75 samples = 21.4 M field evaluations) is ray-generated on the device
(mf_gen_rays) and rendered by one ``JointEncoding.render_rays`` call -- or in ``ray_batch_size`` pieces when memory is to be
bounded (the decoder output of a batch is R x S x 10 floats)."""
import ctypes as C

import torch

from . import _lib as L


@torch.no_grad()
def render_full_img(model, rays_d_cam, pose_local, gt_depth, ray_batch_size=None, u=None):
    """rays_d_cam (H,W,3) camera-frame directions (``dataset.rays_d``), pose_local (4,4) camera -> submap frame, gt_depth (H,W).
    -> rgb (H,W,3), depth (H,W) on the model's device (same return as the reference).  ``model`` must be in eval() mode or
    have ``perturb`` 0 for a deterministic image (the reference renders under no_grad with whatever mode the model is in)."""
    dev = model._device
    Hh, Ww = gt_depth.shape[-2], gt_depth.shape[-1]
    dirs = L.f32c(rays_d_cam.reshape(-1, 3), dev)
    depth = L.f32c(gt_depth.reshape(-1), dev)
    pose = L.f32c(pose_local.reshape(1, 4, 4), dev)
    R = dirs.shape[0]
    rays_o = torch.empty(R, 3, device=dev, dtype=torch.float32)
    rays_d = torch.empty(R, 3, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        L.call("mf_gen_rays", L.ptr(dirs), L.ptr(pose), None, L.ptr(rays_o), L.ptr(rays_d), R, 1, L.stream())   # rays_camera_to_world
    step = R if ray_batch_size is None else int(ray_batch_size)
    rgb = torch.empty(R, 3, device=dev, dtype=torch.float32)
    dep = torch.empty(R, device=dev, dtype=torch.float32)
    for i in range(0, R, step):
        ret = model.render_rays(rays_o[i:i + step], rays_d[i:i + step], depth[i:i + step, None], u=None if u is None else u[i:i + step])
        rgb[i:i + step] = ret["rgb"]
        dep[i:i + step] = ret["depth"]
    return rgb.reshape(Hh, Ww, 3), dep.reshape(Hh, Ww)
