"""Oracle: pixel / ray sampling helpers (integer work -> bit-exact).

TEST INFRASTRUCTURE ONLY.  Restates reference helper_functions/sampling_helper.py:7-68,
datasets/utils.py:4-36 (camera rays) and helper_functions/geometry_helper.py:107-123.
Random keys are explicit inputs: the reference draws ``abs(randn)`` on the CPU
(sampling_helper.py:30,60); here ``keys`` is that tensor.
"""
import torch


def pixel_indices_to_rc(indices, H, W):                       # sampling_helper.py:7-10
    return torch.div(indices, W, rounding_mode="floor"), torch.remainder(indices, W)


def pixel_rc_to_indices(rows, cols, H, W):                    # :13-15
    return rows * W + cols


def sample_pixels_uniformly(img_h, img_w, num_h, num_w):      # :38-48
    interval_h, offset_h = (img_h - num_h) // (num_h + 1), (img_h - num_h) % (num_h + 1)
    interval_w, offset_w = (img_w - num_w) // (num_w + 1), (img_w - num_w) % (num_w + 1)
    row_ids = torch.arange(0, num_h, dtype=torch.int64) * (interval_h + 1) + interval_h + offset_h // 2
    col_ids = torch.arange(0, num_w, dtype=torch.int64) * (interval_w + 1) + interval_w + offset_w // 2
    rows = row_ids[..., None].repeat((1, num_w)).reshape((-1,))
    cols = col_ids[None, ...].repeat((num_h, 1)).reshape((-1,))
    return rows, cols


def topk_indices(samp_v, num):
    """torch.topk(samp_v, num)[1] with the tie rule made explicit: descending
    value, ties broken by ascending index (stable sort)."""
    order = torch.sort(samp_v, descending=True, stable=True)[1]
    return order[:num]


def sample_valid_pixels_random(depth_image, num, keys):       # :28-32
    mask = torch.where(depth_image > 0.0, torch.ones_like(depth_image), torch.zeros_like(depth_image)).flatten()
    return topk_indices(mask * keys.flatten(), num)


def sample_pixels_mix(img_h, img_w, num_h, num_w, depth_image, num, keys):   # :53-68
    rows, cols = sample_pixels_uniformly(img_h, img_w, num_h, num_w)
    mask = torch.where(depth_image > 0.0, torch.ones_like(depth_image), torch.zeros_like(depth_image))
    mask[rows, cols] = 0
    sel = topk_indices(mask.flatten() * keys.flatten(), num - num_h * num_w)
    r2, c2 = pixel_indices_to_rc(sel, img_h, img_w)
    return torch.cat([rows, r2], 0), torch.cat([cols, c2], 0)


def get_camera_rays(H, W, fx, fy, cx, cy):                    # datasets/utils.py:4-36 (OpenGL)
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing="xy")
    return torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)


def rays_camera_to_world2(rays_d_cam, c2w_mats, pose_indices):   # geometry_helper.py:120-123
    rays_o = c2w_mats[pose_indices, :3, -1]
    rays_d = torch.sum(rays_d_cam[..., None, :] * c2w_mats[pose_indices, :3, :3], -1)
    return rays_d, rays_o
