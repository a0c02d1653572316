"""Drop-in for reference helper_functions/sampling_helper.py (same function names / signatures), with
the selection done by integer-exact CUDA kernels.  The random draws the reference makes on the CPU
(``torch.randn_like`` / python ``random``) become explicit ``keys`` (drawn on the device when omitted),
so that oracle and kernel can be compared bit for bit."""
import torch

from . import _lib as L


def pixel_indices_to_rc(indices, H, W):
    return torch.div(indices, W, rounding_mode="floor"), torch.remainder(indices, W)


def pixel_rc_to_indices(rows, cols, H, W):
    return rows * W + cols


def _device_of(t=None):
    if t is not None and t.is_cuda:
        return t.device
    return torch.device("cuda", torch.cuda.current_device())


def sample_pixels_uniformly(img_h, img_w, num_h, num_w, device=None):
    """rows, cols int64 (num_h*num_w) of the uniform lattice (sampling_helper.py:38-48)."""
    dev = torch.device(device) if device is not None else _device_of()
    rows = torch.empty(num_h * num_w, device=dev, dtype=torch.int64)
    cols = torch.empty_like(rows)
    with torch.cuda.device(dev):
        L.call("mf_sample_pixels_uniform", img_h, img_w, num_h, num_w, L.ptr(rows), L.ptr(cols), L.stream())
    return rows, cols


def _topk(depth_image, num, lattice, keys):
    dev = _device_of(depth_image)
    depth = L.f32c(depth_image, dev)
    H, W = depth.shape
    if keys is None:
        keys = torch.randn(H * W, device=dev).abs()
    keys = L.f32c(keys.reshape(-1), dev)
    ws = torch.empty(int(L.lib().mf_topk_workspace_size(H * W)), device=dev, dtype=torch.uint8)
    lh, lw = lattice
    k = num - lh * lw
    idx = torch.empty(max(k, 0), device=dev, dtype=torch.int64)
    rows = torch.empty(num, device=dev, dtype=torch.int64) if lh > 0 else None
    cols = torch.empty(num, device=dev, dtype=torch.int64) if lh > 0 else None
    with torch.cuda.device(dev):
        L.call("mf_sample_pixels_topk", L.ptr(depth), L.ptr(keys), H, W, lh, lw, num, L.ptr(idx), L.ptr(rows), L.ptr(cols),
               L.ptr(ws), L.stream())
    return idx, rows, cols


def sample_valid_pixels_random(depth_image, num, keys=None):
    """Indices (row-major) of `num` pixels with depth > 0, ordered by descending key (sampling_helper.py:28-32)."""
    return _topk(depth_image, num, (0, 0), keys)[0]


def sample_pixels_mix(img_h, img_w, num_h, num_w, depth_image, num, keys=None):
    """Uniform lattice + random valid pixels off the lattice (sampling_helper.py:53-68)."""
    _, rows, cols = _topk(depth_image, num, (num_h, num_w), keys)
    return rows, cols


def sample_pixels_random(img_h, img_w, num, device=None):
    """`num` distinct pixel indices, uniformly at random (sampling_helper.py:20-22; the reference uses python
    random.sample -- here a device permutation)."""
    dev = torch.device(device) if device is not None else _device_of()
    return torch.randperm(img_h * img_w, device=dev)[:num]
