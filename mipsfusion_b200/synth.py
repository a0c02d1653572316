"""Synthetic RGB-D frames rendered from an analytic SDF room with known poses.

No dataset or checkpoint is available offline, so correctness tests, the smoke
run and bench.py all use these frames (BASELINE.json north_star):
640x480 pinhole camera with the FastCaMo-synth intrinsics (reference
configs/FastCaMo-synth/FastCaMo-synth.yaml:93-103), OpenGL ray convention
(reference datasets/utils.py:29), 10 px edge crop -> 460x620 working resolution
(reference datasets/dataset.py:259-263), ~2 % invalid (0) depth pixels.
Host-side numpy/torch only; nothing here is on the GPU hot path.
"""
import math
import numpy as np
import torch

ROOM_MIN = np.array([-0.4, 0.7, -0.95], dtype=np.float32)
ROOM_MAX = np.array([2.75, 6.85, 2.85], dtype=np.float32)
BOUND = [[-0.6, 2.95], [0.5, 7.05], [-1.15, 3.05]]   # configs/FastCaMo-synth/apartment_2.yaml:4


def scene_sdf(p):
    """p: (..., 3) torch fp32 -> signed distance (positive in free space)."""
    lo, hi = torch.from_numpy(ROOM_MIN), torch.from_numpy(ROOM_MAX)
    room = torch.minimum(p - lo, hi - p).min(-1).values              # inside of a box
    sph = (p - torch.tensor([1.2, 3.0, 0.2])).norm(dim=-1) - 0.55
    q = (p - torch.tensor([0.6, 5.2, -0.45])).abs() - torch.tensor([0.45, 0.6, 0.5])
    box = q.clamp(min=0).norm(dim=-1) + q.max(-1).values.clamp(max=0)
    return torch.minimum(torch.minimum(room, sph), box)


def scene_color(p):
    return 0.5 + 0.5 * torch.sin(p * torch.tensor([3.1, 2.3, 4.7]) + torch.tensor([0.0, 1.0, 2.0]))


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """c2w (4,4) fp32, OpenGL camera (looks down -z, +y up)."""
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    f = target - eye; f /= np.linalg.norm(f)
    r = np.cross(f, up); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = r, u, -f, eye
    return torch.from_numpy(T.astype(np.float32))


def camera_rays(H=480, W=640, fx=320.0, fy=320.0, cx=319.5, cy=239.5, crop=10):
    """Camera-frame directions (H-2c, W-2c, 3); reference datasets/utils.py:4-36 + edge crop."""
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing="xy")
    d = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    return d[crop:H - crop, crop:W - crop].contiguous() if crop > 0 else d


def render_frame(c2w, dirs=None, invalid_frac=0.02, seed=0, steps=96):
    """Sphere-trace one frame.  Returns dict(rgb (H,W,3), depth (H,W) [z-depth, metres], direction, c2w)."""
    dirs = camera_rays() if dirs is None else dirs
    H, W, _ = dirs.shape
    d_cam = dirs.reshape(-1, 3)
    d_w = d_cam @ c2w[:3, :3].T
    o = c2w[:3, 3][None].expand_as(d_w)
    t = torch.zeros(d_w.shape[0])
    for _ in range(steps):
        t = t + scene_sdf(o + d_w * t[:, None]).clamp(min=0) * 0.98
    hit = o + d_w * t[:, None]
    depth = t.clone()                         # |d_cam.z| == 1  ->  ray parameter == z-depth
    rgb = scene_color(hit)
    g = torch.Generator().manual_seed(seed)
    bad = torch.rand(depth.shape[0], generator=g) < invalid_frac
    depth[bad] = 0.0
    return {"rgb": rgb.reshape(H, W, 3).contiguous(), "depth": depth.reshape(H, W).contiguous(),
            "direction": dirs, "c2w": c2w}


def trajectory(n, seed=0):
    """n smooth c2w poses orbiting inside the room."""
    poses = []
    for k in range(n):
        a = 2 * math.pi * k / max(n, 8) * 0.35
        eye = (1.2 + 0.5 * math.cos(a), 3.6 + 1.2 * math.sin(a), 1.0 + 0.1 * math.sin(2 * a))
        tgt = (1.2 - 1.0 * math.cos(a + 0.3), 3.0 - 1.5 * math.sin(a + 0.3), 0.4)
        poses.append(look_at(eye, tgt))
    return torch.stack(poses)


def frame_rays(frame):
    """(H*W, 7) [dir_cam(3), rgb(3), depth(1)] -- the layout mipsfusion.py:289-290 builds."""
    return torch.cat([frame["direction"], frame["rgb"], frame["depth"][..., None]], -1).reshape(-1, 7)
