"""Generates tests/golden/mesher.npz by calling the REFERENCE's own `Mesher.point_mask` and `Mesher.get_face_mask`
(model/Mesher.py:221-281), imported unchanged from /root/reference, unbound on a stand-in `self` that carries only what the two
methods read (device, K, config['cam'], kfSet.rays).  The module's top-level imports of open3d / skimage / trimesh (absent
offline, unused by these two methods) are satisfied by empty stand-in modules.  Run in the build container only:
    python tests/golden/make_mesher_golden.py"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("MIPSFUSION_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, REF)
np.bool = bool
for name in ("open3d", "trimesh", "skimage", "skimage.measure"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["skimage"].measure = sys.modules["skimage.measure"]

from model.Mesher import Mesher                     # noqa: E402
from mipsfusion_b200 import synth                   # noqa: E402

g = torch.Generator().manual_seed(5)
W, H, fx, fy, cx, cy = 640, 480, 320.0, 320.0, 319.5, 239.5
poses = synth.trajectory(8)[:5].clone()              # (5,4,4) camera-to-world
k = poses.shape[0]
# stored keyframe rays: (num_kf, n_rays, 7), last channel = depth (model/keyframeSet.py); only its maximum is read (:273)
rays = torch.rand(k + 2, 300, 7, generator=g)
rays[..., -1] = 0.5 + 4.0 * torch.rand(k + 2, 300, generator=g)
kf_ids = torch.tensor([0, 2, 3, 5, 6])
# points: a cloud around the room plus points placed exactly on frustum borders and at z = 0 of a camera
pts = torch.tensor([1.0, 3.5, 1.0]) + 4.0 * (torch.rand(6000, 3, generator=g) - 0.5)
stub = types.SimpleNamespace(device="cpu", K=torch.tensor([[fx, 0., cx], [0., fy, cy], [0., 0., 1.]]),
                             config={"cam": {"W": W, "H": H}}, kfSet=types.SimpleNamespace(rays=rays))
seen = Mesher.point_mask(stub, pts.clone(), kf_ids, poses.clone())
faces = torch.randint(0, pts.shape[0], (9000, 3), generator=g).numpy()
face_seen = Mesher.get_face_mask(stub, seen.numpy(), faces)
out = dict(points=pts.numpy(), poses=poses.numpy(), rays_depth=rays[..., -1].numpy(), kf_ids=kf_ids.numpy(), seen=seen.numpy(),
           faces=faces, face_seen=face_seen, cam=np.array([W, H, fx, fy, cx, cy], np.float64))
print("seen", int(seen.sum()), "of", seen.numel(), "| faces kept", int(face_seen.sum()), "of", face_seen.size)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mesher.npz"), **out)
