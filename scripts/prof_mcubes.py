"""Times the marching-cubes path (N2) at 512^3 on a room-like SDF volume: per-call wall time (CUDA events around the two C-ABI calls,
including their host synchronisations) and the CPU oracle on a 128^3 sub-volume.  Usage: python scripts/prof_mcubes.py [n]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mipsfusion_b200 as mf          # noqa: E402
from mipsfusion_b200 import synth     # noqa: E402,F401

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ax = torch.linspace(0, 1, n, device="cuda")
X, Y, Z = ax[:, None, None], ax[None, :, None], ax[None, None, :]
d = torch.minimum(torch.minimum(X - 0.12, 0.9 - Y), torch.minimum(Z - 0.2, torch.sqrt((X - 0.55) ** 2 + (Y - 0.55) ** 2 + (Z - 0.55) ** 2) - 0.17))
vol = torch.tanh(d * 12.0).contiguous()
for it in range(4):
    torch.cuda.synchronize()
    t = time.perf_counter()
    v, f, info = mf.marching_cubes_device(vol, 0.0, 3.0, return_info=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"n={n} run {it}: {dt * 1e3:.2f} ms  verts {v.shape[0]} faces {f.shape[0]} {info}  {n ** 3 / dt / 1e9:.2f} G voxels/s")
from oracle import marching_cubes as omc   # noqa: E402
sub = vol[:128, :128, :128].cpu().numpy()
t = time.perf_counter(); ov, of_ = omc.marching_cubes(sub, 0.0, 3.0); dt = time.perf_counter() - t
print(f"oracle 128^3: {dt:.3f} s = {128 ** 3 / dt / 1e6:.2f} M voxels/s, faces {of_.shape[0]}")
sv, sf = mf.marching_cubes.marching_cubes(sub, 0.0, 3.0)
print("sub-volume parity:", np.array_equal(sv, ov) and np.array_equal(sf, of_))
