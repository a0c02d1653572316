import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import mipsfusion_b200 as mf
from oracle import tracking as otrk
cfg = H.make_config(16, n_samples_d=50, n_range_d=25)
cfg["training"]["perturb"] = 0
cfg["tracking"] = {"lr_rot": 1e-3, "lr_trans": 1e-3, "wait_iters": 100, "best": True}
of = H.oracle_field(cfg, grid_scale=0.3, seed=13)
model = H.cuda_model(cfg, H.state_of(of))
rays7, _, poses, _ = H.synth_batch_packed(500, seed=31)
c2w = poses[0].clone()
d = torch.eye(4); d[:3, 3] = torch.tensor([0.01, -0.008, 0.006]); ang = 0.01
d[:3, :3] = torch.tensor([[1.0, -ang, 0.0], [ang, 1.0, 0.0], [0.0, 0.0, 1.0]]) / (1 + ang * ang) ** 0.5; d[2, 2] = 1.0
start = c2w @ d
_, losses, seen = otrk.refine_pose(of, start, rays7[:, :3], rays7[:, 3:6], rays7[:, 6:7], 10)
for bw in ("fp32", "tc"):
    print(bw)
    for n in range(0, 6):
        r = mf.FusedPoseRefiner(model, use_best=False, backward=bw)
        last, st = r.refine(start, rays7[:, :3].cuda(), rays7[:, 3:6].cuda(), rays7[:, 6].cuda(), n)
        torch.cuda.synchronize()
        dif = (last.cpu() - seen[n]).abs()
        print(n, "max abs diff", float(dif.max()), "rot", float(dif[:3, :3].max()), "trans", float(dif[:3, 3].max()), "state q", st[:7].cpu().numpy())
# oracle quaternion path for reference
from oracle.shims.pytorch3d.transforms import matrix_to_quaternion
print("oracle q0", matrix_to_quaternion(start[None, :3, :3]))
