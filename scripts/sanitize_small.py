"""Small end-to-end run (both decoders, forward + backward + RO + joint query) for compute-sanitizer."""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import mipsfusion_b200 as mf
from mipsfusion_b200 import _lib as L
cfg = H.make_config(12, n_samples_d=32, n_range_d=11)
of = H.oracle_field(cfg, grid_scale=0.3)
model = H.cuda_model(cfg, H.state_of(of))
ro_, rd_, rgb, d, u = H.synth_batch(200, 43, seed=1)
for want in (False, True):
    a = ro_.cuda().requires_grad_(want); b = rd_.cuda().requires_grad_(want)
    ret = model(a, b, rgb.cuda(), d.cuda(), u=u.cuda())
    (ret["rgb_loss"] + 1000 * ret["sdf_loss"] + 10 * ret["fs_loss"]).backward()
opt = mf.create_map_optimizer(model, 1e-2, 1e-2); opt.step(zero_grad=True)
model.eval()
q = model.run_network(torch.rand(300, 3).cuda())
g = torch.Generator().manual_seed(0)
from mipsfusion_b200 import sampling_helper as sh
dep = torch.rand(60, 80, generator=g).cuda()
sh.sample_pixels_mix(60, 80, 6, 8, dep, 300)
tc = dict(cfg); tc["tracking"] = {"RO": {"particle_size": 64, "initial_scaling_factor": 0.02, "rescaling_factor": 0.5, "n_rows": 6, "n_cols": 8}, "ignore_edge_W": 2, "ignore_edge_H": 2}
i, j = torch.meshgrid(torch.arange(80, dtype=torch.float32), torch.arange(60, dtype=torch.float32), indexing="xy")
dirs = torch.stack([(i - 39.5) / 40, -(j - 29.5) / 40, -torch.ones_like(i)], -1)
ds = types.SimpleNamespace(H=60, W=80, fx=40.0, fy=40.0, cx=39.5, cy=29.5, rays_d=dirs)
r = mf.RandomOptimizer(tc, types.SimpleNamespace(dataset=ds, device="cuda"))
print(r.optimize(model, dep.cpu() * 3, torch.eye(4), torch.eye(4), n_iter=2))
jq = mf.JointSubmapQuery([model, model], [torch.eye(4), torch.eye(4)], [np.array([-0.5, 0.6, -1.0])] * 2, [np.array([2.0, 5.0, 2.0])] * 2,
                         [np.array([1.0, 3.0, 0.5], dtype=np.float32)] * 2)
axes = mf.get_grid_uniform(np.array([-0.5, 0.6, -1.0]), np.array([2.0, 5.0, 2.0]), voxel_size=0.25)
print(jq.query(axes=axes)["sdf"].shape, L.lib().mf_tc_check_error())
torch.cuda.synchronize()
