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
# fused mapper (device batch, host batch, store-fed), keyframe store, device draws, colour joint query, all forward kernels
from mipsfusion_b200.mapper import FusedMapper
model.train()
m = FusedMapper(model)
args4 = [t.cuda().contiguous() for t in (ro_, rd_, rgb, d)]
m.step(*args4)
rays7, pidx, poses, _ = H.synth_batch_packed(200, seed=1)
print(m.step_host(rays7.pin_memory(), pidx.pin_memory(), poses.cuda()))
scfg = {"sampling": {"kf_n_rays_h": 6, "kf_n_rays_w": 8}}
st = mf.KeyframeRayStore(scfg, 60, 80, 4, "cuda")
for k in range(3):
    st.add_keyframe({"direction": dirs, "rgb": torch.rand(60, 80, 3, generator=g), "depth": dep.cpu() * 3, "frame_id": k})
print(m.step_from_store(st, 0, [0, 1, 2], torch.eye(4)[None].repeat(3, 1, 1).cuda(), 40, cur_rays7=st.rays[2, :8].clone()))
print(mf.sample_without_replacement(1000, 10, torch.device("cuda")).shape, st.sample_rays_in_given_kf([2, 0], 9)[0].shape)
model.eval()
print(jq.query(points=torch.rand(500, 3).numpy() * np.array([2.5, 4.4, 3.0]) + np.array([-0.5, 0.6, -1.0]), color=True)["rgb"].shape)
for impl in (2, 3, 1, 0):
    L.call("mf_set_decoder_impl", impl)
    model.run_network(torch.rand(700, 3).cuda()); model.query_sdf(torch.rand(700, 3).cuda())
torch.cuda.synchronize()
print("tc error flag", L.lib().mf_tc_check_error())
