"""Where a 640x480 frame's time goes: RandomOptimizer / gradient pose refinement / mapping; kernel table of the pose route."""
import os, sys, time, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
import mipsfusion_b200 as mf
from mipsfusion_b200 import synth, sampling_helper as sh
from mipsfusion_b200.mapper import FusedMapper
dev = torch.device("cuda", 0)
cfg = H.make_config(bench.HASH, n_samples_d=50, n_range_d=25)
cfg["tracking"] = {"RO": {"particle_size": 2000, "initial_scaling_factor": 0.02, "rescaling_factor": 0.5, "n_rows": 16, "n_cols": 24},
                   "ignore_edge_W": 20, "ignore_edge_H": 20}
of = H.oracle_field(cfg); model = H.cuda_model(cfg, H.state_of(of))
dirs = synth.camera_rays(); c2w = synth.trajectory(4)[1]; frame = synth.render_frame(c2w, dirs)
depth_d, rgb_d, dirs_d = frame["depth"].to(dev), frame["rgb"].to(dev), dirs.to(dev)
ds = types.SimpleNamespace(H=460, W=620, fx=320.0, fy=320.0, cx=309.5, cy=229.5, rays_d=dirs)
ro = mf.RandomOptimizer(cfg, types.SimpleNamespace(dataset=ds, device=str(dev)))
mapper = FusedMapper(model); tw = cfg["training"]
def t_ro():
    model.eval(); p = ro.optimize(model, frame["depth"], c2w.clone(), c2w.clone(), n_iter=5).to(dev); model.train(); return p
def t_go(pose):
    rows, cols = sh.sample_pixels_mix(460, 620, 16, 24, depth_d, 1000)
    d_cam, t_rgb, t_d = dirs_d[rows, cols], rgb_d[rows, cols], depth_d[rows, cols].unsqueeze(-1)
    trans = pose[:3, 3].clone().requires_grad_(True); rot = pose[:3, :3].clone().requires_grad_(True)
    opt = torch.optim.Adam([rot, trans], lr=1e-3)
    for _ in range(10):
        opt.zero_grad()
        rays_o = trans[None, :].repeat(1000, 1); rays_d = torch.sum(d_cam[..., None, :] * rot[None], -1)
        ret = model(rays_o, rays_d, t_rgb, t_d, EMD_w=0.0)
        loss = tw["rgb_weight"] * ret["rgb_loss"] + tw["sdf_weight"] * ret["sdf_loss"] + tw["fs_weight"] * ret["fs_loss"]
        loss.backward(); opt.step()
    return loss
def t_map(pose):
    idx = torch.randint(0, 460 * 620, (2600,), device=dev); r_, c_ = idx // 620, idx % 620
    ro_, rd_ = pose[None, :3, 3].repeat(2600, 1).contiguous(), torch.sum(dirs_d[r_, c_][..., None, :] * pose[None, :3, :3], -1).contiguous()
    for _ in range(15): mapper.step(ro_, rd_, rgb_d[r_, c_].contiguous(), depth_d[r_, c_].contiguous())
pose = t_ro(); t_go(pose); t_map(pose); torch.cuda.synchronize()
for name, fn in (("RO 5 iters", t_ro), ("GO 10 iters", lambda: t_go(pose)), ("map 15 iters", lambda: t_map(pose))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); print(name, round((time.perf_counter() - t0) / 3 * 1e3, 3), "ms", flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t_go(pose); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=70))
