"""Pose-gradient accuracy: fp32 route vs tcgen05-forward + fp32-backward, against the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
for (R, S, nsd, nrd, T, seed) in [(512, 43, 32, 11, 19, 512), (300, 75, 50, 25, 16, 300), (1000, 75, 50, 25, 19, 7), (1000, 75, 50, 25, 19, 8)]:
    cfg = H.make_config(T, n_samples_d=nsd, n_range_d=nrd)
    of = H.oracle_field(cfg, grid_scale=0.3, seed=seed)
    rays_o, rays_d, rgb, d, u = H.synth_batch(R, S, seed=seed, invalid=3)
    roo = rays_o.clone().requires_grad_(True); rdo = rays_d.clone().requires_grad_(True)
    of.total_loss(of.forward(roo, rdo, rgb, d, u)).backward()
    t = cfg["training"]
    for mixed in (False, True):
        model = H.cuda_model(cfg, H.state_of(of)); model.pose_route_tc_forward = mixed
        ro = rays_o.cuda().requires_grad_(True); rd = rays_d.cuda().requires_grad_(True)
        ret = model(ro, rd, rgb.cuda(), d.cuda(), u=u.cuda())
        (t["rgb_weight"] * ret["rgb_loss"] + t["sdf_weight"] * ret["sdf_loss"] + t["fs_weight"] * ret["fs_loss"]).backward()
        # pose-level gradients: sum over rays (translation) and the rotation-like contraction
        eo, ed = H.rel_err(ro.grad.cpu(), roo.grad), H.rel_err(rd.grad.cpu(), rdo.grad)
        so = H.rel_err(ro.grad.sum(0).cpu(), roo.grad.sum(0)); sd = H.rel_err((rd.grad.cpu()[:, :, None] * rays_d[:, None, :]).sum(0), (rdo.grad[:, :, None] * rays_d[:, None, :]).sum(0))
        print(f"R={R} S={S} T={T} mixed={mixed}: per-ray d_o {eo:.2e} d_d {ed:.2e} | summed trans {so:.2e} rot {sd:.2e} | grid grad {H.rel_err(model.embed_fn.params.grad.cpu(), of.grid.grad):.2e}", flush=True)
