"""Shared helpers of the parity tests: build matching (oracle, CUDA) models from the same weights."""
import numpy as np
import torch

from oracle import scene as oscene


def T(a):
    return torch.from_numpy(np.asarray(a))


def make_config(hash_size=19, **training):
    cfg = oscene.default_config()
    cfg["grid"]["hash_size"] = hash_size
    cfg["training"].update(training)
    return cfg


def oracle_field(cfg, state=None, grid_scale=None, seed=0):
    f = oscene.OracleField(cfg, mlp_seed=seed)
    if grid_scale is not None:                       # give the grid visible amplitude (init is U(-1e-4,1e-4))
        g = torch.Generator().manual_seed(100 + seed)
        f.grid = ((torch.rand(f.grid.shape, generator=g) * 2 - 1) * grid_scale).requires_grad_(True)
    if state is not None:
        f.grid = state["embed_fn.params"].clone().requires_grad_(True)
        f.w = {k[len("decoder."):]: v.clone().requires_grad_(True) for k, v in state.items() if k.startswith("decoder.")}
    return f


def state_of(field):
    sd = {"embed_fn.params": field.grid.detach().clone(), "embedpos_fn.params": torch.zeros(0)}
    for k, v in field.w.items():
        sd["decoder." + k] = v.detach().clone()
    return sd


def cuda_model(cfg, state, train=True):
    import mipsfusion_b200 as mf
    bb = torch.tensor(cfg["mapping"]["bound"], dtype=torch.float64)
    nf = torch.tensor(cfg["mapping"]["localMLP_max_len"], dtype=torch.float64)
    m = mf.JointEncoding(cfg, bb, nf)
    m.load_state_dict(state)
    m = m.cuda()
    m.train(train)
    return m


def fixture_state(fx):
    return {k[2:]: T(v) for k, v in fx.items() if k.startswith("w:")}


def _np64(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def rel_err(a, b):
    a, b = _np64(a), _np64(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def synth_batch_packed(R, seed=0, invalid=3):
    """The host batch of the reference's mapping loop (mipsfusion.py:289-318) from a synthetic SDF-room frame:
    rays7 (R,7) = [dir_cam | rgb | depth], pose_idx (R,) int64 (all -1: the current frame), poses (1,4,4), and the
    generator used (for further draws)."""
    from mipsfusion_b200 import synth
    g = torch.Generator().manual_seed(seed)
    c2w = synth.trajectory(4)[1]
    dirs = synth.camera_rays()
    step = max(1, int((dirs.shape[0] * dirs.shape[1] / max(R, 1)) ** 0.5) - 1)
    sub = dirs[::max(step, 1), ::max(step, 1)].contiguous()
    frame = synth.render_frame(c2w, sub, seed=seed)
    rays = synth.frame_rays(frame)
    sel = torch.randperm(rays.shape[0], generator=g)[:R]
    rays = rays[sel].clone()
    assert rays.shape[0] == R, (rays.shape, R)
    rays[:invalid, 6] = 0.0
    return rays.contiguous(), -torch.ones(R, dtype=torch.int64), c2w[None].contiguous(), g


def synth_batch(R, S, seed=0, invalid=3):
    """Rays from a synthetic SDF-room frame: rays_o, rays_d, rgb, depth (R,1), u (R,S)."""
    rays, _, poses, g = synth_batch_packed(R, seed, invalid)
    c2w = poses[0]
    rays_d = torch.sum(rays[:, None, :3] * c2w[None, :3, :3], -1)
    rays_o = c2w[None, :3, 3].repeat(R, 1)
    u = torch.rand(R, S, generator=g)
    return rays_o, rays_d, rays[:, 3:6].contiguous(), rays[:, 6:7].contiguous(), u
