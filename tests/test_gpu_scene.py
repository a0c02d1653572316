"""GPU parity: JointEncoding (query, z sampling, render, losses, backward, Adam) against the oracle
and against vectors produced by the reference's own code (tests/golden/scene.npz)."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import adam as oadam

pytestmark = pytest.mark.gpu
FIELD_RTOL = 1e-3            # BASELINE.json: rel 1e-3 on fields, losses and gradients


def test_scene_golden_forward_backward(golden):
    fx = golden("scene")
    cfg = H.make_config(int(fx["hash_size"]))
    model = H.cuda_model(cfg, H.fixture_state(fx))
    rays = H.T(fx["rays"])
    ro = H.T(fx["rays_o"]).cuda().requires_grad_(True)
    rd = H.T(fx["rays_d"]).cuda().requires_grad_(True)
    u = H.T(fx["u"]).cuda()
    ret = model(ro, rd, rays[:, 3:6].cuda(), rays[:, 6:7].cuda(), u=u)
    assert set(ret.keys()) == {"rgb", "depth", "rgb_loss", "depth_loss", "sdf_loss", "fs_loss", "psnr"}
    for k in ("rgb_loss", "depth_loss", "sdf_loss", "fs_loss", "psnr"):
        np.testing.assert_allclose(float(ret[k]), float(fx[k]), rtol=1e-4, err_msg=k)
    assert H.rel_err(ret["rgb"].detach().cpu(), fx["rgb"]) < 1e-4
    assert H.rel_err(ret["depth"].detach().cpu(), fx["depth"]) < 1e-4
    t = cfg["training"]
    loss = t["rgb_weight"] * ret["rgb_loss"] + t["sdf_weight"] * ret["sdf_loss"] + t["fs_weight"] * ret["fs_loss"]
    np.testing.assert_allclose(float(loss), float(fx["loss"]), rtol=1e-4)
    loss.backward()
    assert H.rel_err(model.embed_fn.params.grad.cpu(), fx["g:embed_fn.params"]) < FIELD_RTOL
    for name, p in model.decoder.named_parameters():
        assert H.rel_err(p.grad.cpu(), fx["g:decoder." + name]) < FIELD_RTOL, name
    assert H.rel_err(ro.grad.cpu(), fx["g_rays_o"]) < FIELD_RTOL
    assert H.rel_err(rd.grad.cpu(), fx["g_rays_d"]) < FIELD_RTOL
    # eval-mode render dict + z bit-exact + point queries
    model.eval()
    rend = model(ro.detach(), rd.detach(), None, rays[:, 6:7].cuda(), u=u)
    assert set(rend.keys()) == {"rgb", "depth", "disp_map", "acc_map", "depth_var", "z_vals", "raw"}
    assert np.array_equal(rend["z_vals"].detach().cpu().numpy(), fx["z_vals"])             # same fp32 op order as torch
    assert H.rel_err(rend["raw"].cpu(), fx["raw"]) < 1e-4
    for k in ("depth_var", "acc_map", "disp_map"):
        assert H.rel_err(rend[k].cpu(), fx[k]) < 1e-4, k
    q = model.run_network(H.T(fx["q_pts"]).cuda())
    assert H.rel_err(q.detach().cpu(), fx["q_out"]) < 1e-4
    assert model.query_sdf(torch.rand(7, 3).cuda()).shape == (7, 1)


@pytest.mark.parametrize("R,S,nsd,nrd,T,bound", [(512, 43, 32, 11, 19, True), (300, 75, 50, 25, 16, False), (2, 43, 32, 11, 14, True)])
def test_scene_vs_oracle(R, S, nsd, nrd, T, bound):
    cfg = H.make_config(T, n_samples_d=nsd, n_range_d=nrd)
    cfg["grid"]["use_bound_normalize"] = bound
    if not bound:
        cfg["mapping"]["localMLP_max_len"] = [7.0, 7.0, 4.0]
    of = H.oracle_field(cfg, grid_scale=0.3, seed=R)
    model = H.cuda_model(cfg, H.state_of(of))
    rays_o, rays_d, rgb, d, u = H.synth_batch(R, S, seed=R, invalid=min(3, R - 1))   # (the reference's depth-loss indexing needs R >= 2)
    roo = rays_o.clone().requires_grad_(True); rdo = rays_d.clone().requires_grad_(True)
    ret_o = of.forward(roo, rdo, rgb, d, u)
    of.total_loss(ret_o).backward()
    ro = rays_o.cuda().requires_grad_(True); rd = rays_d.cuda().requires_grad_(True)
    ret = model(ro, rd, rgb.cuda(), d.cuda(), u=u.cuda())
    t = cfg["training"]
    (t["rgb_weight"] * ret["rgb_loss"] + t["sdf_weight"] * ret["sdf_loss"] + t["fs_weight"] * ret["fs_loss"]).backward()
    for k in ("rgb_loss", "depth_loss", "sdf_loss", "fs_loss"):
        np.testing.assert_allclose(float(ret[k]), float(ret_o[k]), rtol=FIELD_RTOL, err_msg=k)
    assert H.rel_err(ret["rgb"].detach().cpu(), ret_o["rgb"].detach()) < FIELD_RTOL
    assert H.rel_err(ret["depth"].detach().cpu(), ret_o["depth"].detach()) < FIELD_RTOL
    assert H.rel_err(model.embed_fn.params.grad.cpu(), of.grid.grad) < FIELD_RTOL
    for name, p in model.decoder.named_parameters():
        assert H.rel_err(p.grad.cpu(), of.w[name].grad) < FIELD_RTOL, name
    assert H.rel_err(ro.grad.cpu(), roo.grad) < FIELD_RTOL
    assert H.rel_err(rd.grad.cpu(), rdo.grad) < FIELD_RTOL


def test_integer_streams_bit_exact():
    """z values, first-sign-change index and the global mask counts (integer work) match the oracle."""
    import ctypes as C
    from mipsfusion_b200 import _lib as L
    cfg = H.make_config(14, n_samples_d=32, n_range_d=11)
    of = H.oracle_field(cfg, grid_scale=0.5, seed=5)
    model = H.cuda_model(cfg, H.state_of(of))
    R, S = 777, 43
    rays_o, rays_d, rgb, d, u = H.synth_batch(R, S, seed=4, invalid=9)
    with torch.no_grad():
        ret_o = of.forward(rays_o, rays_d, rgb, d, u)
    rgb_c, depth_c, aux, z, raw, counts = model._render(rays_o.cuda(), rays_d.cuda(), rgb.cuda(), d.cuda(), u.cuda(), 0.01)[:6]
    assert np.array_equal(z.cpu().numpy(), ret_o["z_vals"].numpy())
    assert list(counts.cpu().numpy()) == list(ret_o["counts"])
    # sign-change index on the oracle's own raw values (so a 1-ulp sdf difference cannot flip it)
    cfg_c, _ = model._render_cfg(True, 0.01, torch.device("cuda"))
    raw_o = ret_o["raw"].cuda().contiguous()
    inds = torch.empty(R, device="cuda", dtype=torch.int32)
    o3 = torch.empty(R, 3, device="cuda"); o1 = torch.empty(R, device="cuda")
    L.call("mf_render_loss_fwd", L.ptr(raw_o), L.ptr(z), None, None, None, C.byref(cfg_c), L.ptr(o3), L.ptr(o1), None, None,
           L.ptr(inds), None, None, R, S, L.stream())
    assert np.array_equal(inds.cpu().numpy().astype(np.int64), ret_o["inds"].numpy())
    w = model.raw2outputs(raw_o, z)[3]
    sums = w.sum(-1).cpu().numpy()
    assert np.all((np.abs(sums - 1.0) < 1e-4) | (sums < 1e-6))                     # weights are L1-normalised


def test_full_size_properties():
    """BASELINE C1 shape (4096 x 43, T=2^19): size-independent properties instead of an oracle run."""
    cfg = H.make_config(19, n_samples_d=32, n_range_d=11)
    of = H.oracle_field(cfg, grid_scale=0.2, seed=2)
    model = H.cuda_model(cfg, H.state_of(of))
    R, S = 4096, 43
    rays_o, rays_d, rgb, d, u = H.synth_batch(R, S, seed=8)
    args = [t.cuda() for t in (rays_o, rays_d, rgb, d)]
    ret = model(*args, u=u.cuda())
    for k in ("rgb_loss", "depth_loss", "sdf_loss", "fs_loss"):
        assert np.isfinite(float(ret[k])), k
    (1000 * ret["sdf_loss"] + 10 * ret["fs_loss"] + ret["rgb_loss"]).backward()
    g1 = model.embed_fn.params.grad.clone(); m1 = model.decoder.pts_linear[2].weight.grad.clone()
    model.zero_grad(set_to_none=True)
    ret = model(*args, u=u.cuda())
    (2 * (1000 * ret["sdf_loss"] + 10 * ret["fs_loss"] + ret["rgb_loss"])).backward()
    # linearity of the backward in the upstream gradient (atomics reorder sums -> small tolerance)
    assert H.rel_err(model.embed_fn.params.grad.cpu(), (2 * g1).cpu()) < 1e-4
    assert H.rel_err(model.decoder.pts_linear[2].weight.grad.cpu(), (2 * m1).cpu()) < 1e-4
    # a sub-batch of the full batch agrees with the oracle
    sub = slice(0, 256)
    with torch.no_grad():
        raw_o = of.run_network((rays_o[sub, None, :] + rays_d[sub, None, :] * torch.linspace(0.3, 3.0, 16)[None, :, None]))
        raw_c = model.run_network((args[0][sub, None, :] + args[1][sub, None, :] * torch.linspace(0.3, 3.0, 16).cuda()[None, :, None]))
    assert H.rel_err(raw_c.cpu(), raw_o) < FIELD_RTOL


def test_adam_matches_torch_and_oracle():
    import mipsfusion_b200 as mf
    g = torch.Generator().manual_seed(2)
    for eps, wd, n in ((1e-15, 0.0, 100003), (1e-8, 1e-6, 36577)):
        p0 = torch.randn(n, generator=g)
        pt = p0.clone().requires_grad_(True)
        pc = p0.clone().cuda().requires_grad_(True)
        ot = torch.optim.Adam([{"params": [pt], "eps": eps, "weight_decay": wd, "lr": 1e-2}], betas=(0.9, 0.99))
        oc = mf.FusedAdam([{"params": [pc], "eps": eps, "weight_decay": wd, "lr": 1e-2}], betas=(0.9, 0.99))
        q, m, v = p0.clone(), torch.zeros(n), torch.zeros(n)
        for step in range(1, 7):
            grad = torch.randn(n, generator=g) * (0 if step == 3 else 1) * 1e-3
            pt.grad = grad.clone(); pc.grad = grad.clone().cuda()
            ot.step(); oc.step(zero_grad=True)
            oadam.adam_step(q, grad, m, v, step, 1e-2, 0.9, 0.99, eps, wd)
            assert float(pc.grad.abs().sum()) == 0.0
        np.testing.assert_allclose(pc.detach().cpu().numpy(), pt.detach().numpy(), rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(pc.detach().cpu().numpy(), q.numpy(), rtol=2e-6, atol=1e-7)


def test_training_loop_parity():
    """30 mapping steps (forward + backward + Adam) track the oracle's loss curve and end at the same weights."""
    import mipsfusion_b200 as mf
    cfg = H.make_config(12, n_samples_d=32, n_range_d=11)
    of = H.oracle_field(cfg, seed=7)
    model = H.cuda_model(cfg, H.state_of(of))
    R, S = 256, 43
    rays_o, rays_d, rgb, d, _ = H.synth_batch(R, S, seed=11)
    opt_o = oadam.make_optimizer(of)
    opt_c = mf.create_map_optimizer(model, 1e-2, 1e-2)
    g = torch.Generator().manual_seed(0)
    args = [t.cuda() for t in (rays_o, rays_d, rgb, d)]
    t = cfg["training"]
    lo, lc = [], []
    for it in range(30):
        u = torch.rand(R, S, generator=g)
        opt_o.zero_grad()
        ret_o = of.forward(rays_o, rays_d, rgb, d, u)
        loss_o = of.total_loss(ret_o); loss_o.backward(); opt_o.step()
        ret = model(*args, u=u.cuda())
        loss = t["rgb_weight"] * ret["rgb_loss"] + t["sdf_weight"] * ret["sdf_loss"] + t["fs_weight"] * ret["fs_loss"]
        loss.backward(); opt_c.step(zero_grad=True)
        lo.append(float(loss_o)); lc.append(float(loss))
    lo_a, lc_a = np.array(lo), np.array(lc)
    dw = np.abs(model.decoder.pts_linear[0].weight.detach().cpu().numpy() - of.w["pts_linear.0.weight"].detach().numpy())
    print("\n  loss curve max rel diff", float(np.max(np.abs(lc_a - lo_a) / np.abs(lo_a))), "first 5 steps", float(np.max(np.abs(lc_a[:5] - lo_a[:5]) / np.abs(lo_a[:5]))))
    print("  weights after 30 steps: |diff| quantiles 50/99/100 %", np.quantile(dw, 0.5), np.quantile(dw, 0.99), dw.max())
    # The first steps are a pure fp32-vs-tensor-core comparison: inside north_star's 1e-3 (measured 4.6e-4).  After that the
    # comparison measures Adam, not the kernels: m / sqrt(v) turns the last digits of a near-zero gradient component into a
    # +-lr step (lr = 1e-2 here), for ANY two fp32 implementations, and 30 such steps separate the two trajectories by a few
    # 1e-3 on the loss (measured 3.8e-3) and by up to ~2 lr on single weights (measured: median 7e-4, 99 % 8e-3, max 2.2e-2).
    np.testing.assert_allclose(lc_a[:5], lo_a[:5], rtol=1e-3)
    np.testing.assert_allclose(lc, lo, rtol=5e-3)
    assert np.quantile(dw, 0.5) < 2e-3 and np.quantile(dw, 0.99) < 2e-2 and dw.max() < 5e-2, (np.quantile(dw, 0.5), np.quantile(dw, 0.99), dw.max())
    assert lc[-1] < lc[0]


@pytest.mark.gpu
def test_fused_mapper_matches_oracle_and_host_route():
    """FusedMapper.step (the fixed 11-kernel mapping iteration) follows the oracle's loss curve; step_host on the packed
    host batch of mipsfusion.py:289-322 gives the same losses as step on rays generated by the oracle's formula."""
    from mipsfusion_b200.mapper import FusedMapper
    cfg = H.make_config(12, n_samples_d=32, n_range_d=11)
    of = H.oracle_field(cfg, seed=7)
    R, S = 256, 43
    rays7, pose_idx, poses, g = H.synth_batch_packed(R, seed=11)
    rays_o, rays_d, rgb, d, _ = H.synth_batch(R, S, seed=11)
    m_dev = FusedMapper(H.cuda_model(cfg, H.state_of(of)))
    opt_o = oadam.make_optimizer(of)
    dargs = [t.cuda().contiguous() for t in (rays_o, rays_d, rgb, d)]
    lo, ld = [], []
    for it in range(12):
        u = torch.rand(R, S, generator=g)
        opt_o.zero_grad()
        ret_o = of.forward(rays_o, rays_d, rgb, d, u)
        loss_o = of.total_loss(ret_o); loss_o.backward(); opt_o.step()
        lo.append([float(ret_o[k]) for k in ("rgb_loss", "depth_loss", "sdf_loss", "fs_loss")])
        ld.append(m_dev.step(*dargs, u=u.cuda())[:4].cpu().numpy().copy())
    np.testing.assert_allclose(np.array(ld), np.array(lo), rtol=5e-3)       # fp32 tolerance: losses 1e-3, Adam drift on top
    # host route: same kernels behind ray generation; jitter drawn on the device, so compare without perturbation
    import copy
    cfg0 = copy.deepcopy(cfg); cfg0["training"]["perturb"] = 0
    m_dev2 = FusedMapper(H.cuda_model(cfg0, H.state_of(of)))
    m_host2 = FusedMapper(H.cuda_model(cfg0, H.state_of(of)))
    a = m_dev2.step(*dargs).cpu().numpy().copy()
    b = m_host2.step_host(rays7.pin_memory(), pose_idx.pin_memory(), poses.cuda()).numpy().copy()
    np.testing.assert_array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("N,pattern", [(1, "all"), (1025, "all"), (1500, "none"), (2500, "mixed"), (128, "last")])
def test_backward_active_point_list_edge_cases(N, pattern, impl):
    """The backward visits only the rows of the upstream gradient that are non-zero.  Edge cases of that work list
    (empty list, full list, block / tile boundaries, a single active row) against the oracle, both decoders;
    skipped points get an exactly zero dL/dp and all-zero upstream gradients give exactly zero parameter gradients."""
    from mipsfusion_b200 import _lib as L
    cfg = H.make_config(12)
    of = H.oracle_field(cfg, grid_scale=0.3, seed=5)
    model = H.cuda_model(cfg, H.state_of(of))
    g = torch.Generator().manual_seed(N)
    bb = torch.tensor(cfg["mapping"]["bound"], dtype=torch.float32)
    pts = bb[:, 0] + (bb[:, 1] - bb[:, 0]) * torch.rand(N, 3, generator=g)
    G = torch.randn(N, 10, generator=g)
    keep = torch.ones(N, dtype=torch.bool)
    if pattern == "none":
        keep[:] = False
    elif pattern == "mixed":
        keep = torch.rand(N, generator=g) < 0.4
        keep[1024:2048] = False                                        # one compaction block without any active row
    elif pattern == "last":
        keep[:] = False; keep[-1] = True
    G = G * keep[:, None]
    po = pts.clone().requires_grad_(True)
    of.run_network(po).backward(G)
    # impl 1: fp32 decoder, points require grad (dL/dp checked); impl 0: tcgen05 decoder, parameter gradients only
    # (points with requires_grad always route through the fp32 decoder)
    want_dp = impl == 1
    L.call("mf_set_decoder_impl", impl)
    try:
        pc = pts.cuda().requires_grad_(want_dp)
        model.run_network(pc).backward(G.cuda())
        torch.cuda.synchronize()
    finally:
        L.call("mf_set_decoder_impl", 0)
    gg = model.embed_fn.params.grad.cpu()
    gp = pc.grad.cpu() if want_dp else None
    if want_dp:
        assert bool((gp[~keep] == 0).all())
    if pattern == "none":
        assert float(gg.abs().max()) == 0.0 and (gp is None or float(gp.abs().max()) == 0.0)
        assert all(float(p.grad.abs().max()) == 0.0 for p in model.decoder.parameters())
        return
    tol = 1e-5 if impl == 1 else 1e-4              # fp32 decoder (measured <= 5e-7) / tcgen05 decoder, bf16 x 3 (measured <= 1.5e-5)
    errs = {"grid": H.rel_err(gg, of.grid.grad)}
    errs.update({name: H.rel_err(p.grad.cpu(), of.w[name].grad) for name, p in model.decoder.named_parameters()})
    print(f"\n  N={N} {pattern} impl={impl}: max rel err {max(errs.values()):.2e} ({max(errs, key=errs.get)})")
    for name, e in errs.items():
        assert e < tol, (name, e)
    if want_dp:
        assert H.rel_err(gp, po.grad) < 1e-3


@pytest.mark.gpu
def test_one_launch_optimizer_step_equals_per_tensor_path_and_flat_storage_semantics():
    """create_map_optimizer re-seats the decoder on one flat buffer and steps grid + decoder in one launch; the result must
    equal the generic path (one launch per tensor) bit for bit, and the module keeps behaving like the reference's: same
    state_dict, deepcopy gives an independent module, load_state_dict is seen by the next forward."""
    import copy
    import mipsfusion_b200 as mf
    cfg = H.make_config(12, n_samples_d=32, n_range_d=11)
    cfg["training"]["perturb"] = 0
    of = H.oracle_field(cfg, seed=3)
    R, S = 200, 43
    rays_o, rays_d, rgb, d, _ = H.synth_batch(R, S, seed=5)
    args = [t.cuda() for t in (rays_o, rays_d, rgb, d)]
    t = cfg["training"]

    def run(fast):
        model = H.cuda_model(cfg, H.state_of(of))
        sd0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        opt = mf.create_map_optimizer(model, 1e-2, 1e-2)
        assert model.decoder.flat_storage() is not None
        for k, v in model.state_dict().items():                               # flattening changed no value, key or shape
            assert torch.equal(v.cpu(), sd0[k]), k
        if not fast:
            opt.__dict__["_map_model"] = None                                 # generic path: one launch per parameter tensor
        first, losses = None, []
        for it in range(3):
            ret = model(*args)
            loss = t["rgb_weight"] * ret["rgb_loss"] + t["sdf_weight"] * ret["sdf_loss"] + t["fs_weight"] * ret["fs_loss"]
            loss.backward()
            opt.step(zero_grad=True)
            losses.append(float(loss.detach()))
            if it == 0:
                first = {k: v.detach().clone() for k, v in model.state_dict().items()}
        assert (opt.__dict__.get("_flat_state") is not None) == fast
        return model, opt, first, losses

    m_fast, o_fast, first_fast, l_fast = run(True)
    m_gen, _, first_gen, l_gen = run(False)
    # after ONE step from identical weights: the decoder's gradients are reduced in a fixed order, so its update is bit-identical
    # in the two paths; the grid's gradient is summed with float atomics (order varies run to run), so a few of its entries
    # whose gradient cancels to ~0 may take the opposite +-lr step
    for k in first_fast:
        if k.startswith("decoder."):
            assert torch.equal(first_fast[k], first_gen[k]), k
    dg = (first_fast["embed_fn.params"] - first_gen["embed_fn.params"]).abs()
    assert float((dg > 1e-6).float().mean()) < 1e-3, float((dg > 1e-6).float().mean())
    np.testing.assert_allclose(l_fast, l_gen, rtol=1e-4)
    assert all(float(p.grad.abs().max()) == 0.0 for p in m_fast.parameters() if p.grad is not None)      # zero_grad folded in
    # torch-visible optimiser state exists per parameter (views of the flat moments)
    p0 = m_fast.decoder.pts_linear[0].weight
    assert o_fast.state[p0]["exp_avg"].shape == p0.shape and o_fast.state[p0]["step"] == 3
    # deepcopy: independent, ordinary parameters; both modules keep evaluating their own weights
    pts = (torch.rand(64, 3) * torch.tensor([3.5, 6.5, 4.2]) + torch.tensor([-0.6, 0.5, -1.15])).cuda()
    m_copy = copy.deepcopy(m_fast)
    with torch.no_grad():
        out_a = m_fast.run_network(pts).clone()
        assert torch.equal(m_copy.run_network(pts), out_a)
        m_copy.decoder.pts_linear[0].weight.add_(0.01)
        assert torch.equal(m_fast.run_network(pts), out_a)
        assert not torch.equal(m_copy.run_network(pts), out_a)
    # load_state_dict into the flat-backed module is picked up (the weight image is rebuilt)
    m_fast.load_state_dict(H.state_of(of))
    m_ref = H.cuda_model(cfg, H.state_of(of))
    with torch.no_grad():
        assert torch.equal(m_fast.run_network(pts), m_ref.run_network(pts))


@pytest.mark.gpu
def test_forward_on_batch_slices_equals_forward_on_contiguous_copies():
    """The reference's loop slices one (R,10) device tensor into rays_o / rays_d / target_rgb / target_d (mipsfusion.py:316-322);
    the targets are then read in place through their row stride (mf_sample_z_ld, mf_render_loss_fwd_ld,
    mf_render_loss_bwd_scalars) -- nothing may change."""
    cfg = H.make_config(12, n_samples_d=32, n_range_d=11)
    of = H.oracle_field(cfg, grid_scale=0.3)
    R, S = 333, 43
    rays_o, rays_d, rgb, d, u = H.synth_batch(R, S, seed=5)
    batch = torch.cat([rays_o, rays_d, rgb, d.reshape(R, 1)], -1).cuda()
    outs = []
    for sliced in (True, False):
        model = H.cuda_model(cfg, H.state_of(of))
        a = [batch[:, 0:3], batch[:, 3:6], batch[:, 6:9], batch[:, 9:10]]
        if not sliced:
            a = [t.contiguous() for t in a]
        ret = model(*a, u=u.cuda())
        loss = ret["rgb_loss"] + 1000.0 * ret["sdf_loss"] + 10.0 * ret["fs_loss"]
        loss.backward()
        outs.append((float(loss), ret["rgb"].detach().clone(), ret["depth"].detach().clone(), model.embed_fn.params.grad.detach().clone()))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    # (grid gradients are float atomics: equal up to summation order)
    assert H.rel_err(outs[0][3].cpu().numpy(), outs[1][3].cpu().numpy()) < 1e-5
    # evaluation route with a strided depth view
    model.eval()
    with torch.no_grad():
        r1 = model.render_rays(batch[:, 0:3], batch[:, 3:6], batch[:, 9:10], u=u.cuda())
        r2 = model.render_rays(batch[:, 0:3].contiguous(), batch[:, 3:6].contiguous(), batch[:, 9:10].contiguous(), u=u.cuda())
    assert torch.equal(r1["rgb"], r2["rgb"]) and torch.equal(r1["depth"], r2["depth"])
