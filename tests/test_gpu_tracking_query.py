"""GPU parity: pixel samplers (bit-exact), RandomOptimizer scoring/update, joint multi-submap query."""
import types

import numpy as np
import pytest
import torch

import helpers as H
from oracle import joint_query as ojq
from oracle import ro as oro
from oracle import sampling as osamp

pytestmark = pytest.mark.gpu


def test_uniform_lattice_golden(golden):
    from mipsfusion_b200 import sampling_helper as sh
    fx = golden("lattice")
    k = 0
    while f"args{k}" in fx:
        r, c = sh.sample_pixels_uniformly(*[int(v) for v in fx[f"args{k}"]])
        assert np.array_equal(r.cpu().numpy(), fx[f"rows{k}"]) and np.array_equal(c.cpu().numpy(), fx[f"cols{k}"])
        k += 1


def test_random_samplers_golden_and_full_size(golden):
    from mipsfusion_b200 import sampling_helper as sh
    fx = golden("sampling")                                    # produced by the reference's own samplers
    depth, keys = H.T(fx["depth"]).cuda(), H.T(fx["keys"]).cuda()
    assert np.array_equal(sh.sample_valid_pixels_random(depth, 200, keys).cpu().numpy(), fx["valid_random_200"])
    r, c = sh.sample_pixels_mix(60, 80, 6, 8, depth, 300, keys)
    assert np.array_equal(r.cpu().numpy(), fx["mix_rows"]) and np.array_equal(c.cpu().numpy(), fx["mix_cols"])
    # working resolution 460x620, with duplicate keys and more requested pixels than valid ones (tie rule)
    g = torch.Generator().manual_seed(3)
    Hh, Ww = 460, 620
    d = torch.rand(Hh, Ww, generator=g)
    d[torch.rand(Hh, Ww, generator=g) < 0.3] = 0.0
    k = torch.randn(Hh * Ww, generator=g).abs()
    k[1000:1200] = k[5]                                        # exact duplicates
    for num in (1000, 2048, 4096 + 384):
        r, c = sh.sample_pixels_mix(Hh, Ww, 16, 24, d.cuda(), num, k.cuda())
        ro_, co_ = osamp.sample_pixels_mix(Hh, Ww, 16, 24, d, num, k)
        assert np.array_equal(r.cpu().numpy(), ro_.numpy()) and np.array_equal(c.cpu().numpy(), co_.numpy()), num
    d2 = torch.zeros(40, 50); d2[3, 4:30] = 1.0                # only 26 valid pixels, 100 requested
    k2 = torch.randn(2000, generator=g).abs()
    assert np.array_equal(sh.sample_valid_pixels_random(d2.cuda(), 100, k2.cuda()).cpu().numpy(),
                          osamp.sample_valid_pixels_random(d2, 100, k2).numpy())


def test_gen_rays():
    import ctypes as C
    from mipsfusion_b200 import _lib as L
    g = torch.Generator().manual_seed(0)
    R, K = 1000, 5
    dirs = torch.randn(R, 3, generator=g); poses = torch.randn(K, 4, 4, generator=g)
    idx = torch.randint(-1, K, (R,), generator=g)
    rd_o, ro_o = osamp.rays_camera_to_world2(dirs, poses, idx)
    ro = torch.empty(R, 3, device="cuda"); rd = torch.empty(R, 3, device="cuda")
    dirs_c, poses_c, idx_c = dirs.cuda(), poses.cuda(), idx.cuda()          # keep the device copies alive across the calls
    L.call("mf_gen_rays", L.ptr(dirs_c), L.ptr(poses_c), L.ptr(idx_c), L.ptr(ro), L.ptr(rd), R, K, L.stream())
    assert np.array_equal(ro.cpu().numpy(), ro_o.numpy())
    np.testing.assert_allclose(rd.cpu().numpy(), rd_o.numpy(), rtol=1e-6, atol=1e-6)
    po = poses.clone().requires_grad_(True)
    rd2, ro2 = osamp.rays_camera_to_world2(dirs, po, idx)
    go, gd = torch.randn(R, 3, generator=g), torch.randn(R, 3, generator=g)
    (ro2 * go).sum().add((rd2 * gd).sum()).backward()
    dp = torch.zeros(K, 4, 4, device="cuda")
    go_c, gd_c = go.cuda(), gd.cuda()
    L.call("mf_gen_rays_bwd", L.ptr(dirs_c), L.ptr(idx_c), L.ptr(go_c), L.ptr(gd_c), L.ptr(dp), R, K, L.stream())
    assert H.rel_err(dp.cpu(), po.grad) < 1e-4


def _ro_setup(fx, model):
    import mipsfusion_b200 as mf
    cfg = dict(model.config)
    cfg["tracking"] = {"RO": {"particle_size": int(fx["particles"].shape[0]), "initial_scaling_factor": 0.02, "rescaling_factor": 0.5,
                              "n_rows": 6, "n_cols": 8}, "ignore_edge_W": 2, "ignore_edge_H": 2}
    ds = types.SimpleNamespace(H=60, W=80, fx=40.0, fy=40.0, cx=39.5, cy=29.5, rays_d=H.T(fx["dirs"]))
    slam = types.SimpleNamespace(dataset=ds, device="cuda")
    return mf.RandomOptimizer(cfg, slam, particles=H.T(fx["particles"]))


def test_random_optimizer_golden(golden):
    fx, fs = golden("ro"), golden("scene")
    model = H.cuda_model(H.make_config(int(fs["hash_size"])), H.fixture_state(fs), train=False)
    ro = _ro_setup(fx, model)
    assert np.array_equal(ro.row_indices.cpu().numpy(), fx["rows"]) and np.array_equal(ro.col_indices.cpu().numpy(), fx["cols"])
    init = H.T(fx["init"])
    depth, dirs = H.T(fx["depth"]), H.T(fx["dirs"])
    rows, cols = H.T(fx["rows"]), H.T(fx["cols"])
    # one scoring pass vs the reference's get_fitness
    search = torch.full((6,), 0.02, device="cuda")
    fit, msdf, pst7 = ro.score(model, init[:3, :3].contiguous().cuda(), init[:3, 3].contiguous().cuda(), search,
                               depth[rows, cols].contiguous().cuda(), dirs[rows, cols, :].contiguous().cuda())
    np.testing.assert_allclose(fit.cpu().numpy(), fx["fitness"], rtol=1e-3)
    np.testing.assert_allclose(msdf.cpu().numpy(), fx["mean_sdf"], rtol=1e-3)
    # integer decisions on this pass: better mask / count / argmin (bit exact where the oracle's margin allows)
    f_ref = fx["fitness"]
    margin = np.abs(f_ref - f_ref[0]) > 1e-3 * np.abs(f_ref[0])
    better = (fit < fit[0]).cpu().numpy()
    assert np.array_equal(better[margin], (f_ref < f_ref[0])[margin])
    assert int(torch.argmin(fit)) == int(np.argmin(f_ref))
    # the full 3-iteration optimisation vs the reference's RandomOptimizer.optimize
    pose = ro.optimize(model, depth, init.clone(), H.T(fx["c2w"]), n_iter=3)
    assert not pose.is_cuda and pose.shape == (4, 4)
    np.testing.assert_allclose(pose.numpy(), fx["pose"], atol=1e-4)             # 1e-4 on the pose (north_star)
    info = ro.last_info.cpu().numpy()
    of = H.oracle_field(H.make_config(int(fs["hash_size"])), H.fixture_state(fs))
    _, infos = oro.optimize(of, depth, dirs, rows, cols, init, H.T(fx["particles"]), 3, 0.1)
    assert [int(i[1]) for i in info] == [int(x["success"]) for x in infos]
    assert [int(i[2]) for i in info] == [x["argmin"] for x in infos]
    for i, x in zip(info, infos):
        assert abs(int(i[0]) - x["count"]) <= 1                                  # near-ties against fitness[0]


def test_ro_baseline_shape_shards_agree():
    """BASELINE shape (1024 candidates x 2048 pixels): scoring a candidate shard equals the same rows of
    the full run (the multi-GPU partition), and row 0 (the zero particle) reproduces the current pose."""
    import ctypes as C
    import mipsfusion_b200 as mf
    from mipsfusion_b200 import _lib as L, synth
    cfg = H.make_config(19)
    of = H.oracle_field(cfg, grid_scale=0.3, seed=3)
    model = H.cuda_model(cfg, H.state_of(of), train=False)
    Cn, P = 1024, 2048
    g = torch.Generator().manual_seed(0)
    particles = torch.randn(Cn, 6, generator=g).clamp(-2, 2); particles[0] = 0
    c2w = synth.trajectory(4)[1]
    dirs = synth.camera_rays()
    rows, cols = osamp.sample_pixels_uniformly(460, 620, 32, 64)
    frame = synth.render_frame(c2w, dirs[rows, cols][None].contiguous())
    target_d = frame["depth"].reshape(-1).cuda(); rays_d = dirs[rows, cols].contiguous().cuda()
    rot, trans = c2w[:3, :3].contiguous().cuda(), c2w[:3, 3].contiguous().cuda()
    search = torch.full((6,), 0.02, device="cuda")
    particles_c = particles.cuda()
    field = model._field()

    def score(b, n):
        fit = torch.empty(n, device="cuda"); ms = torch.empty(n, device="cuda"); p7 = torch.empty(n, 7, device="cuda")
        scratch = torch.empty(n * (P + 12), device="cuda")
        L.call("mf_ro_score", L.ptr(particles_c), L.ptr(search), L.ptr(rot), L.ptr(trans), L.ptr(rays_d), L.ptr(target_d),
               C.byref(field), 0.1, 1000.0, b, n, P, L.ptr(fit), L.ptr(ms), L.ptr(p7), L.ptr(scratch), L.stream())
        return fit, ms, p7
    full = score(0, Cn)
    part = score(256, 128)
    assert torch.equal(full[0][256:384], part[0]) and torch.equal(full[2][256:384], part[2])
    # candidate 0 against the oracle (a 2048-point query)
    with torch.no_grad():
        pst7 = oro.pose_6D_to_7D(particles[:4] * 0.02)
        R_, t_ = oro.get_abs_pose(c2w[:3, :3], c2w[:3, 3:], pst7)
        fit_o, _ = oro.get_fitness(of, R_, t_, frame["depth"].reshape(-1, 1), dirs[rows, cols], 0.1)
    np.testing.assert_allclose(full[0][:4].cpu().numpy(), fit_o.numpy(), rtol=1e-3)


def _submaps(n, cfg):
    from mipsfusion_b200 import synth
    fields, models, poses, amin, amax, cents = [], [], [], [], [], []
    traj = synth.trajectory(8)
    for i in range(n):
        of = H.oracle_field(cfg, grid_scale=0.4, seed=20 + i)
        fields.append(of); models.append(H.cuda_model(cfg, H.state_of(of), train=False))
        T_ = torch.eye(4); T_[:3, :3] = traj[i][:3, :3]; T_[:3, 3] = torch.tensor([0.3 * i, 0.2 * i, 0.0])
        poses.append(T_)
        lo = np.array([-0.2 + 0.5 * i, 1.0 + 0.7 * i, -0.6]); hi = lo + np.array([1.6, 2.4, 2.2])
        amin.append(lo); amax.append(hi); cents.append(((lo + hi) / 2 + 0.1).astype(np.float32))
    return fields, models, poses, amin, amax, cents


def test_joint_query_vs_oracle():
    import mipsfusion_b200 as mf
    cfg = H.make_config(12)
    cfg["grid"]["use_bound_normalize"] = False            # submap-local frames: normalise by localMLP_max_len
    fields, models, poses, amin, amax, cents = _submaps(3, cfg)
    axes = mf.get_grid_uniform(np.array([-0.2, 1.0, -0.6]), np.array([2.4, 4.8, 1.6]), voxel_size=0.11)
    jq = mf.JointSubmapQuery(models, poses, amin, amax, cents)
    xx, yy, zz = np.meshgrid(*axes)
    pts64 = np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T
    G = pts64.shape[0]
    rng = np.random.RandomState(0)
    vis = rng.rand(G, 3) < 0.8
    # the oracle tests containment on the float64 points and feeds the network their float32 rounding, as the reference does
    res_o = ojq.joint_query(pts64, fields, poses, amin, amax, cents, vis_masks=vis)
    p64 = pts64
    contain64 = np.stack([np.all((p64 >= amin[i]) & (p64 <= amax[i]), -1) for i in range(3)], -1)
    res = jq.query(axes=axes, vis=vis, want_contain=True)
    assert np.array_equal(res["contain"].cpu().numpy(), contain64)                  # submap assignment: bit exact
    assert np.array_equal(res["mask"].cpu().numpy(), (contain64 & vis).any(-1))
    assert np.array_equal(res_o["contain"], contain64)
    err = H.rel_err(res["sdf"].cpu().numpy(), res_o["sdf"])
    assert err < 1e-3, err
    # explicit points (vertex colour pass)
    sel = rng.choice(G, 500, replace=False)
    res_c = jq.query(points=pts64[sel], vis=vis[sel], color=True)
    res_co = ojq.joint_query(pts64[sel], fields, poses, amin, amax, cents, vis_masks=vis[sel], color=True)
    err_c = H.rel_err(res_c["rgb"].cpu().numpy(), res_co["rgb"])
    assert err_c < 1e-3, err_c
    # no submap sees the point -> -1 / masked
    far = jq.query(points=np.array([[50.0, 50.0, 50.0]]))
    assert float(far["sdf"][0]) == -1.0 and not bool(far["mask"][0])


def test_render_full_img_vs_oracle():
    """Logger.render_full_img (Logger.py:193-214): whole-image rendering on the device vs the oracle's render_rays on the same
    rays (a 46 x 62 image keeps the oracle in seconds), one call and batched."""
    import mipsfusion_b200 as mf
    from mipsfusion_b200 import synth
    cfg = H.make_config(14)
    cfg["training"]["perturb"] = 0
    of = H.oracle_field(cfg, grid_scale=0.3, seed=9)
    model = H.cuda_model(cfg, H.state_of(of), train=False)
    dirs = synth.camera_rays()[::10, ::10].contiguous()                 # (46, 62, 3)
    c2w = synth.trajectory(4)[2]
    frame = synth.render_frame(c2w, dirs)
    rgb, depth = mf.render_full_img(model, dirs, c2w, frame["depth"])
    rgb_b, depth_b = mf.render_full_img(model, dirs, c2w, frame["depth"], ray_batch_size=1000)
    assert rgb.shape == (46, 62, 3) and depth.shape == (46, 62)
    assert torch.equal(rgb, rgb_b) and torch.equal(depth, depth_b)
    rays_d = torch.sum(dirs.reshape(-1, 1, 3) * c2w[None, :3, :3], -1)
    rays_o = c2w[None, :3, 3].repeat(rays_d.shape[0], 1)
    with torch.no_grad():
        ret = of.render_rays(rays_o, rays_d, target_d=frame["depth"].reshape(-1, 1))
    np.testing.assert_allclose(rgb.reshape(-1, 3).cpu().numpy(), ret["rgb"].numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(depth.reshape(-1).cpu().numpy(), ret["depth"].numpy(), rtol=1e-4, atol=1e-4)
