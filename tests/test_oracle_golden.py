"""Pins the oracle against vectors produced by the reference's own Python
(tests/golden/make_golden.py, run where /root/reference exists)."""
import numpy as np
import torch

from oracle import sampling as osamp, scene as oscene, ro as oro, joint_query as ojq, adam as oadam
from oracle.decoder import mlp_reg


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_uniform_lattice_known_answers(golden):
    fx = golden("lattice")
    k = 0
    while f"args{k}" in fx:
        r, c = osamp.sample_pixels_uniformly(*[int(v) for v in fx[f"args{k}"]])
        assert np.array_equal(r.numpy(), fx[f"rows{k}"]) and np.array_equal(c.numpy(), fx[f"cols{k}"])
        k += 1
    # SURVEY.md section 4 known-answer table
    r, c = osamp.sample_pixels_uniformly(460, 620, 16, 24)
    assert (int(r[0]), int(r[-1]), int(c[0]), int(c[1]), int(c[2]), int(c[-1]), r.numel()) == (27, 432, 33, 57, 81, 585, 384)
    r, c = osamp.sample_pixels_uniformly(460, 620, 150, 200)
    assert (int(r[0]), int(r[-1]), int(c[0]), int(c[-1]), r.numel()) == (6, 453, 11, 608, 30000)


def test_random_samplers(golden):
    fx = golden("sampling")
    depth, keys = T(fx["depth"]), T(fx["keys"])
    assert np.array_equal(osamp.sample_valid_pixels_random(depth, 200, keys).numpy(), fx["valid_random_200"])
    r, c = osamp.sample_pixels_mix(60, 80, 6, 8, depth, 300, keys)
    assert np.array_equal(r.numpy(), fx["mix_rows"]) and np.array_equal(c.numpy(), fx["mix_cols"])
    r, c = osamp.pixel_indices_to_rc(T(fx["rc_idx"]), 60, 80)
    assert np.array_equal(r.numpy(), fx["rc_rows"]) and np.array_equal(c.numpy(), fx["rc_cols"])


def test_losses(golden):
    fx = golden("losses")
    z, d, sdf, prob = (T(fx[k]) for k in ("z", "d", "sdf", "prob"))
    for tag, emd in (("emd", 0.01), ("noemd", 0.0)):
        fs, sl, counts = oscene.get_sdf_loss(z, d, sdf, prob, 0.1, 5, emd)
        assert np.array_equal(fs.numpy(), fx[f"fs_{tag}"]) and np.array_equal(sl.numpy(), fx[f"sdf_{tag}"])
        assert list(counts) == list(fx["counts"])
    fm, sm, fw, sw, _ = oscene.get_masks(z, d, 0.1)
    assert np.array_equal(fm.numpy(), fx["front_mask"]) and np.array_equal(sm.numpy(), fx["sdf_mask"])
    assert np.float32(fw) == fx["fs_weight"] and np.float32(sw) == fx["sdf_weight"]


def test_decoder(golden):
    fx = golden("decoder")
    w = {k[2:]: T(v) for k, v in fx.items() if k.startswith("w:")}
    out = mlp_reg(w, T(fx["embed"]), T(fx["embed_pos"]), T(fx["pts"]))
    np.testing.assert_allclose(out.numpy(), fx["out"], rtol=0, atol=1e-6)


def field_from_fixture(fx):
    cfg = oscene.default_config()
    cfg["grid"]["hash_size"] = int(fx["hash_size"])
    f = oscene.OracleField(cfg)
    f.grid = T(fx["w:embed_fn.params"]).clone().requires_grad_(True)
    f.w = {k[len("w:decoder."):]: T(v).clone().requires_grad_(True) for k, v in fx.items() if k.startswith("w:decoder.")}
    return f


def test_scene_forward_backward(golden):
    fx = golden("scene")
    f = field_from_fixture(fx)
    rays = T(fx["rays"])
    ro = T(fx["rays_o"]).clone().requires_grad_(True)
    rd = T(fx["rays_d"]).clone().requires_grad_(True)
    ret = f.forward(ro, rd, rays[:, 3:6], rays[:, 6:7], T(fx["u"]))
    np.testing.assert_array_equal(ret["z_vals"].detach().numpy(), fx["z_vals"])
    np.testing.assert_allclose(ret["raw"].detach().numpy(), fx["raw"], rtol=0, atol=2e-6)
    for k in ("rgb", "depth", "rgb_loss", "depth_loss", "sdf_loss", "fs_loss", "psnr"):
        np.testing.assert_allclose(ret[k].detach().numpy(), fx[k], rtol=2e-6, atol=1e-7, err_msg=k)
    loss = f.total_loss(ret)
    np.testing.assert_allclose(loss.detach().numpy(), fx["loss"], rtol=2e-6)
    loss.backward()
    np.testing.assert_allclose(f.grid.grad.numpy(), fx["g:embed_fn.params"], rtol=1e-4, atol=1e-7)
    for k, p in f.w.items():
        np.testing.assert_allclose(p.grad.numpy(), fx["g:decoder." + k], rtol=1e-4, atol=1e-6, err_msg=k)
    np.testing.assert_allclose(ro.grad.numpy(), fx["g_rays_o"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(rd.grad.numpy(), fx["g_rays_d"], rtol=1e-4, atol=1e-6)
    with torch.no_grad():
        q = f.run_network(T(fx["q_pts"]))
    np.testing.assert_allclose(q.numpy(), fx["q_out"], rtol=0, atol=2e-6)
    rend = f.render_rays(ro.detach(), rd.detach(), rays[:, 6:7], T(fx["u"]))
    for k in ("depth_var", "acc_map", "disp_map"):
        np.testing.assert_allclose(rend[k].detach().numpy(), fx[k], rtol=1e-5, atol=1e-7, err_msg=k)


def test_random_optimizer(golden):
    fx, fs = golden("ro"), golden("scene")
    f = field_from_fixture(fs)
    rows, cols = T(fx["rows"]), T(fx["cols"])
    depth, dirs = T(fx["depth"]), T(fx["dirs"])
    init = T(fx["init"])
    with torch.no_grad():
        pst7 = oro.pose_6D_to_7D(T(fx["particles"]) * 0.02)
        R_, t_ = oro.get_abs_pose(init[:3, :3], init[:3, 3:], pst7)
        np.testing.assert_allclose(R_.numpy(), fx["abs_rot"], atol=1e-7)
        fit, msdf = oro.get_fitness(f, R_, t_, depth[rows, cols].unsqueeze(-1), dirs[rows, cols, :], 0.1)
    np.testing.assert_allclose(fit.numpy(), fx["fitness"], rtol=1e-5)
    np.testing.assert_allclose(msdf.numpy(), fx["mean_sdf"], rtol=1e-5)
    pose, infos = oro.optimize(f, depth, dirs, rows, cols, init, T(fx["particles"]), 3, 0.1)
    np.testing.assert_allclose(pose.numpy(), fx["pose"], atol=1e-6)
    assert all(0 <= i["argmin"] < fit.numel() for i in infos)


def test_blend_weights(golden):
    fx = golden("blend")
    W = ojq.compute_weights(fx["entropy"], fx["dist_w"], fx["mask"])
    np.testing.assert_allclose(W, fx["weights"], rtol=1e-6, atol=0)
    d = np.abs(fx["dist"])
    np.testing.assert_allclose(ojq.pdf_gauss(d, 0.0, d.max() / 3.0), fx["dist_weight"], rtol=1e-6)


def test_adam_matches_torch():
    g = torch.Generator().manual_seed(2)
    for eps, wd in ((1e-15, 0.0), (1e-8, 1e-6)):
        p = torch.randn(1000, generator=g).requires_grad_(True)
        opt = torch.optim.Adam([{"params": [p], "eps": eps, "weight_decay": wd, "lr": 1e-2}], betas=(0.9, 0.99))
        q, m, v = p.detach().clone(), torch.zeros(1000), torch.zeros(1000)
        for step in range(1, 5):
            grad = torch.randn(1000, generator=g) * (step % 2)      # includes an all-zero gradient step
            p.grad = grad.clone()
            opt.step()
            oadam.adam_step(q, grad, m, v, step, 1e-2, 0.9, 0.99, eps, wd)
            np.testing.assert_allclose(q.numpy(), p.detach().numpy(), rtol=1e-6, atol=1e-7)
