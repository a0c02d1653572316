"""Generates tests/golden/*.npz by running the REFERENCE's own Python, imported
unchanged from /root/reference, on seeded inputs (CPU).

Run in the build container only (the reference checkout does not travel):
    python tests/golden/make_golden.py
What is the reference's own arithmetic (decoder, rendering, losses, samplers,
RandomOptimizer, Mesher weight blend) is pinned by these vectors.  The tcnn
encodings are not available offline, so inside JointEncoding they are supplied
by oracle/shims/tinycudann (= the oracle): fixtures that depend on them pin the
reference's *composition* of the encodings, not tcnn itself (parity unpinned).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("MIPSFUSION_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, REF)
OUT = os.path.dirname(os.path.abspath(__file__))

np.bool = bool                                                   # numpy>=1.24 dropped the alias the reference uses
torch.Tensor.cuda = lambda self, *a, **k: self                   # model/decoder.py:29 on a CPU box

from helper_functions import sampling_helper as ref_samp        # noqa: E402
from helper_functions import utils as ref_utils                 # noqa: E402
from model.decoder import MLP_reg                                # noqa: E402
from model.scene_rep import JointEncoding                        # noqa: E402
from RandomOptimizer import RandomOptimizer                      # noqa: E402
from vis import math_helper as ref_mh                            # noqa: E402

from oracle import scene as oscene                               # noqa: E402


def small_config(hash_size=10):
    cfg = oscene.default_config()
    cfg["grid"]["hash_size"] = hash_size
    cfg["tracking"] = {"RO": {"particle_size": 48, "initial_scaling_factor": 0.02, "rescaling_factor": 0.5,
                              "n_rows": 6, "n_cols": 8}, "ignore_edge_W": 2, "ignore_edge_H": 2}
    return cfg


def gen_lattice():
    out = {}
    for k, a in enumerate([(460, 620, 16, 24), (460, 620, 150, 200), (460, 620, 24, 32), (460, 620, 15, 20),
                           (480, 640, 16, 24), (60, 80, 6, 8), (7, 9, 7, 9)]):
        r, c = ref_samp.sample_pixels_uniformly(*a)
        out[f"args{k}"] = np.asarray(a); out[f"rows{k}"] = r.numpy(); out[f"cols{k}"] = c.numpy()
    np.savez_compressed(os.path.join(OUT, "lattice.npz"), **out)


def gen_sampling():
    g = torch.Generator().manual_seed(11)
    H, W = 60, 80
    depth = torch.rand(H, W, generator=g) * 3
    depth[torch.rand(H, W, generator=g) < 0.1] = 0.0
    keys_n = torch.randn(H * W, generator=g)                     # the randn draw; reference takes abs()
    out = {"depth": depth.numpy(), "keys": keys_n.abs().numpy()}
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: keys_n.reshape(t.shape).clone()
    try:
        out["valid_random_200"] = ref_samp.sample_valid_pixels_random(depth, 200).numpy()
        r, c = ref_samp.sample_pixels_mix(H, W, 6, 8, depth, 300)
        out["mix_rows"], out["mix_cols"] = r.numpy(), c.numpy()
    finally:
        torch.randn_like = orig
    idx = torch.arange(0, H * W, 7)
    r, c = ref_samp.pixel_indices_to_rc(idx, H, W)
    out["rc_idx"], out["rc_rows"], out["rc_cols"] = idx.numpy(), r.numpy(), c.numpy()
    np.savez_compressed(os.path.join(OUT, "sampling.npz"), **out)


def gen_losses():
    g = torch.Generator().manual_seed(5)
    R, S = 32, 75
    z = torch.sort(torch.rand(R, S, generator=g) * 5, -1)[0]
    d = torch.rand(R, 1, generator=g) * 4 + 0.3
    d[:4] = 0.0
    sdf = torch.rand(R, S, generator=g) * 2 - 1
    prob = torch.softmax(torch.randn(R, S, 5, generator=g), -1)
    out = {"z": z.numpy(), "d": d.numpy(), "sdf": sdf.numpy(), "prob": prob.numpy()}
    for tag, emd in (("emd", 0.01), ("noemd", 0.0)):
        fs, sl = ref_utils.get_sdf_loss(z, d, sdf, prob, 0.1, 5, emd, "l2")
        out[f"fs_{tag}"], out[f"sdf_{tag}"] = fs.numpy(), sl.numpy()
    fm, sm, fw, sw = ref_utils.get_masks(z, d, 0.1)
    out["front_mask"], out["sdf_mask"] = fm.numpy(), sm.numpy()
    out["fs_weight"], out["sdf_weight"] = np.float32(fw), np.float32(sw)
    out["counts"] = np.asarray([int(torch.count_nonzero(fm)), int(torch.count_nonzero(sm))])
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)


def gen_decoder():
    torch.manual_seed(0)
    dec = MLP_reg(None, input_ch=32, input_ch_pos=48)
    g = torch.Generator().manual_seed(3)
    N = 96
    embed = (torch.rand(N, 32, generator=g) * 2 - 1) * 0.3
    pts = torch.rand(N, 3, generator=g)
    embed_pos = torch.sin(torch.rand(N, 48, generator=g) * 6.0)
    out = dec(embed, embed_pos, pts)
    fx = {"embed": embed.numpy(), "embed_pos": embed_pos.numpy(), "pts": pts.numpy(), "out": out.detach().numpy()}
    for k, v in dec.state_dict().items():
        fx["w:" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "decoder.npz"), **fx)


def build_ref_model(cfg):
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"]))                 # fp64, mipsfusion.py:94
    nf = torch.from_numpy(np.array(cfg["mapping"]["localMLP_max_len"]))
    torch.manual_seed(0)
    model = JointEncoding(cfg, bb, nf)
    # make the MLP weights non-degenerate for the SDF head and give the grid visible amplitude
    g = torch.Generator().manual_seed(21)
    with torch.no_grad():
        model.embed_fn.params.copy_((torch.rand(model.embed_fn.params.shape, generator=g) * 2 - 1) * 0.5)
    return model


def gen_scene():
    from mipsfusion_b200 import synth
    cfg = small_config(10)
    model = build_ref_model(cfg)
    model.train()
    g = torch.Generator().manual_seed(9)
    c2w = synth.trajectory(4)[1]
    dirs = synth.camera_rays()
    frame = synth.render_frame(c2w, dirs[::23, ::31].contiguous())           # 20 x 20 pixel sub-lattice
    rays = synth.frame_rays(frame)
    R = 40
    sel = torch.randperm(rays.shape[0], generator=g)[:R]
    rays = rays[sel]
    rays[:3, 6] = 0.0                                                        # a few invalid-depth rays
    rays_d = torch.sum(rays[:, None, :3] * c2w[None, :3, :3], -1)
    rays_o = c2w[None, :3, 3].repeat(R, 1)
    rays_o.requires_grad_(True); rays_d.requires_grad_(True)
    S = cfg["training"]["n_samples_d"] + cfg["training"]["n_range_d"]
    u = torch.rand(R, S, generator=g)
    orig = torch.rand
    torch.rand = lambda *a, **k: u.clone()                                   # scene_rep.py:176
    try:
        ret = model.forward(rays_o, rays_d, rays[:, 3:6], rays[:, 6:7])
        rend = model.render_rays(rays_o.detach(), rays_d.detach(), target_d=rays[:, 6:7])
    finally:
        torch.rand = orig
    t = cfg["training"]
    loss = t["rgb_weight"] * ret["rgb_loss"] + t["sdf_weight"] * ret["sdf_loss"] + t["fs_weight"] * ret["fs_loss"]
    loss.backward()
    fx = {"c2w": c2w.numpy(), "rays": rays.numpy(), "rays_o": rays_o.detach().numpy(), "rays_d": rays_d.detach().numpy(),
          "u": u.numpy(), "loss": loss.detach().numpy(), "z_vals": rend["z_vals"].numpy(), "raw": rend["raw"].detach().numpy(),
          "rgb": ret["rgb"].detach().numpy(), "depth": ret["depth"].detach().numpy(),
          "depth_var": rend["depth_var"].detach().numpy(), "acc_map": rend["acc_map"].detach().numpy(),
          "disp_map": rend["disp_map"].detach().numpy(),
          "g_rays_o": rays_o.grad.numpy(), "g_rays_d": rays_d.grad.numpy(), "hash_size": np.asarray(10)}
    for k in ("rgb_loss", "depth_loss", "sdf_loss", "fs_loss", "psnr"):
        fx[k] = ret[k].detach().numpy()
    for k, v in model.state_dict().items():
        fx["w:" + k] = v.numpy()
    for k, p in model.named_parameters():
        if p.grad is not None:
            fx["g:" + k] = p.grad.numpy()
    # points-only query (run_network) on scattered points incl. outside the bound
    pts = (torch.rand(128, 3, generator=g) * torch.tensor([4.0, 7.0, 4.6]) + torch.tensor([-0.8, 0.3, -1.3]))
    with torch.no_grad():
        fx["q_pts"] = pts.numpy(); fx["q_out"] = model.run_network(pts).numpy()
    np.savez_compressed(os.path.join(OUT, "scene.npz"), **fx)
    return cfg, model


def gen_ro(cfg, model):
    H, W = 60, 80
    g = torch.Generator().manual_seed(13)
    fx_, cx_, cy_ = 40.0, 39.5, 29.5
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing="xy")
    dirs = torch.stack([(i - cx_) / fx_, -(j - cy_) / fx_, -torch.ones_like(i)], -1)
    from mipsfusion_b200 import synth
    c2w = synth.trajectory(4)[2]
    frame = synth.render_frame(c2w, dirs, invalid_frac=0.05, seed=3)
    depth = frame["depth"]
    ds = types.SimpleNamespace(H=H, W=W, fx=fx_, fy=fx_, cx=cx_, cy=cy_, rays_d=dirs)
    slam = types.SimpleNamespace(dataset=ds, device="cpu")
    np.random.seed(7)
    ro = RandomOptimizer(cfg, slam)
    model.eval()
    init = c2w.clone()
    init[:3, 3] += torch.tensor([0.02, -0.015, 0.01])
    pose = ro.optimize(model, depth, init.clone(), c2w.clone(), n_iter=3)
    # one scoring pass with the template at the initial search size (fitness is the kernel's output)
    pst7 = ro.pose_6D_to_7D(ro.pre_sampled_particle * 0.02)
    R_, t_ = ro.get_abs_pose(init[:3, :3], init[:3, 3:], pst7)
    td = depth[ro.row_indices, ro.col_indices].unsqueeze(-1)
    rd = dirs[ro.row_indices, ro.col_indices, :]
    with torch.no_grad():
        fit, msdf = ro.get_fitness(model, R_, t_, c2w, td, rd)
    np.savez_compressed(os.path.join(OUT, "ro.npz"), depth=depth.numpy(), dirs=dirs.numpy(), c2w=c2w.numpy(),
                        init=init.numpy(), particles=ro.pre_sampled_particle.numpy(), pose=pose.numpy(),
                        rows=ro.row_indices.numpy(), cols=ro.col_indices.numpy(), fitness=fit.numpy(),
                        mean_sdf=msdf.numpy(), abs_rot=R_.numpy(), abs_trans=t_.numpy())


def gen_blend():
    rng = np.random.RandomState(4)
    G, M = 257, 4
    ent = rng.rand(G, M).astype(np.float32) * 2
    dw = rng.rand(G, M).astype(np.float32)
    mask = rng.rand(G, M) < 0.5
    mask[:5] = False
    dist = rng.rand(300).astype(np.float32) * 3
    np.savez_compressed(os.path.join(OUT, "blend.npz"), entropy=ent, dist_w=dw, mask=mask,
                        weights=ref_mh.compute_weights(ent, dw, mask), dist=dist,
                        dist_weight=ref_mh.convert_dist_to_weight(dist))


def gen_overlap():
    """InactiveMap.get_SDF_dif / get_SDF_dif2 (the reference's own methods, called unbound on a stub) with two reference
    JointEncoding models; loss and its gradients w.r.t. the two first-keyframe poses."""
    import types as _types
    sys.modules.setdefault("Logger", _types.ModuleType("Logger"))          # Logger.py needs matplotlib; InactiveMap only imports the name
    if not hasattr(sys.modules["Logger"], "Logger"):
        sys.modules["Logger"].Logger = object
    import InactiveMap as ref_im
    from mipsfusion_b200 import synth
    cfg = small_config(10)
    models = []
    for seed in (21, 22):
        m = build_ref_model(cfg)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            m.embed_fn.params.copy_((torch.rand(m.embed_fn.params.shape, generator=g) * 2 - 1) * 0.5)
            for p_ in m.decoder.parameters():
                p_.add_(0.05 * torch.randn(p_.shape, generator=g))
        m.eval()
        models.append(m)
    g = torch.Generator().manual_seed(5)
    poses = synth.trajectory(6)
    dirs = synth.camera_rays()
    frame = synth.render_frame(poses[2], dirs[::19, ::27].contiguous())
    rays = synth.frame_rays(frame)
    N = 96
    rays = rays[torch.randperm(rays.shape[0], generator=g)[:N]].clone()
    rays[:5, 6] = 0.0
    kf_idx = torch.randint(0, 3, (N,), generator=g)
    ovlp = torch.stack([poses[2], poses[3], poses[4]])[kf_idx]                      # (N,4,4) world poses of the overlapping keyframes
    first1 = poses[0].clone().requires_grad_(True); first2 = poses[1].clone().requires_grad_(True)
    stub = _types.SimpleNamespace(device="cpu", trunc_value=cfg["training"]["trunc"], model_list=models)
    stub.infer_pts = lambda *a: ref_im.InactiveMap.infer_pts(stub, *a)
    loss = ref_im.InactiveMap.get_SDF_dif(stub, rays, ovlp, 0, 1, first1, first2)
    loss.backward()
    out = {"rays": rays.numpy(), "ovlp": ovlp.numpy(), "first1": first1.detach().numpy(), "first2": first2.detach().numpy(),
           "loss": loss.detach().numpy(), "g_first1": first1.grad.numpy().copy(), "g_first2": first2.grad.numpy().copy(),
           "trunc": np.asarray(cfg["training"]["trunc"]), "hash_size": np.asarray(10)}
    first1.grad = None; first2.grad = None
    mask = (torch.rand(N, 1, generator=g) > 0.3)
    loss2 = ref_im.InactiveMap.get_SDF_dif2(stub, rays[:, 6:7], rays[:, :3], mask, poses[3][None], 1, 0, first1, first2)
    loss2.backward()
    out.update(mask2=mask.numpy(), pose2=poses[3][None].numpy(), loss2=loss2.detach().numpy(),
               g2_first1=first1.grad.numpy().copy(), g2_first2=first2.grad.numpy().copy())
    for i, m in enumerate(models):
        for k, v in m.state_dict().items():
            out[f"w{i}:" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "overlap.npz"), **out)


def gen_keyframes():
    """The reference's own KeyframeSet (ray part) with python random.sample patched to replay recorded draws."""
    import random as pyrandom
    from model import keyframeSet as ref_kf
    g = torch.Generator().manual_seed(77)
    H, W, nh, nw, num_kf = 48, 64, 6, 8, 6
    ks = object.__new__(ref_kf.KeyframeSet)                      # the constructor needs the whole SLAM config; only the ray part is used
    ks.H, ks.W, ks.n_rays_h, ks.n_rays_w = H, W, nh, nw
    ks.num_rays_to_save = nh * nw
    ks.row_indices, ks.col_indices = ref_samp.sample_pixels_uniformly(H, W, nh, nw)
    ks.rays = torch.zeros((num_kf, ks.num_rays_to_save, 7))
    ks.frame_ids = None
    frames = []
    for k in range(num_kf):
        batch = {"direction": torch.randn(1, H, W, 3, generator=g), "rgb": torch.rand(1, H, W, 3, generator=g),
                 "depth": torch.rand(1, H, W, generator=g) * 4, "frame_id": 5 * k}
        ks.add_keyframe(batch)
        frames.append(batch)
    out = {"H": H, "W": W, "nh": nh, "nw": nw, "store": ks.rays.numpy(),
           "direction": torch.stack([f["direction"][0] for f in frames]).numpy(), "rgb": torch.stack([f["rgb"][0] for f in frames]).numpy(),
           "depth": torch.stack([f["depth"][0] for f in frames]).numpy()}
    draws = []
    real_sample = pyrandom.sample
    def recording_sample(population, k):
        r = real_sample(population, k); draws.append(r); return r
    ref_kf.random.sample = recording_sample
    pyrandom.seed(1234)
    cases = [("one", 2, [2], 20), ("two", 1, [1, 4], 30), ("many", 0, [0, 2, 3, 5], 37), ("five", 1, [1, 0, 2, 3, 4, 5], 64)]
    for name, first, related, pix in cases:
        draws.clear()
        rays, kf_ids, kf_indices = ks.sample_rays_in_submap(torch.tensor(first), torch.tensor(related), pix)
        out[f"{name}:first"] = first; out[f"{name}:related"] = np.array(related); out[f"{name}:pix"] = pix
        for i, d in enumerate(draws):
            out[f"{name}:draw{i}"] = np.array(d, dtype=np.int64)
        out[f"{name}:n_draws"] = len(draws)
        out[f"{name}:rays"] = rays.numpy(); out[f"{name}:kf_ids"] = kf_ids.numpy(); out[f"{name}:kf_indices"] = kf_indices.numpy()
    draws.clear()
    rays, kf_ids, kf_indices = ks.sample_rays_in_given_kf(torch.tensor([4, 1, 3]), 25)
    out["given:ids"] = np.array([4, 1, 3]); out["given:draw"] = np.array(draws[0], dtype=np.int64)
    out["given:rays"] = rays.numpy(); out["given:kf_ids"] = kf_ids.numpy(); out["given:kf_indices"] = kf_indices.numpy()
    ref_kf.random.sample = real_sample
    np.savez_compressed(os.path.join(OUT, "keyframes.npz"), **out)


def gen_manager():
    """The containment tests of the reference's OWN Manager (Manager.py:159-244), called unbound on a stand-in `self` that carries
    only the attributes the two methods read (dataset.H / W, kfSet.localMLP_info, min_cr_localMLP_len): lattice scores,
    the chosen submap and containing ratios on a synthetic frame with a few zero-depth pixels."""
    from Manager import Manager
    from mipsfusion_b200 import synth
    g = torch.Generator().manual_seed(11)
    dirs = synth.camera_rays()[::2, ::2].contiguous()            # (230, 310, 3): a half-resolution frame keeps the fixture small
    H, W = dirs.shape[0], dirs.shape[1]
    c2w = synth.trajectory(6)[2]
    depth = synth.render_frame(c2w, dirs)["depth"].clone()
    depth[torch.rand(H, W, generator=g) < 0.03] = 0.0            # missing depth
    k = 6
    centers = torch.tensor([1.0, 3.5, 1.0]) + 1.5 * (torch.rand(k, 3, generator=g) - 0.5)
    lens = 1.0 + 3.0 * torch.rand(k, 3, generator=g)
    info = torch.cat([torch.zeros(k, 1), centers, lens], -1)     # kfSet.localMLP_info rows: [flag, center 3, length 3]
    stub = types.SimpleNamespace(dataset=types.SimpleNamespace(H=H, W=W), kfSet=types.SimpleNamespace(localMLP_info=info),
                                 min_cr_localMLP_len=torch.tensor([2.0, 2.0, 2.0]))
    ids = torch.arange(k)
    out = dict(depth=depth.numpy(), c2w=c2w.numpy(),                # (dirs: synth.camera_rays()[::2, ::2], regenerated by the tests)
               centers=centers.numpy(), lens=lens.numpy(),
               min_len=stub.min_cr_localMLP_len.numpy())
    out["top_id"] = np.int64(Manager.find_highest_containing_ratio(stub, depth, dirs, c2w, ids))
    # the scores the method ranks (its Step 2, recomputed with the reference's own helpers)
    from helper_functions.geometry_helper import pts_in_bbox as ref_pib
    ih, iw = ref_samp.sample_pixels_uniformly(H, W, 15, 20)
    td, dc = depth[ih, iw], dirs[ih, iw]
    pts = (c2w[:3, -1].repeat(300, 1)[..., None, :] + torch.sum(dc[..., None, :] * c2w[None, :3, :3], -1)[..., None, :] * td[..., :, None]).reshape(-1, 3)
    # NB the reference's expression broadcasts (P,1,3) * (P,1) to (P,P,3): every ray direction is paired with every depth
    # (Manager.py:174) -- 90,000 points for the 15 x 20 lattice; that is the behaviour the scores below come from
    out["scores"] = torch.count_nonzero(ref_pib(pts, centers - 0.5 * lens, centers + 0.5 * lens), dim=0).numpy()
    sub = torch.randperm(pts.shape[0], generator=g)[:2000]
    out["pts_sub"] = pts[sub].numpy()
    out["mask_sub"] = ref_pib(pts[sub], centers - 0.5 * lens, centers + 0.5 * lens).numpy()
    out["ratios"] = np.array([float(Manager.compute_containing_ratio(stub, depth, dirs, c2w, torch.tensor(j))) for j in range(k)], dtype=np.float64)
    out["ratios_small"] = np.array([float(Manager.compute_containing_ratio(stub, depth, dirs, c2w, torch.tensor(j), rays_h=30, rays_w=40)) for j in range(k)], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "manager.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "keyframes":
        gen_keyframes(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "manager":
        gen_manager(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "overlap":
        gen_overlap(); sys.exit(0)
    gen_overlap()
    gen_keyframes()
    gen_manager()
    gen_lattice(); gen_sampling(); gen_losses(); gen_decoder()
    cfg, model = gen_scene()
    gen_ro(cfg, model)
    gen_blend()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
