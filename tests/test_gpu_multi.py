"""Two-GPU checks of the data-parallel mapper (skipped on a single-GPU box): the peer-memory route (one sharded
reduce + Adam + broadcast kernel over NVLink) must give the same parameters as the NCCL all-reduce + replicated Adam
route, and all replicas must stay bit-identical."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as H

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from mipsfusion_b200.mapper import FusedMapper
    cfg = H.make_config(12, n_samples_d=32, n_range_d=11)
    cfg["training"]["perturb"] = 0
    of = H.oracle_field(cfg, seed=3)
    R, S = 256, 43
    rays_o, rays_d, rgb, d, _ = H.synth_batch(R, S, seed=20 + rank)
    args = [t.to(dev).contiguous() for t in (rays_o, rays_d, rgb, d)]
    res, mappers = {}, {}
    for name, pm in (("nccl", False), ("peer", True)):
        m = FusedMapper(H.cuda_model(cfg, H.state_of(of)), group=dist.group.WORLD, peer_memory=pm)
        assert (m.arena is not None) == pm
        mappers[name] = m
    if mappers["peer"].arena.multicast_base:                       # NVSwitch multicast (NVLS) variant of the same kernel
        mappers["peer_mc"] = FusedMapper(H.cuda_model(cfg, H.state_of(of)), group=dist.group.WORLD, peer_memory=True, multicast=True)
    # (1) the update alone on identical, seeded per-rank gradients: the two routes must agree bit for bit at world 2
    gen = torch.Generator().manual_seed(100 + rank)
    gg = (torch.randn(mappers["nccl"].grid.numel(), generator=gen) * 1e-3).to(dev)
    gm = (torch.randn(mappers["nccl"].mlp.numel(), generator=gen) * 1e-3).to(dev)
    gg[::3] = 0.0
    upd = {}
    for name, m in mappers.items():
        for _ in range(2):                                          # two updates: both gradient buffers of the peer route
            m.g_grid.copy_(gg); m.g_mlp.copy_(gm)
            if m.arena is not None:
                m.apply_gradients_sharded()
            else:
                from mipsfusion_b200 import dist as D
                D.average_gradients_([m.g_grid, m.g_mlp], m.group)
                m.apply_gradients()
        torch.cuda.synchronize()
        upd[name] = (m.grid.detach().cpu().numpy().copy(), m.mlp.detach().cpu().numpy().copy(),
                     float(m.g_grid.abs().max()), float(m.g_mlp.abs().max()))
    # (2) whole mapping steps
    for name, m in mappers.items():
        losses = [m.step(*args).cpu().numpy().copy() for _ in range(4)]
        torch.cuda.synchronize()
        res[name] = (m.grid.detach().cpu().numpy().copy(), m.mlp.detach().cpu().numpy().copy(), np.array(losses))
    g = res["peer"][0]
    others = [torch.empty_like(torch.from_numpy(g)).to(dev) for _ in range(world)]
    dist.all_gather(others, torch.from_numpy(g).to(dev))
    same = all(bool(torch.equal(o, others[0])) for o in others)
    if rank == 0:
        extra = {}
        if "peer_mc" in upd:
            extra = dict(upd_grid_mc=upd["peer_mc"][0], upd_mlp_mc=upd["peer_mc"][1], loss_mc=res["peer_mc"][2])
        np.savez(out, **extra, upd_grid_nccl=upd["nccl"][0], upd_grid_peer=upd["peer"][0], upd_mlp_nccl=upd["nccl"][1], upd_mlp_peer=upd["peer"][1],
                 cleared=np.array([upd["peer"][2], upd["peer"][3]]), loss_nccl=res["nccl"][2], loss_peer=res["peer"][2], same=np.array(same))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_adam_matches_nccl_route(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, 29641, out), nprocs=2, join=True)
    r = np.load(out)
    assert bool(r["same"]), "replicas diverged"
    # identical gradients in, identical parameters out (two ranks: the sum order cannot differ); gradient buffers cleared
    np.testing.assert_array_equal(r["upd_grid_peer"], r["upd_grid_nccl"])
    np.testing.assert_array_equal(r["upd_mlp_peer"], r["upd_mlp_nccl"])
    assert r["cleared"].max() == 0.0
    if "upd_grid_mc" in r:
        np.testing.assert_array_equal(r["upd_grid_mc"], r["upd_grid_nccl"])
        np.testing.assert_array_equal(r["upd_mlp_mc"], r["upd_mlp_nccl"])
        np.testing.assert_allclose(r["loss_mc"], r["loss_nccl"], rtol=2e-3)
    # whole steps: the scatter order of the grid gradient is not deterministic, so only the losses are compared (fp32 tolerance)
    np.testing.assert_allclose(r["loss_peer"], r["loss_nccl"], rtol=2e-3)


def _submap_worker(rank, world, port, out):
    """Submap-parallel placement on two GPUs: submap m on rank m % 2."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import mipsfusion_b200 as mf
    from mipsfusion_b200.submap_parallel import SubmapParallel
    from oracle import overlap as oov
    res = {}
    sp = SubmapParallel(dist.group.WORLD)
    # (1) joint query, 4 submaps: submap-sharded (partial sums all-reduced) == all submaps on one GPU
    cfg = H.make_config(12)
    cfg["grid"]["use_bound_normalize"] = False
    fields = [H.oracle_field(cfg, grid_scale=0.4, seed=30 + i) for i in range(4)]
    models = [H.cuda_model(cfg, H.state_of(f), train=False) for f in fields]       # (replicated here so that rank 0 can also run the reference)
    lo = np.array([-0.2, 1.0, -0.6]); poses, amin, amax, cents = [], [], [], []
    for i in range(4):
        a = lo + np.array([0.5 * i, 0.4 * i, 0.0]); b = a + np.array([1.6, 2.4, 2.2])
        T = torch.eye(4); T[:3, 3] = torch.tensor((a + b) / 2, dtype=torch.float32)
        poses.append(T); amin.append(a); amax.append(b); cents.append(((a + b) / 2 + 0.1).astype(np.float32))
    axes = mf.get_grid_uniform(lo, lo + np.array([3.1, 3.6, 2.2]), voxel_size=0.09)
    jq = mf.JointSubmapQuery(models, poses, amin, amax, cents, device=dev)
    sharded = jq.query(axes=axes, group=dist.group.WORLD, shard="submaps", want_contain=True)
    single = jq.query(axes=axes, want_contain=True)
    res["jq_sdf_err"] = float((sharded["sdf"] - single["sdf"]).abs().max())
    res["jq_masks_equal"] = bool(torch.equal(sharded["mask"], single["mask"]) and torch.equal(sharded["contain"], single["contain"]))
    # (2) weight hand-off of one T=2^19 submap: rank 0 -> rank 1, then broadcast
    big = H.make_config(19)
    src_field = H.oracle_field(big, grid_scale=0.1, seed=5)
    m = H.cuda_model(big, H.state_of(src_field if rank == 0 else H.oracle_field(big, seed=6)), train=False)
    torch.cuda.synchronize(); dist.barrier()
    a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_ev.record(); nbytes = sp.handoff(m, src=0, dst=1); b_ev.record(); torch.cuda.synchronize()
    res["handoff_ms"], res["handoff_bytes"] = a_ev.elapsed_time(b_ev), nbytes
    ok = torch.equal(m.embed_fn.params.data.cpu(), src_field.grid.detach())
    ok &= all(torch.equal(p.data.cpu(), src_field.w[k].detach()) for k, p in m.decoder.named_parameters())
    res["handoff_exact"] = bool(ok)
    pts = torch.rand(257, 3, generator=torch.Generator().manual_seed(1)) * torch.tensor([3.5, 6.5, 4.2]) + torch.tensor([-0.6, 0.5, -1.15])
    with torch.no_grad():
        res["handoff_query_err"] = H.rel_err(m.run_network(pts.to(dev)).cpu(), src_field.run_network(pts))   # the received field evaluates
    # (3) cross-rank overlap SDF difference (InactiveMap.get_SDF_dif2): rank r owns submap r
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "overlap.npz"))
    ocfg = H.make_config(int(fx["hash_size"]))
    st = lambda i: {k[len(f"w{i}:"):]: torch.from_numpy(v) for k, v in fx.items() if k.startswith(f"w{i}:")}
    local = {rank: H.cuda_model(ocfg, st(rank), train=False)}
    t = lambda k: torch.from_numpy(fx[k]).to(dev)
    rays = t("rays"); target_d, dirs = rays[:, 6:7], rays[:, :3]
    mask = torch.where(target_d > 0., torch.ones_like(target_d), torch.zeros_like(target_d))
    f1 = t("first1").requires_grad_(True); f2 = t("first2").requires_grad_(True)
    loss = sp.overlap_sdf_difference(local, 0, 1, target_d, dirs, mask, t("ovlp"), f1, f2, float(fx["trunc"]))
    loss.backward()
    res["ovl_loss"] = float(loss.detach())
    res["ovl_grad"] = (f1.grad if rank == 0 else f2.grad).cpu().numpy()
    res["ovl_other_none"] = (f2.grad if rank == 0 else f1.grad) is None
    np.savez(out + f".{rank}.npz", **res)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_submap_parallel_on_two_gpus(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "sp")
    mp.spawn(_submap_worker, args=(2, 29653, out), nprocs=2, join=True)
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "overlap.npz"))
    for rank in range(2):
        r = np.load(out + f".{rank}.npz")
        assert bool(r["jq_masks_equal"]) and float(r["jq_sdf_err"]) < 1e-5, (rank, float(r["jq_sdf_err"]))
        np.testing.assert_allclose(float(r["ovl_loss"]), float(fx["loss"]), rtol=1e-3)
        ref = fx["g_first1"] if rank == 0 else fx["g_first2"]
        assert bool(r["ovl_other_none"])
        print(f"\n  rank {rank}: hand-off {int(r['handoff_bytes']) / 1e6:.1f} MB in {float(r['handoff_ms']):.3f} ms, "
              f"overlap pose-gradient rel err {H.rel_err(r['ovl_grad'], ref):.2e}")
        assert H.rel_err(r["ovl_grad"], ref) < 5e-2                  # (one sample on a cell boundary, see test_overlap.py)
    r1 = np.load(out + ".1.npz")
    assert bool(r1["handoff_exact"]) and float(r1["handoff_query_err"]) < 1e-4


def _ro_worker(rank, world, port, out):
    import types
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import mipsfusion_b200 as mf
    from mipsfusion_b200 import synth
    cfg = H.make_config(14)
    cfg["tracking"] = {"RO": {"particle_size": 300, "initial_scaling_factor": 0.02, "rescaling_factor": 0.5, "n_rows": 12, "n_cols": 16},
                       "ignore_edge_W": 20, "ignore_edge_H": 20}
    of = H.oracle_field(cfg, grid_scale=0.3, seed=4)
    model = H.cuda_model(cfg, H.state_of(of), train=False)
    dirs = synth.camera_rays()
    c2w = synth.trajectory(4)[1]
    depth = synth.render_frame(c2w, dirs)["depth"]
    ds = types.SimpleNamespace(H=460, W=620, fx=320.0, fy=320.0, cx=309.5, cy=229.5, rays_d=dirs)
    g = torch.Generator().manual_seed(1)
    particles = torch.randn(300, 6, generator=g).clamp(-2, 2); particles[0] = 0
    slam = types.SimpleNamespace(dataset=ds, device=str(dev))
    start = c2w.clone(); start[:3, 3] += torch.tensor([0.01, -0.01, 0.005])
    res = {}
    for name, kw in (("single", dict(group=None)), ("nccl", dict(group=dist.group.WORLD, peer_memory=False)),
                     ("peer", dict(group=dist.group.WORLD, peer_memory=True))):
        ro = mf.RandomOptimizer(cfg, slam, particles=particles.clone(), **kw)
        poses = [ro.optimize(model, depth, start.clone(), start.clone(), n_iter=7).numpy().copy() for _ in range(2)]   # 14 exchanges: both buffer parities
        torch.cuda.synchronize()
        assert (ro.__dict__.get("_arena") is not None) == (name == "peer")
        res[name] = (np.stack(poses), ro.last_info.cpu().numpy().copy())
    from mipsfusion_b200 import _lib as L
    assert L.lib().mf_tc_check_error() == 0
    np.savez(out + f".{rank}.npz", **{f"{k}_pose": v[0] for k, v in res.items()}, **{f"{k}_info": v[1] for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_random_optimizer_sharded_routes_equal_single_gpu(tmp_path):
    """RandomOptimizer with the 300 candidates sharded over two GPUs -- NCCL all-gather + update, and the one-kernel exchange +
    update over NVLink peer memory -- against the single-GPU loop: per-candidate fitness is reduced in a fixed order and the
    update kernel sums the candidates in a fixed order, so the tracked poses and the per-iteration decisions (better count,
    success, argmin) must be bit-identical on every rank."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "ro")
    mp.spawn(_ro_worker, args=(2, 29667, out), nprocs=2, join=True)
    for rank in range(2):
        r = np.load(out + f".{rank}.npz")
        for name in ("nccl", "peer"):
            np.testing.assert_array_equal(r[f"{name}_pose"], r["single_pose"])
            np.testing.assert_array_equal(r[f"{name}_info"], r["single_info"])
