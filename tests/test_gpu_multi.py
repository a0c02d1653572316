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
