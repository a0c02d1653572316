"""world_size-2 gloo tests (CPU) of the host-side sharding logic of the multi-GPU paths: candidate shards +
all-gather, submap / point-range shards + reductions, and the data-parallel mapping recipe (global mask counts,
gradient averaging), with the oracle standing in for the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mipsfusion_b200 import dist as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world_size=2):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(fn, r, world_size, port, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    out = {}
    while not q.empty():
        r, res = q.get()
        out[r] = res
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(out) == world_size
    return out


def _worker(fn, rank, world_size, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        q.put((rank, fn(rank, world_size)))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 1024, 2000):
        for ws in (1, 2, 3, 8):
            seen = []
            for r in range(ws):
                b, n, per = D.shard_range(total, ws, r)
                assert n >= 0 and b + n <= total and n <= per
                seen += list(range(b, b + n))
            assert seen == list(range(total))
            assert sorted(sum((D.round_robin(total, ws, r) for r in range(ws)), [])) == list(range(total))


def _candidate_gather(rank, ws):
    Cn = 1001                                             # not divisible by the world size
    g = torch.Generator().manual_seed(0)
    full = torch.randn(Cn, 9, generator=g)
    b, n, per = D.shard_range(Cn, ws, rank)
    local = torch.zeros(per, 9)
    local[:n] = full[b:b + n]
    out = D.allgather_rows(local, Cn, dist.group.WORLD)
    return bool(torch.equal(out, full))


def test_candidate_allgather():
    assert all(_run(_candidate_gather).values())


def _submap_reduce(rank, ws):
    G, M = 500, 5
    g = torch.Generator().manual_seed(1)
    contrib = torch.rand(M, G, 2, generator=g)            # per-submap partial sums (w*sdf, w)
    mask = torch.rand(M, G, generator=g) < 0.4
    mine = D.round_robin(M, ws, rank)
    acc = torch.zeros(G, 2)
    m_any = torch.zeros(G, dtype=torch.int32)
    for m in mine:
        acc += contrib[m] * mask[m][:, None]
        m_any = torch.maximum(m_any, mask[m].to(torch.int32))
    D.allreduce_sum_(acc, dist.group.WORLD)
    D.allreduce_max_(m_any, dist.group.WORLD)
    ref = (contrib * mask[..., None]).sum(0)
    md = torch.tensor([float(rank + 1), 5.0 - rank])
    D.allreduce_max_(md, dist.group.WORLD)
    return bool(torch.allclose(acc, ref, atol=1e-5) and torch.equal(m_any.bool(), mask.any(0)) and md.tolist() == [2.0, 5.0])


def test_submap_sharded_reduction():
    assert all(_run(_submap_reduce).values())


def _dp_mapping(rank, ws):
    """Each rank: its own ray batch, GLOBAL mask counts, local loss -> gradient average == single-process
    gradient of the concatenated batch (the oracle plays the kernels)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import helpers as H
    from oracle import scene as oscene
    cfg = H.make_config(10, n_samples_d=16, n_range_d=7)
    R, S = 24, 23
    of = H.oracle_field(cfg, grid_scale=0.3, seed=3)
    batches = [H.synth_batch(R, S, seed=10 + r) for r in range(ws)]
    ro, rd, rgb, d, u = batches[rank]
    rend = of.render_rays(ro, rd, target_d=d, u=u)
    # --- the DP recipe of FusedMapper.step ---
    trunc = cfg["training"]["trunc"]
    fm, sm, _, _, counts = oscene.get_masks(rend["z_vals"], d, trunc)
    cnt = torch.tensor(counts, dtype=torch.int64)
    D.allreduce_sum_(cnt, dist.group.WORLD)
    n = float(cnt[0] + cnt[1])
    fs_w, sdf_w = 1.0 - float(cnt[0]) / n, 1.0 - float(cnt[1]) / n
    sdf, prob, z = rend["raw"][..., 3], rend["raw"][..., 5:], rend["z_vals"]
    idx = torch.arange(0, 5).to(prob)
    fs = torch.nn.functional.mse_loss(sdf * fm, fm) * fs_w + 0.01 * torch.mean(torch.sum(prob * (4 - idx) * fm[..., None], -1)) / 250
    gt = (((d - z) + trunc) / (2 * trunc)) * 4
    sl = torch.nn.functional.mse_loss((z + sdf * trunc) * sm, d * sm) * sdf_w + \
        0.01 * torch.mean(torch.sum(torch.abs(gt[:, :, None] - idx[None, None]) * sm[..., None] * prob, -1)) / 5000
    valid = (d.squeeze(-1) > 0) & (d.squeeze(-1) < cfg["cam"]["depth_trunc"])
    rgb_l = torch.nn.functional.mse_loss(rend["rgb"] * valid[:, None], rgb * valid[:, None])
    (rgb_l + 1000 * sl + 10 * fs).backward()
    grads = [of.grid.grad] + [p.grad for p in of.w.values()]
    D.average_gradients_(grads, dist.group.WORLD)
    # --- single-process reference on the concatenated batch ---
    of2 = H.oracle_field(cfg, grid_scale=0.3, seed=3)
    cat = [torch.cat([b[i] for b in batches], 0) for i in range(5)]
    ret = of2.forward(*cat)
    of2.total_loss(ret).backward()
    ref = [of2.grid.grad] + [p.grad for p in of2.w.values()]
    err = max(float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(grads, ref))
    return err


def test_data_parallel_mapping_recipe_matches_global_batch():
    errs = _run(_dp_mapping)
    assert max(errs.values()) < 1e-4, errs


def _overlap_and_handoff(rank, ws):
    """Submap-parallel placement (mipsfusion_b200.submap_parallel): the cross-rank overlap SDF difference equals the
    single-process composition (loss and the gradient w.r.t. each submap's first-keyframe pose, which only its owner holds),
    and a weight hand-off reproduces the source submap bit for bit.  The oracle fields play the models."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import helpers as H
    from oracle import overlap as oov
    from mipsfusion_b200.submap_parallel import SubmapParallel
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "overlap.npz"))
    cfg = H.make_config(int(fx["hash_size"]))
    state = lambda i: {k[len(f"w{i}:"):]: torch.from_numpy(v) for k, v in fx.items() if k.startswith(f"w{i}:")}
    models = [H.oracle_field(cfg, state(i)) for i in range(2)]
    t = lambda k: torch.from_numpy(fx[k])
    trunc = float(fx["trunc"])
    sp = SubmapParallel(dist.group.WORLD)
    assert [sp.owner(m) for m in range(4)] == [0, 1, 0, 1] and sp.local_ids(5) == [m for m in range(5) if m % ws == rank]
    rays = t("rays")
    target_d, dirs = rays[:, 6:7], rays[:, :3]
    mask = torch.where(target_d > 0., torch.ones_like(target_d), torch.zeros_like(target_d))
    f1 = t("first1").requires_grad_(True); f2 = t("first2").requires_grad_(True)
    local = {m: models[m] for m in sp.local_ids(2)}                  # rank r owns submap r
    loss = sp.overlap_sdf_difference(local, 0, 1, target_d, dirs, mask, t("ovlp"), f1, f2, trunc)
    loss.backward()
    mine = f1.grad if rank == 0 else f2.grad
    other = f2.grad if rank == 0 else f1.grad
    ok = abs(float(loss) - float(fx["loss"])) <= 2e-5 * abs(float(fx["loss"]))
    ref = fx["g_first1"] if rank == 0 else fx["g_first2"]
    ok &= float(np.abs(mine.numpy() - ref).max()) <= 2e-5 * float(np.abs(ref).max())
    ok &= other is None                                              # the other submap's pose is a constant on this rank
    # hand-off of submap 0 (rank 0) to rank 1: a module-like shim over the oracle field
    class Shim:
        pass
    def shim(field):
        s = Shim(); s.embed_fn = Shim(); s.decoder = Shim()
        s.embed_fn.params = field.grid
        s.decoder.ordered_params = lambda: [field.w[k] for k in field.w]
        return s
    dst_field = H.oracle_field(cfg, seed=99)
    moved = sp.handoff(shim(models[0] if rank == 0 else dst_field), src=0, dst=1)
    if rank == 1:
        ok &= bool(torch.equal(dst_field.grid, models[0].grid)) and all(torch.equal(dst_field.w[k], models[0].w[k]) for k in dst_field.w)
    ok &= moved == 4 * (models[0].grid.numel() + sum(v.numel() for v in models[0].w.values()))
    return bool(ok)


def test_submap_parallel_overlap_query_and_handoff():
    assert all(_run(_overlap_and_handoff).values())
