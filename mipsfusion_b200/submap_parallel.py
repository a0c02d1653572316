"""Submap-parallel placement for one multi-GPU box (SURVEY.md 8e "Submap-parallel online SLAM", BASELINE configs[3]).

The reference keeps the active submap in one OS process and optimises inactive submaps in a second one, copying model
weights between them whenever the active submap changes (mipsfusion.py:607-653: ``copy.deepcopy`` + ``load_state_dict`` +
``share_memory``; InactiveMap.py:61-96,203-308).  With one process per GPU the same structure is: submap ``m`` lives on rank
``m % G``; rank 0 tracks and maps the active submap; the other ranks run the inactive-submap BA; and the three places where
submaps meet become collectives over NVLink:

* :meth:`handoff` -- one submap's weights (hash grid 36 MB + decoder 146 KB) move to another rank or to all of them
  (``ncclBroadcast`` / send-recv), the replacement of the deepcopy hand-off at a submap switch;
* :meth:`overlap_sdf_difference` -- InactiveMap.get_SDF_dif / get_SDF_dif2 (InactiveMap.py:128-192) when the two submaps sit
  on different ranks: the sample rays are replicated (<= 768 x 7 floats), each owner queries its own field, the two SDF
  vectors (2 x N floats) are exchanged, and each owner back-propagates into the pose of its own submap;
* the joint query with ``shard="submaps"`` (:class:`mipsfusion_b200.JointSubmapQuery`) for the mesher.

Only torch.distributed is used (NCCL on the GPUs, gloo in the CPU tests); the field evaluations are the library's kernels.
"""
import torch

from . import dist as D


class SubmapParallel:
    def __init__(self, group=None):
        self.group = group
        self.world, self.rank = D.world(group)

    # ---- placement --------------------------------------------------------------------------------
    def owner(self, submap_id):
        """Rank that holds submap ``submap_id`` (round robin: 16 submaps on 8 GPUs = 2 per GPU)."""
        return int(submap_id) % self.world

    def local_ids(self, n_submaps):
        return D.round_robin(int(n_submaps), self.world, self.rank)

    # ---- weight hand-off ---------------------------------------------------------------------------
    @staticmethod
    def _tensors(model):
        """The tensors of one submap in state_dict order: the flat hash grid and the ten decoder tensors."""
        return [model.embed_fn.params.data] + [p.data for p in model.decoder.ordered_params()]

    @torch.no_grad()
    def handoff(self, model, src, dst=None):
        """Move the weights of ``model`` from rank ``src`` to rank ``dst`` (or to every rank when ``dst`` is None).  Every
        participating rank passes its own module of the same architecture; non-participants of a point-to-point hand-off return
        immediately.  The decoder travels as one flat buffer.  -> number of bytes this rank sent or received."""
        import torch.distributed as dist
        if self.world == 1:
            return 0
        grid, dec = self._tensors(model)[0], self._tensors(model)[1:]
        flat = torch.cat([t.reshape(-1) for t in dec])
        nbytes = 4 * (grid.numel() + flat.numel())
        if dst is None:
            dist.broadcast(grid, src=dist.get_global_rank(self.group, src) if self.group is not None else src, group=self.group)
            dist.broadcast(flat, src=dist.get_global_rank(self.group, src) if self.group is not None else src, group=self.group)
        else:
            if self.rank == src:
                dist.send(grid, dst=dst, group=self.group); dist.send(flat, dst=dst, group=self.group)
            elif self.rank == dst:
                dist.recv(grid, src=src, group=self.group); dist.recv(flat, src=src, group=self.group)
            else:
                return 0
        if self.rank != src:
            o = 0
            for t in dec:
                t.copy_(flat[o:o + t.numel()].view(t.shape))
                o += t.numel()
            for p in [model.embed_fn.params] + list(model.decoder.ordered_params()):
                torch.autograd.graph.increment_version(p)          # cached kernel-layout weights are rebuilt on the next call
        return nbytes

    # ---- cross-rank overlap query -------------------------------------------------------------------
    def overlap_sdf_difference(self, local_models, id1, id2, target_d, rays_d_cam, mask, ovlp_kf_pose, first_kf_pose1, first_kf_pose2,
                               trunc_value):
        """InactiveMap.get_SDF_dif2 (InactiveMap.py:177-192) for submaps on different ranks.  ``local_models``: dict submap id ->
        model of the submaps THIS rank owns (any object with ``run_network``).  All other arguments are replicated on every rank.
        Every rank returns the same loss value; on an owner it carries the autograd graph of its own submap's pose
        (``first_kf_pose1`` on the owner of ``id1``, ``first_kf_pose2`` on the owner of ``id2``) -- the other submap's SDF enters
        as a constant, which is exactly its role in the derivative."""
        import torch.distributed as dist
        mask = mask.to(target_d)
        N = target_d.shape[0]
        own = [int(id1) in local_models, int(id2) in local_models]
        sdf = [None, None]
        for j, (sid, first) in enumerate(((id1, first_kf_pose1), (id2, first_kf_pose2))):
            if own[j]:
                local_poses = first.inverse() @ ovlp_kf_pose
                sdf[j] = _infer_sdf(local_poses, local_models[int(sid)], rays_d_cam, target_d, trunc_value)
        if self.world > 1:
            buf = torch.zeros(2, N, 1, device=target_d.device, dtype=torch.float32)
            for j in range(2):
                # the owner with the lowest rank contributes (a submap may be replicated after a broadcast hand-off)
                if own[j] and self.rank == self.owner((id1, id2)[j]):
                    buf[j] = sdf[j].detach()
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)       # 2 x N floats
            for j in range(2):
                if not own[j]:
                    sdf[j] = buf[j]
        loss = torch.sum(torch.square(sdf[0] * mask - sdf[1] * mask))          # geometry_helper.py:225-229
        return loss / (torch.count_nonzero(mask) + 0.001)


def _infer_sdf(local_poses, model, rays_d_cam, target_d, trunc_value):
    """InactiveMap.infer_pts (InactiveMap.py:128-139), SDF part."""
    rays_d = torch.sum(rays_d_cam[..., None, None, :] * local_poses[..., None, :3, :3], -1)
    rays_o = local_poses[..., None, :3, -1].repeat(1, rays_d.shape[1], 1).reshape(-1, 3)
    rays_d = rays_d.reshape(-1, 3)
    pts_local = (rays_o[..., None, :] + rays_d[..., None, :] * target_d[..., :, None]).reshape(-1, 3)
    return model.run_network(pts_local)[..., 3:4] * trunc_value
