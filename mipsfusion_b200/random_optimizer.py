"""Drop-in for reference RandomOptimizer.py: particle-swarm-template pose search whose hot loop
(candidate transform -> field query -> per-candidate fitness -> swarm update) runs as CUDA kernels
(mf_ro_score / mf_ro_update) with no host synchronisation inside the iteration loop.

Candidates shard naturally across GPUs: pass ``group`` (a torch.distributed process group) and every
rank scores C/G candidates, followed by one small all-gather of (fitness, mean_sdf, pst7) that the update kernel reads in
place (``iterate``)."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import dist as D
from .sampling_helper import sample_pixels_uniformly


def pose_compose(rot_mat, trans_vec):
    """reference helper_functions/geometry_helper.py:43-47 (returns a CPU tensor)."""
    T = torch.eye(4)
    T[:3, :3] = rot_mat
    T[:3, 3] = trans_vec.squeeze()
    return T


class RandomOptimizer:
    def __init__(self, cfg, mipsfusion, particles=None, group=None, peer_memory="auto"):
        self.cfg = cfg
        self.slam = mipsfusion
        self.dataset = self.slam.dataset
        self.device = torch.device(self.slam.device)
        ro = self.cfg["tracking"]["RO"]
        self.particle_size = ro["particle_size"]
        self.scaling_coefficient1 = ro["initial_scaling_factor"]
        self.scaling_coefficient2 = ro["rescaling_factor"]
        self.sdf_weight = 1000.
        self.trunc_value = self.cfg["training"]["trunc"]
        if particles is None:                                      # RandomOptimizer.py:26-32
            particles = np.random.multivariate_normal(np.zeros(6), np.eye(6), self.particle_size).astype(np.float32)
            particles = torch.from_numpy(particles)
            particles[0, :] = 0
            particles = torch.clamp(particles, -2., 2.)
        self.pre_sampled_particle = L.f32c(torch.as_tensor(particles), self.device)
        self.iW = self.cfg["tracking"]["ignore_edge_W"]
        self.iH = self.cfg["tracking"]["ignore_edge_H"]
        self.rays_dir = L.f32c(self.dataset.rays_d, self.device)   # (H, W, 3) camera-frame directions
        self.row_indices, self.col_indices = sample_pixels_uniformly(self.dataset.H, self.dataset.W, ro["n_rows"], ro["n_cols"],
                                                                     device=self.device)
        self.fx, self.fy, self.cx, self.cy = self.dataset.fx, self.dataset.fy, self.dataset.cx, self.dataset.cy
        self.group = group
        self.peer_memory = peer_memory     # multi-GPU exchange: True (require NVLink peer memory), False (NCCL all-gather), "auto"
        self.last_info = None          # per-iteration (count, success, argmin) of the last optimize() call
        self.last_fitness = None

    # ---- sharding ---------------------------------------------------------------------------------
    def _shard(self):
        Cn = self.pre_sampled_particle.shape[0]
        ws, rk = D.world(self.group)
        b, n, _ = D.shard_range(Cn, ws, rk)
        return b, n, ws, rk

    def score(self, model, rot_cur, trans_cur, search_size, target_d, rays_d_cam):
        """Fitness of every candidate (RandomOptimizer.py:113-131).  All arguments are device tensors:
        rot_cur (3,3), trans_cur (3), search_size (6), target_d (P), rays_d_cam (P,3).
        -> fitness (C), mean_sdf (C), pst7 (C,7)."""
        dev = self.device
        Cn, P = self.pre_sampled_particle.shape[0], target_d.shape[0]
        b, n, ws, rk = self._shard()
        per = (Cn + ws - 1) // ws
        packed = torch.zeros(per, 9, device=dev, dtype=torch.float32)        # [fitness, mean_sdf, pst7]
        fit = torch.empty(per, device=dev); msdf = torch.empty(per, device=dev); pst7 = torch.empty(per, 7, device=dev)
        scratch = torch.empty(max(n, 1) * (P + 12), device=dev, dtype=torch.float32)
        field = model._field()
        with torch.cuda.device(dev):
            L.call("mf_ro_score", L.ptr(self.pre_sampled_particle), L.ptr(search_size), L.ptr(rot_cur), L.ptr(trans_cur),
                   L.ptr(rays_d_cam), L.ptr(target_d), C.byref(field), float(self.trunc_value), float(self.sdf_weight),
                   int(b), int(n), int(P), L.ptr(fit), L.ptr(msdf), L.ptr(pst7), L.ptr(scratch), L.stream())
        if ws == 1:
            return fit[:Cn], msdf[:Cn], pst7[:Cn]
        packed[:, 0], packed[:, 1], packed[:, 2:] = fit, msdf, pst7
        gathered = D.allgather_rows(packed, Cn, self.group)          # one small all-gather: 9 floats per candidate
        return gathered[:, 0].contiguous(), gathered[:, 1].contiguous(), gathered[:, 2:].contiguous()

    def update(self, fitness, mean_sdf, pst7, rot_cur, trans_cur, search_size):
        """Steps 3-5 of the loop body (RandomOptimizer.py:202-224), in place on the device state."""
        dev = self.device
        Cn = fitness.shape[0]
        better = torch.empty(Cn, device=dev, dtype=torch.uint8)
        info = torch.empty(4, device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            L.call("mf_ro_update", L.ptr(fitness), L.ptr(mean_sdf), L.ptr(pst7), int(Cn), float(self.scaling_coefficient2),
                   L.ptr(rot_cur), L.ptr(trans_cur), L.ptr(search_size), L.ptr(better), L.ptr(info), L.stream())
        return better, info

    # ---- one iteration without intermediate tensors ---------------------------------------------------------------------
    def _iter_buffers(self, P):
        Cn = self.pre_sampled_particle.shape[0]
        b, n, ws, rk = self._shard()
        per = (Cn + ws - 1) // ws
        buf = self.__dict__.get("_ibuf")
        if buf is None or buf["key"] != (Cn, P, ws, rk):
            dev = self.device
            f32 = dict(device=dev, dtype=torch.float32)
            buf = dict(key=(Cn, P, ws, rk), local=torch.zeros(9 * per, **f32), scratch=torch.empty(max(n, 1) * (P + 12), **f32),
                       gathered=torch.empty(ws * 9 * per, **f32) if ws > 1 else None)
            self.__dict__["_ibuf"] = buf
        if ws > 1 and self.peer_memory and "_arena" not in self.__dict__:
            # symmetric (peer-mapped) arena for the exchange inside the update kernel; every rank reaches this point in its first
            # iteration (the rendezvous is collective).  Not available (no NVLink peer access, gloo tests): NCCL all-gather.
            arena = None
            try:
                arena = D.PeerArena({"gath": 2 * ws * 9 * per, "flags": ws}, self.device, self.group, use_multicast=False)
                torch.cuda.current_stream(self.device).synchronize()
                arena.barrier()
                torch.cuda.current_stream(self.device).synchronize()
            except Exception as e:
                if self.peer_memory is True:
                    raise
                self.__dict__["_arena_error"] = repr(e)
                arena = None
            self.__dict__["_arena"] = arena
            self.__dict__["_arena_per"] = per
            self.__dict__["_seq"] = 0
        return buf, b, n, ws, per

    def iterate(self, model, rot_cur, trans_cur, search_size, target_d, rays_d_cam, better=None, info=None):
        """One loop body (RandomOptimizer.py:192-224): score this rank's candidates and apply the swarm update, in place on
        ``rot_cur`` / ``trans_cur`` / ``search_size``.  The three per-candidate results of a rank are one contiguous block
        [fitness | mean_sdf | pst7] on persistent buffers; with a process group the blocks are all-gathered once (9 floats per
        candidate) and the update kernel reads the gathered blocks in place -- 4 launches + 1 collective per iteration, no
        allocation, no pack / unpack copies.  better (C) uint8 / info (4) int32: optional outputs.  -> (blocks, per)."""
        dev = self.device
        Cn, P = self.pre_sampled_particle.shape[0], target_d.shape[0]
        buf, b, n, ws, per = self._iter_buffers(P)
        loc = buf["local"]
        field = model._field()
        with torch.cuda.device(dev):
            st = L.stream()
            L.call("mf_ro_score", L.ptr(self.pre_sampled_particle), L.ptr(search_size), L.ptr(rot_cur), L.ptr(trans_cur),
                   L.ptr(rays_d_cam), L.ptr(target_d), C.byref(field), float(self.trunc_value), float(self.sdf_weight),
                   int(b), int(n), int(P), loc.data_ptr(), loc.data_ptr() + 4 * per, loc.data_ptr() + 8 * per, L.ptr(buf["scratch"]), st)
            blocks = loc
            arena = self.__dict__.get("_arena") if ws > 1 else None
            if arena is not None and self.__dict__["_arena_per"] == per:
                # exchange + update in one kernel over NVLink peer memory (mf_ro_update_peer): no collective call
                self.__dict__["_seq"] += 1
                seq = self.__dict__["_seq"]
                bases = (C.c_uint64 * ws)(*arena.peer_bases)
                L.call("mf_ro_update_peer", L.ptr(loc), bases, int(ws), int(arena.rank), int(arena.offsets["gath"]), int(arena.offsets["flags"]),
                       int(seq & 0xFFFFFFFF), int(Cn), int(per), float(self.scaling_coefficient2), L.ptr(rot_cur), L.ptr(trans_cur),
                       L.ptr(search_size), L.ptr(better), L.ptr(info), st)
                n = ws * 9 * per
                blocks = arena.view("gath")[(seq & 1) * n:(seq & 1) * n + n]
                return blocks, per
            if ws > 1:
                import torch.distributed as dist
                blocks = buf["gathered"]
                dist.all_gather_into_tensor(blocks, loc, group=self.group)
            L.call("mf_ro_update_gathered", L.ptr(blocks), int(Cn), int(per), float(self.scaling_coefficient2), L.ptr(rot_cur),
                   L.ptr(trans_cur), L.ptr(search_size), L.ptr(better), L.ptr(info), st)
        return blocks, per

    def _lattice(self, off):
        """Pixel lattice shifted by ``off`` (RandomOptimizer.py:186-190) and its camera-frame directions (cached: they do not
        depend on the frame)."""
        cache = self.__dict__.setdefault("_lat", {})
        e = cache.get(off)
        if e is None:
            ih, iw = self.row_indices + off, self.col_indices + off
            e = cache[off] = (ih, iw, self.rays_dir[ih, iw, :].contiguous())
        return e

    @torch.no_grad()
    def optimize(self, model, depth_img, initial_pose, last_frame_pose, n_iter=10):
        """reference RandomOptimizer.py:165-227; returns the tracked pose as a CPU (4,4) tensor."""
        if n_iter <= 0:
            return initial_pose
        dev = self.device
        init = L.f32c(torch.as_tensor(initial_pose), dev)
        rot_cur = init[:3, :3].contiguous()
        trans_cur = init[:3, 3].contiguous()
        search = torch.full((6,), float(self.scaling_coefficient1), device=dev, dtype=torch.float32)
        depth = L.f32c(torch.as_tensor(depth_img), dev)            # one H2D copy of the depth image
        Cn = self.pre_sampled_particle.shape[0]
        infos = torch.empty(n_iter, 4, device=dev, dtype=torch.int32)
        betters = torch.empty(n_iter, Cn, device=dev, dtype=torch.uint8)
        fits = []
        for i in range(n_iter):
            ih, iw, rays_d_cam = self._lattice(i % 5)
            target_d = depth[ih, iw].contiguous()
            blocks, per = self.iterate(model, rot_cur, trans_cur, search, target_d, rays_d_cam, better=betters[i], info=infos[i])
            fits.append(blocks.view(-1, 9 * per)[:, :per].reshape(-1)[:Cn].clone() if per < Cn else blocks[:Cn].clone())   # diagnostics
        self.last_info = infos                      # device tensors; reading them is the caller's sync point
        self.last_fitness, self.last_better = fits, list(betters)
        return pose_compose(rot_cur.cpu(), trans_cur.cpu())
