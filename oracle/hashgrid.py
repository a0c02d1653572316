"""Oracle: multi-resolution hash-grid encoding (tiny-cuda-nn 1.7 "HashGrid").

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **parity unpinned**: the
arithmetic lives in the third-party dependency tinycudann==1.7 (reference
environment.yaml:74), whose source is not under /root/reference.  This file
restates its published algorithm (NVlabs/tiny-cuda-nn v1.7,
include/tiny-cuda-nn/encodings/grid.h: grid_scale / grid_resolution /
pos_fract / grid_index / coherent-prime grid_hash / kernel_grid), anchored on
the reference's call sites:

  * model/encodings.py:11-26   -- config handed to tcnn.Encoding
  * model/scene_rep.py:40      -- desired_resolution=256, hash_size from cfg
  * model/scene_rep.py:122     -- embed_fn(inputs_flat)

uint32 wrap-around is emulated in int64 with ``& 0xFFFFFFFF``; the independent
plain-C restatement with real uint32/fmaf arithmetic is oracle/hashgrid_ref.c.
"""
import numpy as np
import torch

M32 = 0xFFFFFFFF
PRIME1 = 2654435761
PRIME2 = 805459861


def per_level_scale_of(desired_resolution=256, base_resolution=16, n_levels=16):
    # reference model/encodings.py:13
    return float(np.exp2(np.log2(desired_resolution / base_resolution) / (n_levels - 1)))


def level_table(log2_hashmap_size=19, n_levels=16, base_resolution=16,
                per_level_scale=None, n_features=2):
    """Per-level (scale, resolution, entries, offset); tcnn GridEncodingTemplated ctor."""
    if per_level_scale is None:
        per_level_scale = per_level_scale_of(256, base_resolution, n_levels)
    # tcnn reads per_level_scale from JSON as float, takes std::log2 of it (float)
    log2_pls = np.log2(np.float32(per_level_scale)).astype(np.float32)
    scales, ress, sizes, offsets, strides_hash = [], [], [], [0], []
    for l in range(n_levels):
        # grid_scale(): exp2f(level * log2_per_level_scale) * base_resolution - 1.0f.
        # exp2f is taken as the correctly rounded fp32 result (evaluated in fp64, rounded
        # once): libm / numpy / CUDA exp2f disagree in the last ulp (level 6), and the
        # scale decides integer cell indices, so the oracle fixes it normatively.
        arg = np.float32(np.float32(l) * log2_pls)
        s = np.float32(np.exp2(np.float64(arg))) * np.float32(base_resolution) - np.float32(1.0)
        s = np.float32(s)
        res = int(np.ceil(s)) + 1                      # grid_resolution()
        dense = res ** 3
        max_params = M32 // 2
        n = max_params if float(dense) > float(max_params) else dense
        n = (n + 7) // 8 * 8                           # next_multiple(.., 8)
        n = min(n, 1 << log2_hashmap_size)             # GridType::Hash
        scales.append(s); ress.append(res); sizes.append(n); offsets.append(offsets[-1] + n)
    return {
        "n_levels": n_levels, "n_features": n_features, "base_resolution": base_resolution,
        "log2_hashmap_size": log2_hashmap_size, "per_level_scale": per_level_scale,
        "scale": np.asarray(scales, dtype=np.float32),
        "resolution": np.asarray(ress, dtype=np.int64),
        "size": np.asarray(sizes, dtype=np.int64),
        "offset": np.asarray(offsets, dtype=np.int64),
        "n_params": offsets[-1] * n_features,
    }


def _level_index(pg, res, size):
    """grid_index<3, CoherentPrime>: pg is (N, 3) int64 holding uint32 values."""
    stride, idx = 1, torch.zeros_like(pg[:, 0])
    for d in range(3):
        if stride <= size:
            idx = (idx + pg[:, d] * stride) & M32
            stride = (stride * res) & M32
    if size < stride:  # hashed level
        idx = pg[:, 0] ^ ((pg[:, 1] * PRIME1) & M32) ^ ((pg[:, 2] * PRIME2) & M32)
    return idx % size


def grid_corners(x, table):
    """Integer + weight stream of the encoding.

    x: (N, 3) float32.  Returns (idx, w, pos_grid):
      idx      (N, L, 8) int64  -- entry index inside the level (before offset)
      w        (N, L, 8) float32 -- trilinear corner weight (differentiable in x)
      pos_grid (N, L, 3) int64  -- uint32 cell coordinate
    Corner c has bit d set <=> +1 along dim d (tcnn kernel_grid corner loop).
    """
    assert x.dtype == torch.float32 and x.shape[-1] == 3
    L = table["n_levels"]
    idx_all, w_all, pg_all = [], [], []
    for l in range(L):
        scale = float(table["scale"][l]); res = int(table["resolution"][l]); size = int(table["size"][l])
        # pos_fract(): pos = fmaf(scale, x, 0.5f); exact product in fp64, one rounding
        pos = (x.double() * scale + 0.5).float()
        tmp = torch.floor(pos)
        pg = tmp.detach().to(torch.int64) & M32           # (uint32)(int)floorf
        f = pos - tmp
        idx_l, w_l = [], []
        for c in range(8):
            w = torch.ones_like(f[:, 0])
            pl = []
            for d in range(3):
                if (c >> d) & 1:
                    w = w * f[:, d]; pl.append((pg[:, d] + 1) & M32)
                else:
                    w = w * (1.0 - f[:, d]); pl.append(pg[:, d])
            idx_l.append(_level_index(torch.stack(pl, -1), res, size)); w_l.append(w)
        idx_all.append(torch.stack(idx_l, -1)); w_all.append(torch.stack(w_l, -1)); pg_all.append(pg)
    return torch.stack(idx_all, 1), torch.stack(w_all, 1), torch.stack(pg_all, 1)


def hashgrid_encode(x, params, table):
    """(N,3) fp32, flat params (n_params,) -> (N, L*F) fp32.  Autograd gives
    grad_params (scatter-add, tcnn kernel_grid_backward) and grad_x
    (kernel_grid_backward_input: floor() carries no gradient)."""
    L, F = table["n_levels"], table["n_features"]
    x = x.to(torch.float32)
    idx, w, _ = grid_corners(x, table)
    p2 = params.view(-1, F)
    outs = []
    for l in range(L):
        rows = idx[:, l, :] + int(table["offset"][l])          # (N, 8)
        feat = p2[rows]                                         # (N, 8, F)
        acc = torch.zeros(x.shape[0], F, dtype=torch.float32)
        for c in range(8):                                      # corner order as tcnn
            acc = acc + w[:, l, c, None] * feat[:, c, :]
        outs.append(acc)
    return torch.cat(outs, -1)


def init_params(table, seed=1337):
    """U(-1e-4, 1e-4) like tcnn's default initialisation (values are explicit
    inputs to both oracle and kernels; tcnn's pcg32 stream is not reproduced)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(table["n_params"], generator=g, dtype=torch.float32) * 2 - 1) * 1e-4
