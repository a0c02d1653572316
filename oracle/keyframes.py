"""ORACLE (test infrastructure only): CPU restatement of the ray part of the reference's KeyframeSet
(model/keyframeSet.py:76-79 sample_single_keyframe_rays, :170-175 add_keyframe, :386-437 sample_rays_in_submap,
:446-455 sample_rays_in_given_kf) with the python ``random.sample`` draws made explicit arguments.
Pinned against the reference's own class in tests/golden/keyframes.npz (random.sample patched to return the same lists)."""
import torch


def store_keyframe(direction, rgb, depth, rows, cols, W):
    """(H*W,3), (H*W,3), (H*W) + lattice rows / cols -> (n_rays, 7)   [keyframeSet.py:76-79,170-173]"""
    idx = rows * W + cols
    rays = torch.cat([direction.reshape(-1, 3), rgb.reshape(-1, 3), depth.reshape(-1, 1)], dim=-1)
    return rays[idx]


def split_counts(pix_num, related_kf_num):
    first = max(pix_num // related_kf_num, pix_num // 10)                        # :392
    if related_kf_num == 1:
        return first, 0, 0
    if related_kf_num == 2:
        return first, pix_num - first, 0                                        # :413-415
    last = max(pix_num // related_kf_num, pix_num // 5)                          # :402
    return first, pix_num - first - last, last                                  # :409


def sample_rays_in_submap(rays, first_kf_Id, related_kf_ids, pix_num, idx_first, idx_other=None, idx_last=None):
    """rays (num_kf, n_rays, 7); related_kf_ids (n,) int64; idx_* int64 index draws -> sampled_rays, kf_ids, kf_indices"""
    n_rays = rays.shape[1]
    n_rel = related_kf_ids.shape[0]
    first_rays = rays[first_kf_Id].reshape(-1, 7)[idx_first]                     # :394
    first_idx = torch.zeros_like(idx_first); first_ids = torch.ones_like(idx_first) * first_kf_Id
    if n_rel == 1:
        return first_rays, first_ids, first_idx
    other_ids = related_kf_ids[1:-1] if n_rel > 2 else related_kf_ids[1:]        # :408,413
    other_rays = rays[other_ids].reshape(-1, 7)[idx_other]                       # :420
    other_idx = torch.div(idx_other, n_rays, rounding_mode="floor")              # :422
    other_kf = other_ids[other_idx]
    other_idx = other_idx + 1                                                    # :424
    if n_rel > 2:
        last_id = related_kf_ids[-1]
        last_rays = rays[last_id].reshape(-1, 7)[idx_last]                       # :404
        last_idx = torch.ones_like(idx_last) * (n_rel - 1); last_ids = torch.ones_like(idx_last) * last_id
        return (torch.cat([first_rays, other_rays, last_rays], 0), torch.cat([first_ids, other_kf, last_ids], 0),
                torch.cat([first_idx, other_idx, last_idx], 0))                  # :427-429
    return torch.cat([first_rays, other_rays], 0), torch.cat([first_ids, other_kf], 0), torch.cat([first_idx, other_idx], 0)


def sample_rays_in_given_kf(rays, given_kf_ids, idx):
    n_rays = rays.shape[1]
    sampled = rays[given_kf_ids].reshape(-1, 7)[idx]                             # :451
    kf_indices = torch.div(idx, n_rays, rounding_mode="floor")
    return sampled, given_kf_ids[kf_indices], kf_indices


# ---- device draws: keyed Feistel permutation (restates mf_feistel_perm of csrc/optim_sampling_kernels.cu; no reference
# counterpart -- the reference calls python random.sample) ---------------------------------------------------------------
def _mix32(x):
    import numpy as np
    x = np.asarray(x, dtype=np.uint32)
    x = x ^ (x >> np.uint32(16)); x = (x * np.uint32(0x7feb352d)).astype(np.uint32)
    x = x ^ (x >> np.uint32(15)); x = (x * np.uint32(0x846ca68b)).astype(np.uint32)
    return x ^ (x >> np.uint32(16))


def feistel_sample(n, k, seed):
    """k distinct indices of range(n): out[j] = perm(j) (int64 numpy array)."""
    import numpy as np
    b = 1
    while (1 << (2 * b)) < n:
        b += 1
    mask = np.uint32((1 << b) - 1)
    out = np.empty(k, dtype=np.int64)
    with np.errstate(over="ignore"):
        keys = [_mix32(np.uint32((seed + 0x9e3779b9 * (r + 1)) & 0xffffffff)) for r in range(4)]
        x = np.arange(k, dtype=np.uint64)
        todo = np.ones(k, dtype=bool)
        while todo.any():
            xs = x[todo]
            l = ((xs >> np.uint64(b)).astype(np.uint32)) & mask
            r = xs.astype(np.uint32) & mask
            for key in keys:
                t = l ^ (_mix32(r ^ key) & mask)
                l, r = r, t
            xs = (l.astype(np.uint64) << np.uint64(b)) | r.astype(np.uint64)
            x[todo] = xs
            done_now = xs < np.uint64(n)
            idxs = np.nonzero(todo)[0]
            todo[idxs[done_now]] = False
    out[:] = x.astype(np.int64)
    return out
