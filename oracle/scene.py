"""Oracle: JointEncoding (scene representation) -- query, z-sampling, SDF->weights
rendering and the rgb / depth / sdf / free-space losses.

TEST INFRASTRUCTURE ONLY.  Restates reference model/scene_rep.py:58-238 and
helper_functions/utils.py:21-111 in plain torch (CPU fp32; the coordinate
normalisation is fp64 exactly as in the reference, whose bound tensors are
float64 -- mipsfusion.py:94-96).  Random jitter ``u`` is an explicit input
(the reference draws it with CPU torch.rand, scene_rep.py:176).
"""
import math
import torch
import torch.nn.functional as F

from . import hashgrid as hg
from .frequency import frequency_encode
from .decoder import mlp_reg, init_weights


def default_config():
    """Hot-path keys of configs/FastCaMo-synth/FastCaMo-synth.yaml (:80-118)."""
    return {
        "grid": {"enc": "HashGrid", "tcnn_encoding": True, "hash_size": 19, "voxel_sdf": 0.04,
                 "use_bound_normalize": True},
        "pos": {"enc": "Frequency", "n_bins": 8},
        "cam": {"H": 480, "W": 640, "fx": 320.0, "fy": 320.0, "cx": 319.5, "cy": 239.5,
                "crop_edge": 10, "near": 0, "far": 5, "depth_trunc": 100.0},
        "training": {"rgb_weight": 1.0, "depth_weight": 0.0, "sdf_weight": 1000, "fs_weight": 10,
                     "n_samples_d": 50, "range_d": 0.2, "n_range_d": 25, "n_samples": 75, "perturb": 1,
                     "norm_factor": 1.0, "trunc": 0.1, "rgb_missing": 0.0},
        "data": {"sc_factor": 1},
        "mapping": {"bound": [[-0.6, 2.95], [0.5, 7.05], [-1.15, 3.05]],
                    "localMLP_max_len": [7.0, 7.0, 7.0], "lr_embed": 0.01, "lr_decoder": 0.01},
    }


class OracleField:
    """State of one submap: grid params + MLP weights + config (JointEncoding.__init__,
    model/scene_rep.py:11-45)."""

    def __init__(self, config, bound_box=None, coords_norm_factor=None, grid_seed=1337, mlp_seed=0):
        self.config = config
        bb = config["mapping"]["bound"] if bound_box is None else bound_box
        self.bounding_box = torch.as_tensor(bb, dtype=torch.float64)                 # mipsfusion.py:94
        nf = config["mapping"]["localMLP_max_len"] if coords_norm_factor is None else coords_norm_factor
        self.coords_norm_factor = torch.as_tensor(nf, dtype=torch.float64)           # mipsfusion.py:96
        self.table = hg.level_table(config["grid"]["hash_size"])
        self.n_bins = config["pos"]["n_bins"]
        self.grid = hg.init_params(self.table, grid_seed).requires_grad_(True)
        self.w = {k: v.requires_grad_(True) for k, v in init_weights(mlp_seed).items()}

    def parameters(self):
        return [self.grid] + list(self.w.values())

    # -- model/scene_rep.py:118-128
    def query_color_sdf(self, query_points):
        x = torch.reshape(query_points, [-1, query_points.shape[-1]]) / self.config["training"]["norm_factor"]
        embed = hg.hashgrid_encode(x.to(torch.float32), self.grid, self.table)       # tcnn casts to fp32
        embed_pos = frequency_encode(x.to(torch.float32), self.n_bins)
        x32 = x.to(torch.float32)
        return mlp_reg(self.w, embed, embed_pos, x32)

    def query_sdf(self, p):                       # :106-107
        return self.query_color_sdf(p)[..., 3:4]

    def query_color(self, p):                     # :109-110
        return torch.sigmoid(self.query_color_sdf(p)[..., :3])

    def query_sdf_entropy_prob(self, p):          # :113-114
        return self.query_color_sdf(p)[..., 3:]

    # -- model/scene_rep.py:134-146
    def normalize(self, inputs_flat):
        if self.config["grid"]["tcnn_encoding"]:
            if self.config["grid"]["use_bound_normalize"]:
                return (inputs_flat - self.bounding_box[:, 0]) / (self.bounding_box[:, 1] - self.bounding_box[:, 0])
            return (inputs_flat + self.coords_norm_factor) / (2 * self.coords_norm_factor)
        return inputs_flat

    def run_network(self, inputs):
        inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
        out = self.query_color_sdf(self.normalize(inputs_flat))
        return torch.reshape(out, list(inputs.shape[:-1]) + [out.shape[-1]])

    # -- model/scene_rep.py:58-78
    def sdf2weights(self, sdf, z_vals):
        tr = self.config["training"]["trunc"]
        weights = torch.sigmoid(sdf / tr) * torch.sigmoid(-sdf / tr)
        signs = sdf[:, 1:] * sdf[:, :-1]
        mask = torch.where(signs < 0.0, torch.ones_like(signs), torch.zeros_like(signs))
        inds = torch.argmax(mask, axis=1)[..., None]
        z_min = torch.gather(z_vals, 1, inds)
        mask = torch.where(z_vals < z_min + self.config["data"]["sc_factor"] * tr,
                           torch.ones_like(z_vals), torch.zeros_like(z_vals))
        weights = weights * mask
        return weights / (torch.sum(weights, axis=-1, keepdims=True) + 1e-8), inds[..., 0]

    # -- model/scene_rep.py:81-103
    def raw2outputs(self, raw, z_vals):
        rgb = torch.sigmoid(raw[..., :3])
        weights, inds = self.sdf2weights(raw[..., 3], z_vals)
        rgb_map = torch.sum(weights[..., None] * rgb, -2)
        depth_map = torch.sum(weights * z_vals, -1)
        depth_var = torch.sum(weights * torch.square(z_vals - depth_map.unsqueeze(-1)), dim=-1)
        disp_map = 1.0 / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / torch.sum(weights, -1))
        acc_map = torch.sum(weights, -1)
        return rgb_map, disp_map, acc_map, weights, depth_map, depth_var, inds

    # -- model/scene_rep.py:157-176 ; u replaces torch.rand(z_vals.shape)
    def sample_z(self, n_rays, target_d, u):
        tr = self.config["training"]; cam = self.config["cam"]
        if target_d is not None:
            z_samples = torch.linspace(-tr["range_d"], tr["range_d"], steps=tr["n_range_d"]).to(target_d)
            z_samples = z_samples[None, :].repeat(n_rays, 1) + target_d
            z_samples[target_d.squeeze(-1) <= 0] = torch.linspace(cam["near"], cam["far"], steps=tr["n_range_d"]).to(target_d)
            if tr["n_samples_d"] > 0:
                z_vals = torch.linspace(cam["near"], cam["far"], tr["n_samples_d"])[None, :].repeat(n_rays, 1)
                z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
            else:
                z_vals = z_samples
        else:
            z_vals = torch.linspace(cam["near"], cam["far"], tr["n_samples"])[None, :].repeat(n_rays, 1)
        if tr["perturb"] > 0.0:
            mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
            upper = torch.cat([mids, z_vals[..., -1:]], -1)
            lower = torch.cat([z_vals[..., :1], mids], -1)
            z_vals = lower + (upper - lower) * u
        return z_vals

    # -- model/scene_rep.py:153-187
    def render_rays(self, rays_o, rays_d, target_d=None, u=None):
        z_vals = self.sample_z(rays_o.shape[0], target_d, u)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        raw = self.run_network(pts)
        rgb_map, disp_map, acc_map, weights, depth_map, depth_var, inds = self.raw2outputs(raw, z_vals)
        return {"rgb": rgb_map, "depth": depth_map, "disp_map": disp_map, "acc_map": acc_map,
                "depth_var": depth_var, "z_vals": z_vals, "raw": raw, "inds": inds, "weights": weights}

    # -- model/scene_rep.py:190-238
    def forward(self, rays_o, rays_d, target_rgb, target_d, u, EMD_w=0.01):
        rend = self.render_rays(rays_o, rays_d, target_d=target_d, u=u)
        valid = (target_d.squeeze(-1) > 0.0) * (target_d.squeeze(-1) < self.config["cam"]["depth_trunc"])
        rgb_weight = valid.clone().unsqueeze(-1)
        rgb_weight[rgb_weight == 0] = self.config["training"]["rgb_missing"]        # stays bool (:212-213)
        rgb_loss = F.mse_loss(rend["rgb"] * rgb_weight, target_rgb * rgb_weight)
        psnr = -10.0 * torch.log(rgb_loss) / math.log(10.0)
        depth_loss = F.mse_loss(rend["depth"].squeeze()[valid], target_d.squeeze(-1)[valid])
        truncation = self.config["training"]["trunc"] * self.config["data"]["sc_factor"]
        fs_loss, sdf_loss, counts = get_sdf_loss(rend["z_vals"], target_d, rend["raw"][..., 3], rend["raw"][..., 5:],
                                                 truncation, 5, EMD_w)
        return {"rgb": rend["rgb"], "depth": rend["depth"], "rgb_loss": rgb_loss, "depth_loss": depth_loss,
                "sdf_loss": sdf_loss, "fs_loss": fs_loss, "psnr": psnr,
                "z_vals": rend["z_vals"], "raw": rend["raw"], "inds": rend["inds"], "counts": counts}

    def total_loss(self, ret):                    # mipsfusion.py:142-152
        t = self.config["training"]
        return t["rgb_weight"] * ret["rgb_loss"] + t["depth_weight"] * ret["depth_loss"] + \
            t["sdf_weight"] * ret["sdf_loss"] + t["fs_weight"] * ret["fs_loss"]


def get_masks(z_vals, target_d, truncation):
    """helper_functions/utils.py:21-49 (+ the integer counts, returned for parity)."""
    front_mask = torch.where(z_vals < (target_d - truncation), torch.ones_like(z_vals), torch.zeros_like(z_vals))
    back_mask = torch.where(z_vals > (target_d + truncation), torch.ones_like(z_vals), torch.zeros_like(z_vals))
    depth_mask = torch.where(target_d > 0.0, torch.ones_like(target_d), torch.zeros_like(target_d))
    sdf_mask = (1.0 - front_mask) * (1.0 - back_mask) * depth_mask
    num_fs = torch.count_nonzero(front_mask)
    num_sdf = torch.count_nonzero(sdf_mask)
    num = num_sdf + num_fs
    return front_mask, sdf_mask, 1.0 - num_fs / num, 1.0 - num_sdf / num, (int(num_fs), int(num_sdf))


def get_sdf_loss(z_vals, target_d, predicted_sdf, sdf_prob, truncation, cate_num=5, EMD_w=0.01):
    """helper_functions/utils.py:71-111 (loss_type 'l2')."""
    max_id = cate_num - 1
    front_mask, sdf_mask, fs_weight, sdf_weight, counts = get_masks(z_vals, target_d, truncation)
    index_range = torch.arange(0, cate_num).to(sdf_prob)
    fs_loss = F.mse_loss(predicted_sdf * front_mask, torch.ones_like(predicted_sdf) * front_mask) * fs_weight
    sdf_loss = F.mse_loss((z_vals + predicted_sdf * truncation) * sdf_mask, target_d * sdf_mask) * sdf_weight
    if EMD_w > 0:
        fs_all = sdf_prob * (max_id - index_range).to(sdf_prob) * front_mask[..., None]
        fs1 = torch.mean(torch.sum(fs_all, dim=-1)) / 250
        gt_cls = (((target_d - z_vals) + truncation) / (2.0 * truncation)) * max_id
        sdf_all = torch.abs(gt_cls[:, :, None] - index_range[None, None, :]) * sdf_mask[..., None] * sdf_prob
        sdf1 = torch.mean(torch.sum(sdf_all, dim=-1)) / 5000
        fs_loss = fs_loss + fs1 * EMD_w
        sdf_loss = sdf_loss + sdf1 * EMD_w
    return fs_loss, sdf_loss, counts
