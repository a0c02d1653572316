"""Oracle: MLP_reg decoder, functional restatement of reference model/decoder.py:53-75.

TEST INFRASTRUCTURE ONLY.  Weight names follow the reference state_dict
(model/decoder.py:32-50): pts_linear.{0,2}, rgb_linear.0, sdf_linear.{0,2}.
"""
import torch
import torch.nn.functional as F

PARAM_SHAPES = [                       # (name, (out, in)) -- model/decoder.py:32-50 with
    ("pts_linear.0", (128, 51)),       # input_ch=32, input_ch_pos=48 (+3 xyz)
    ("pts_linear.2", (128, 128)),
    ("rgb_linear.0", (3, 115)),
    ("sdf_linear.0", (128, 96)),
    ("sdf_linear.2", (5, 128)),
]
N_PARAMS = sum(o * i + o for _, (o, i) in PARAM_SHAPES)     # 36,577


def init_weights(seed=0):
    """nn.Linear default init, same construction order as the reference module."""
    torch.manual_seed(seed)
    w = {}
    for name, (o, i) in PARAM_SHAPES:
        lin = torch.nn.Linear(i, o)
        w[name + ".weight"] = lin.weight.detach().clone()
        w[name + ".bias"] = lin.bias.detach().clone()
    return w


def mlp_reg(w, embed, embed_pos, query_pts, n_hidden_sdf=64, n_class=5):
    """-> (N, 10): rgb_raw(3), sdf(1), entropy(1), prob(5).  model/decoder.py:53-75"""
    e = torch.cat([query_pts, embed_pos], -1)                                        # :54
    h = F.linear(F.relu(F.linear(e, w["pts_linear.0.weight"], w["pts_linear.0.bias"])),
                 w["pts_linear.2.weight"], w["pts_linear.2.bias"])                   # :56
    sdf_emb, rgb_emb = h[:, :n_hidden_sdf], h[:, n_hidden_sdf:]                      # :57-58
    rgb = F.linear(torch.cat([rgb_emb, e], -1), w["rgb_linear.0.weight"], w["rgb_linear.0.bias"])   # :61-62
    z = F.linear(F.relu(F.linear(torch.cat([sdf_emb, embed], -1), w["sdf_linear.0.weight"], w["sdf_linear.0.bias"])),
                 w["sdf_linear.2.weight"], w["sdf_linear.2.bias"])
    prob = torch.softmax(z, -1)                                                      # :65-66
    entropy = -1.0 * torch.sum(prob * torch.log2(prob + 1e-5), -1, keepdim=True)     # :68
    ids = torch.arange(0.0, n_class, 1.0)
    sdf = torch.sum(prob * ids[None], -1, keepdim=True)                              # :71
    sdf = (sdf / (n_class - 1) - 0.5) * 2                                            # :72
    return torch.cat([rgb, sdf, entropy, prob], -1)                                  # :74
