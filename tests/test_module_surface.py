"""The nn.Module surface the reference relies on (SURVEY.md 8b "Ownership"): `copy.deepcopy` (mipsfusion.py:616,632),
`.share_memory()` (InactiveMap.py:67,107), pickling of the whole object graph under the `spawn` start method
(mipsfusion.py:37,665), strict `state_dict` / `load_state_dict` round trips with the reference's keys and shapes
(`embed_fn.params`, `embedpos_fn.params`, `decoder.*`; Logger.py:33-69).  CPU only: no kernel is called."""
import copy
import io
import multiprocessing as mp
import os
import pickle
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as H

REF_KEYS = ["embedpos_fn.params", "embed_fn.params",
            "decoder.pts_linear.0.weight", "decoder.pts_linear.0.bias", "decoder.pts_linear.2.weight", "decoder.pts_linear.2.bias",
            "decoder.rgb_linear.0.weight", "decoder.rgb_linear.0.bias", "decoder.sdf_linear.0.weight", "decoder.sdf_linear.0.bias",
            "decoder.sdf_linear.2.weight", "decoder.sdf_linear.2.bias"]
REF_SHAPES = {"embedpos_fn.params": (0,), "decoder.pts_linear.0.weight": (128, 51), "decoder.pts_linear.0.bias": (128,),
              "decoder.pts_linear.2.weight": (128, 128), "decoder.pts_linear.2.bias": (128,), "decoder.rgb_linear.0.weight": (3, 115),
              "decoder.rgb_linear.0.bias": (3,), "decoder.sdf_linear.0.weight": (128, 96), "decoder.sdf_linear.0.bias": (128,),
              "decoder.sdf_linear.2.weight": (5, 128), "decoder.sdf_linear.2.bias": (5,)}


def _model(hash_size=12, seed=3):
    import mipsfusion_b200 as mf
    cfg = H.make_config(hash_size)
    bb = torch.tensor(cfg["mapping"]["bound"], dtype=torch.float64)
    nf = torch.tensor(cfg["mapping"]["localMLP_max_len"], dtype=torch.float64)
    torch.manual_seed(seed)
    return cfg, mf.JointEncoding(cfg, bb, nf)


def _same(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k


def test_state_dict_keys_shapes_and_strict_round_trip():
    cfg, m = _model()
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(REF_KEYS)
    for k, shp in REF_SHAPES.items():
        assert tuple(sd[k].shape) == shp, (k, sd[k].shape)
    assert sd["embed_fn.params"].ndim == 1 and sd["embed_fn.params"].dtype == torch.float32
    # a reference-style checkpoint (torch.save of the state_dict, Logger.py:33-69) loads strictly into a fresh module
    buf = io.BytesIO()
    torch.save(sd, buf)
    buf.seek(0)
    _, m2 = _model(seed=4)
    w0 = m2.decoder.pts_linear[0].weight.detach().clone()
    assert not torch.equal(w0, m.decoder.pts_linear[0].weight)
    missing = m2.load_state_dict(torch.load(buf), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    _same(m, m2)
    # recover_initial_param restores the construction-time weights (model/scene_rep.py:50-55)
    g0 = m2.initial_dict["embed_fn.params"].clone()
    with torch.no_grad():
        m2.embed_fn.params.add_(1.0)
    m2.recover_initial_param()
    assert torch.equal(m2.embed_fn.params, g0) and torch.equal(m2.decoder.pts_linear[0].weight, w0)


def test_deepcopy_is_independent():
    _, m = _model()
    c = copy.deepcopy(m)
    _same(m, c)
    with torch.no_grad():
        c.embed_fn.params.add_(1.0)
        c.decoder.pts_linear[0].weight.mul_(2.0)
    assert not torch.equal(m.embed_fn.params, c.embed_fn.params)
    assert not torch.equal(m.decoder.pts_linear[0].weight, c.decoder.pts_linear[0].weight)
    assert c.embed_fn.n_output_dims == 32 and c.embedpos_fn.n_output_dims == 48


def test_share_memory_and_plain_pickle():
    _, m = _model()
    m.share_memory()
    assert m.embed_fn.params.is_shared() and m.decoder.sdf_linear[2].bias.is_shared()
    m2 = pickle.loads(pickle.dumps(m))
    _same(m, m2)
    assert m2.config["grid"]["hash_size"] == m.config["grid"]["hash_size"]


def _child(q, model, marker):
    # runs in a spawned interpreter: the model arrived by pickle (mipsfusion.py:665 passes `self` to mp.Process)
    sd = model.state_dict()
    q.put((sorted(sd.keys()), float(sd["embed_fn.params"].double().sum()), float(sd["decoder.sdf_linear.0.weight"].double().sum()), marker))


def test_pickle_under_spawn():
    _, m = _model()
    m.share_memory()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_child, args=(q, m, 17))
    p.start()
    keys, s_grid, s_w, marker = q.get(timeout=180)
    p.join(timeout=60)
    assert p.exitcode == 0 and marker == 17
    assert keys == sorted(REF_KEYS)
    sd = m.state_dict()
    np.testing.assert_allclose(s_grid, float(sd["embed_fn.params"].double().sum()), rtol=0, atol=0)
    np.testing.assert_allclose(s_w, float(sd["decoder.sdf_linear.0.weight"].double().sum()), rtol=0, atol=0)
