"""CPU checks of the boundary: the shared library loads and exports every symbol the header declares,
and the ctypes prototypes cover exactly that set.  No compute entry point is called."""
import os
import re

import pytest

from mipsfusion_b200 import _lib as L


def header_functions():
    src = open(L.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_prototypes_agree():
    assert header_functions() == sorted(L.PROTOTYPES.keys())


def header_signatures():
    """name -> number of parameters, parsed from the declarations in the header."""
    src = open(L.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(mf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_prototype_arity_matches_header():
    """Every ctypes prototype has as many arguments as the C declaration it binds (guards against signature drift)."""
    sig = header_signatures()
    assert sorted(sig) == sorted(L.PROTOTYPES.keys())
    for name, (_, argtypes) in L.PROTOTYPES.items():
        assert len(argtypes) == sig[name], (name, len(argtypes), sig[name])


def test_library_loads_and_exports_every_symbol():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    dll = L.load_library()
    for name in header_functions():
        assert hasattr(dll, name), name
    assert dll.mf_abi_version() == 2
    assert dll.mf_mlp_prep_size() > L.MF_MLP_PARAMS


def test_meta_matches_oracle_table():
    """Host-only entry point (no CUDA): the level table must equal the oracle's bit for bit."""
    import numpy as np
    from oracle import hashgrid as hg
    from mipsfusion_b200.encodings import grid_meta
    for T in (19, 16, 10):
        t = hg.level_table(T)
        m = grid_meta(T, 16, 2, 16, hg.per_level_scale_of())
        assert np.array_equal(np.asarray(m.scale[:16], dtype=np.float32), t["scale"])
        assert list(m.resolution[:16]) == list(t["resolution"]) and list(m.size[:16]) == list(t["size"])
        assert list(m.offset[:17]) == list(t["offset"])
        dense = [int(r) ** 3 <= int(s) for r, s in zip(t["resolution"], t["size"])]
        assert [h == 0 for h in m.hashed[:16]] == dense


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(L.MipsFusionB200Error):
        L.load_library(str(tmp_path / "nope.so"))


def test_cpu_tensors_are_refused():
    import torch
    with pytest.raises(L.MipsFusionB200Error):
        L.ptr(torch.zeros(3))
