"""Generates tests/golden/mcubes.npz from the REFERENCE's own marching cubes: its NumpyMarchingCubes C++ sources compiled where
they lie under /root/reference into oracle/_ref/_mcubes_ref.so (recipe: oracle/Makefile) and called on the seeded volumes of
tests/mc_volumes.py.  Run in the build container only:
    make -C oracle && python tests/golden/make_mcubes_golden.py
The outputs (vertices float64, faces uint64) are what `marching_cubes.marching_cubes(volume, isovalue, truncation)` returns
(_mcubes.pyx:20-25).  Vertices are stored as float32 (the reference's values are floats widened to double, marching_cubes.cpp:448-452)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import marching_cubes as omc          # noqa: E402
import mc_volumes                                  # noqa: E402

out = {}
for name, (vol, iso, trunc) in mc_volumes.cases().items():
    v, f = omc.reference_marching_cubes(vol, iso, trunc)
    assert np.array_equal(v.astype(np.float32).astype(np.float64), v)
    out[name + "_volume"] = vol
    out[name + "_args"] = np.array([iso, trunc], np.float64)
    out[name + "_verts"] = v.astype(np.float32)
    out[name + "_faces"] = f.astype(np.uint32)
    print(name, vol.shape, v.shape, f.shape)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mcubes.npz"), **out)
