"""TEST INFRASTRUCTURE (checker only).  Python face of the marching-cubes oracle (N2).

`marching_cubes(volume, isovalue, truncation)` mirrors the reference's `marching_cubes.marching_cubes`
(external/NumpyMarchingCubes/marching_cubes/src/_mcubes.pyx:20-25; call sites utils/utils.py:78,159): (V,3) float64 vertices in
voxel units and (F,3) uint64 faces.  It runs the C++ restatement oracle/mcubes_oracle.cpp (libmcubes_oracle.so).
`reference_marching_cubes` is the reference's own routine compiled from its sources (oracle/_ref/_mcubes_ref.so, built by
oracle/Makefile where /root/reference exists); None when that binary is absent."""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libmcubes_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libmcubes_oracle.so is missing: run `make -C oracle`")
        _lib = ctypes.CDLL(path)
        _lib.mcubes_oracle.restype = ctypes.c_int
        _lib.mcubes_oracle.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_float, ctypes.c_float,
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_long),
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long)]
        _lib.mcubes_oracle_free.argtypes = [ctypes.c_void_p]
    return _lib


def to_float32_volume(volume):
    """What the reference reads: every element through PyArray_SafeGet<double> and then `float d = ...` (marching_cubes.h:14-17,
    marching_cubes.cpp:82)."""
    v = np.asarray(volume)
    if v.ndim != 3:
        raise RuntimeError("Only three-dimensional arrays are supported.")         # pywrapper.cpp:11-12
    return np.ascontiguousarray(v.astype(np.float64).astype(np.float32))


def marching_cubes(volume, isovalue, truncation, return_soup_count=False):
    lib = _load()
    vol = to_float32_volume(volume)
    pv, pf = ctypes.c_void_p(), ctypes.c_void_p()
    nv, nf, ns = ctypes.c_long(), ctypes.c_long(), ctypes.c_long()
    lib.mcubes_oracle(vol.ctypes.data, vol.shape[0], vol.shape[1], vol.shape[2], float(isovalue), float(truncation),
                      ctypes.byref(pv), ctypes.byref(nv), ctypes.byref(pf), ctypes.byref(nf), ctypes.byref(ns))
    verts = np.ctypeslib.as_array(ctypes.cast(pv, ctypes.POINTER(ctypes.c_float)), (max(nv.value, 1) * 3,))[:nv.value * 3].copy()
    faces = np.ctypeslib.as_array(ctypes.cast(pf, ctypes.POINTER(ctypes.c_uint32)), (max(nf.value, 1) * 3,))[:nf.value * 3].copy()
    lib.mcubes_oracle_free(pv)
    lib.mcubes_oracle_free(pf)
    out = verts.astype(np.float64).reshape(-1, 3), faces.astype(np.uint64).reshape(-1, 3)
    return out + (ns.value,) if return_soup_count else out


def _load_ref():
    d = os.path.join(_HERE, "_ref")
    if not os.path.exists(os.path.join(d, "_mcubes_ref.so")):
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    import _mcubes_ref
    return _mcubes_ref


def reference_marching_cubes(volume, isovalue, truncation):
    """The reference binary (same reshape as _mcubes.pyx:23-24); raises when oracle/_ref was not built."""
    m = _load_ref()
    if m is None:
        raise RuntimeError("oracle/_ref/_mcubes_ref.so is missing (built only where /root/reference exists)")
    v, f = m.marching_cubes(np.asarray(volume), float(np.float32(isovalue)), float(np.float32(truncation)))
    return v.reshape(-1, 3), f.reshape(-1, 3)


def have_reference():
    return _load_ref() is not None
