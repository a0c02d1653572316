"""Marching cubes on the device: ``marching_cubes(volume, isovalue, truncation)`` with the signature
and results of the reference's ``marching_cubes.marching_cubes`` (external/NumpyMarchingCubes/marching_cubes/src/_mcubes.pyx:20-25,
pywrapper.cpp:9-54, marching_cubes.cpp:418-462; call sites utils/utils.py:78,159 ``mcubes.marching_cubes(raw.squeeze(), isolevel,
truncation=3.0)``): vertices (V,3) float64 in voxel units, faces (F,3) uint64 -- bit for bit the reference's arrays, including
its vertex order (first-come clusters on the 1e-5 lattice) and its face clean-up.

``marching_cubes_device`` keeps everything on the GPU (float32 vertices, int64 faces) for callers that go on to colour or filter
the mesh there; ``volume`` may be a CUDA tensor (e.g. the blended SDF grid of ``JointSubmapQuery``), so the grid never visits
the host.  There is no CPU fallback."""
import numpy as np
import torch

from . import _lib as L


def _device_volume(volume, device):
    if isinstance(volume, torch.Tensor):
        v = volume
        if v.dim() != 3:
            raise RuntimeError("Only three-dimensional arrays are supported.")          # pywrapper.cpp:11-12
        if v.dtype != torch.float32:                                                    # element -> double -> float (marching_cubes.cpp:82)
            v = v.to(torch.float64).to(torch.float32)
        if not v.is_cuda:
            v = v.to(device if device is not None else "cuda")
        return v.contiguous()
    a = np.asarray(volume)
    if a.ndim != 3:
        raise RuntimeError("Only three-dimensional arrays are supported.")
    a = np.ascontiguousarray(a.astype(np.float64).astype(np.float32))
    return torch.from_numpy(a).to(device if device is not None else "cuda")


def marching_cubes_device(volume, isovalue, truncation, device=None, return_info=False):
    """-> (verts (V,3) float32 CUDA, faces (F,3) int64 CUDA[, info dict])."""
    if not torch.cuda.is_available():
        raise L.MipsFusionB200Error("mipsfusion_b200 kernels need a CUDA device (no CPU fallback)")
    vol = _device_volume(volume, device)
    dev = vol.device
    nx, ny, nz = (int(s) for s in vol.shape)
    iso, trunc = float(np.float32(isovalue)), float(np.float32(truncation))             # `float isovalue, float truncation` (_mcubes.pyx:20)
    with torch.cuda.device(dev):
        lib = L.lib()
        ws = torch.empty(max(int(lib.mf_mcubes_count_workspace_size(nx, ny, nz)), 256), device=dev, dtype=torch.uint8)
        n_tris_d = torch.zeros(1, device=dev, dtype=torch.int64)
        L.call("mf_mcubes_count", L.ptr(vol), nx, ny, nz, iso, trunc, L.ptr(ws), L.ptr(n_tris_d), L.stream())
        n_tris = int(n_tris_d.item())
        counts = torch.zeros(3, device=dev, dtype=torch.int64)
        if n_tris == 0:
            verts = torch.empty(0, 3, device=dev, dtype=torch.float32)
            faces = torch.empty(0, 3, device=dev, dtype=torch.int64)
            info = {"soup_triangles": 0, "rounds": 0}
            return (verts, faces, info) if return_info else (verts, faces)
        size = int(lib.mf_mcubes_mesh_workspace_size(n_tris))
        if size <= 0:
            raise L.MipsFusionB200Error(f"marching cubes: {n_tris} triangles exceed the 32-bit vertex index range")
        mws = torch.empty(size, device=dev, dtype=torch.uint8)
        verts = torch.empty(3 * n_tris, 3, device=dev, dtype=torch.float32)
        faces = torch.empty(n_tris, 3, device=dev, dtype=torch.int32)
        L.call("mf_mcubes_mesh", L.ptr(ws), nx, ny, nz, iso, n_tris, L.ptr(mws), L.ptr(verts), L.ptr(faces), L.ptr(counts), L.stream())
        nv, nf, rounds = (int(c) for c in counts.tolist())
    verts, faces = verts[:nv], faces[:nf].to(torch.int64)
    if return_info:
        return verts, faces, {"soup_triangles": n_tris, "rounds": rounds}
    return verts, faces


def marching_cubes(volume, isovalue, truncation):
    """Drop-in for the reference's ``marching_cubes.marching_cubes``: numpy in (or a tensor), numpy out."""
    verts, faces = marching_cubes_device(volume, isovalue, truncation)
    return verts.cpu().numpy().astype(np.float64), faces.cpu().numpy().astype(np.uint64)
