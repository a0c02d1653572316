"""Per-submap mesh extraction on the device: ``extract_mesh2`` and ``getVoxels`` of reference utils/utils.py:12-34,121-207 (called
by ``Logger.extract_all_mesh``, Logger.py:173-189, with ``model.query_sdf`` / ``model.query_color``), the caller of the in-tree
marching cubes.  Same arguments; the grid is generated, transformed, normalised, queried and polygonised on the device slab by
slab (the reference materialises the whole grid, queries 64 Ki-point chunks and moves every chunk to the host), then the vertex
colours are queried.  Returns the arrays the reference hands to ``trimesh.Trimesh`` (trimesh itself is absent offline) and, if
``mesh_savepath`` is given, writes them as a binary PLY."""
import os

import numpy as np
import torch

from .marching_cubes import marching_cubes_device


def getVoxels(x_max, x_min, y_max, y_min, z_max, z_min, voxel_size=None, resolution=None):
    """utils/utils.py:12-34."""
    x_max, x_min, y_max, y_min, z_max, z_min = (float(v) for v in (x_max, x_min, y_max, y_min, z_max, z_min))
    if voxel_size is not None:
        Nx = round((x_max - x_min) / voxel_size + 0.0005)
        Ny = round((y_max - y_min) / voxel_size + 0.0005)
        Nz = round((z_max - z_min) / voxel_size + 0.0005)
        return torch.linspace(x_min, x_max, Nx + 1), torch.linspace(y_min, y_max, Ny + 1), torch.linspace(z_min, z_max, Nz + 1)
    return torch.linspace(x_min, x_max, resolution), torch.linspace(y_min, y_max, resolution), torch.linspace(z_min, z_max, resolution)


def _transform_points(pts, mat):
    """helper_functions/geometry_helper.py:76-82."""
    return torch.transpose(mat[:3, :3] @ torch.transpose(pts, 0, 1) + mat[:3, 3:], 0, 1)


def write_ply(path, vertices, triangles, colors=None):
    """Binary little-endian PLY (what ``mesh.export('*.ply')`` produces: float vertices, uchar RGBA colours, int32 faces)."""
    v = np.asarray(vertices, dtype="<f4")
    f = np.asarray(triangles).astype("<i4")
    hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {v.shape[0]}", "property float x", "property float y", "property float z"]
    if colors is not None:
        hdr += ["property uchar red", "property uchar green", "property uchar blue", "property uchar alpha"]
    hdr += [f"element face {f.shape[0]}", "property list uchar int vertex_indices", "end_header"]
    os.makedirs(os.path.split(path)[0] or ".", exist_ok=True)
    with open(path, "wb") as fh:
        fh.write(("\n".join(hdr) + "\n").encode())
        if colors is not None:
            c = np.asarray(colors, dtype=np.float64)
            rgba = np.concatenate([np.clip(np.round(c[:, :3] * 255), 0, 255), np.full((c.shape[0], 1), 255.0)], 1).astype(np.uint8)
            rec = np.empty(v.shape[0], dtype=[("p", "<f4", 3), ("c", "u1", 4)])
            rec["p"], rec["c"] = v, rgba
            fh.write(rec.tobytes())
        else:
            fh.write(v.tobytes())
        rec = np.empty(f.shape[0], dtype=[("n", "u1"), ("i", "<i4", 3)])
        rec["n"], rec["i"] = 3, f
        fh.write(rec.tobytes())


@torch.no_grad()
def extract_mesh2(query_fn, first_kf_c2w, config, bounding_box, marching_cube_bound=None, color_func=None, voxel_size=None,
                  resolution=None, isolevel=0.0, scene_name='', mesh_savepath='', slab_points=1 << 24):
    """utils/utils.py:121-207.  -> dict(vertices (V,3) float64 numpy, triangles (F,3) uint64 numpy, colors (V,3) float32 numpy | None,
    sdf_volume (nx,ny,nz) float32 CUDA)."""
    dev = bounding_box.device
    if dev.type != "cuda":
        raise RuntimeError("mipsfusion_b200.extract_mesh2 needs CUDA tensors (no CPU fallback)")
    if marching_cube_bound is None:
        marching_cube_bound = bounding_box
    x_min, y_min, z_min = marching_cube_bound[:, 0]
    x_max, y_max, z_max = marching_cube_bound[:, 1]
    tx, ty, tz = getVoxels(x_max, x_min, y_max, y_min, z_max, z_min, voxel_size, resolution)
    nx, ny, nz = tx.shape[0], ty.shape[0], tz.shape[0]
    first_kf_w2l = first_kf_c2w.to(dev).inverse()
    txd, tyd, tzd = tx.to(dev), ty.to(dev), tz.to(dev)
    tcnn = bool(config['grid']['tcnn_encoding'])
    vol = torch.empty(nx, ny, nz, device=dev, dtype=torch.float32)
    planes = max(1, int(slab_points) // max(1, ny * nz))
    for i0 in range(0, nx, planes):                                              # slabs of x planes: same values as the full meshgrid
        q = torch.stack(torch.meshgrid(txd[i0:i0 + planes], tyd, tzd, indexing='ij'), -1).to(torch.float32)
        flat_world = q.reshape(-1, 3).to(bounding_box[:, 0])
        flat = _transform_points(flat_world.to(first_kf_w2l), first_kf_w2l)      # world -> submap frame (:143-144)
        if tcnn:
            flat = (flat - bounding_box[:, 0]) / (bounding_box[:, 1] - bounding_box[:, 0])
        vol[i0:i0 + planes] = query_fn(flat[:, None, :]).to(torch.float32).reshape(q.shape[0], ny, nz)
    verts_d, tris_d = marching_cubes_device(vol, isolevel, 3.0)                  # mcubes.marching_cubes(raw.squeeze(), isolevel, truncation=3.0)
    vertices = verts_d.cpu().numpy().astype(np.float64)
    triangles = tris_d.cpu().numpy().astype(np.uint64)
    # normalize vertex positions, rescale and translate, metric units (:166-181)
    vertices[:, :3] /= np.array([[nx - 1, ny - 1, nz - 1]])
    txn, tyn, tzn = tx.numpy(), ty.numpy(), tz.numpy()
    scale = np.array([txn[-1] - txn[0], tyn[-1] - tyn[0], tzn[-1] - tzn[0]])
    offset = np.array([txn[0], tyn[0], tzn[0]])
    vertices[:, :3] = scale[np.newaxis, :] * vertices[:, :3] + offset
    vertices[:, :3] = vertices[:, :3] / config['data']['sc_factor'] - config['data']['translation']
    color = None
    if color_func is not None and vertices.shape[0] > 0:
        vert_flat = torch.from_numpy(vertices).to(bounding_box)
        if tcnn:                                                                  # (:183-187: local frame, NOT normalised, as written)
            vert_flat = _transform_points(vert_flat.to(first_kf_w2l), first_kf_w2l)
        color = color_func(vert_flat[:, None, :]).to(torch.float32).reshape(vert_flat.shape[0], -1).cpu().numpy()
    if mesh_savepath:
        write_ply(mesh_savepath, vertices, triangles, color)
    return {"vertices": vertices, "triangles": triangles, "colors": color, "sdf_volume": vol}
