"""Field extraction for marching cubes: drop-in for extract_fields / get_mesh_sdf's grid evaluation
(/root/reference/src/render_mc.py:20-101, :314-358).  The grid coordinates are generated inside the
query kernel (no N^3 x 7 sample tensor, no host float64 temporaries).  extract_mesh_CAP (:201-256), the CAP-UDF marching
cubes that consumes these fields, runs on the device as well (dudf_cap_mesh); the MeshUDF mesher stays a downstream consumer
(SURVEY.md §8f)."""
import numpy as np
import torch

from .engine import Q_ABS_INV_TANH, Q_NEG_NORMALIZE
from .inverses import inverse_torch


def extract_fields(decoder, latent_vec, N, gt_mode, device, alpha, first=0, count=None, out=None):
    """Returns df (N,N,N) fp32 = inverse(gt_mode, |f|) and vecs (N,N,N,3) fp32 = -normalize(grad f) on `device`.
    Where grad f is exactly zero the sign-aligned top Hessian eigenvector is used (render_mc.py:77-93).
    `first`/`count` restrict the evaluation to a range of the flat index (slab sharding); the
    returned tensors are then flat: df (count,), vecs (count,3).  `out` = (df, vecs) preallocated flat tensors of `count` rows
    to write into (the sharded driver passes slices of the gathered output)."""
    if latent_vec is not None and torch.is_tensor(latent_vec) and latent_vec.numel() != 0:
        raise ValueError("extract_fields: latent conditioning is not supported")
    eng = decoder._engine_synced()
    total = N ** 3
    full = count is None
    count = total - first if count is None else count
    flags = Q_NEG_NORMALIZE | (Q_ABS_INV_TANH if gt_mode == "tanh" else 0)
    df, vecs, _ = eng.query_grid(N, first, count, decoder.precision, flags, alpha, want_vecs=True, out=out)
    if gt_mode != "tanh":
        if out is not None:
            df.copy_(inverse_torch(gt_mode, df.abs(), alpha))
        else:
            df = inverse_torch(gt_mode, df.abs(), alpha)
    # zero-gradient fallback (rare): Hessian eigenvector, aligned with the (zero) gradient like the reference
    small = torch.linalg.norm(vecs, dim=-1) < 0.04
    if bool(small.any()):
        idx = torch.nonzero(small).reshape(-1)
        from .render_st import grid_points
        pts = grid_points(N, idx + first, vecs.device)
        _, _, H, _ = eng.query(pts, 2, "fp32")
        vecs[idx] = eng.field_vectors(vecs[idx].contiguous(), H)
    if full and first == 0:
        return df.reshape(N, N, N), vecs.reshape(N, N, N, 3)
    return df, vecs


def grid_values(decoder, N, device=None, first=0, count=None):
    """Raw field values on the grid (the evaluation loop of get_mesh_sdf, render_mc.py:331-345)."""
    eng = decoder._engine_synced()
    count = N ** 3 - first if count is None else count
    f, _, _ = eng.query_grid(N, first, count, decoder.precision, 0, 0.0, want_vecs=False)
    return f.reshape(N, N, N) if (first == 0 and count == N ** 3) else f


class TriangleSoup:
    """What extract_mesh_CAP returns where the reference returns trimesh.Trimesh(v_all, t_all, process=False): vertices
    (3T, 3) float64, faces (T, 3) int64 (no vertex is shared between triangles), and export() for .obj / .ply."""

    def __init__(self, vertices, faces):
        self.vertices, self.faces = vertices, faces

    def export(self, path):
        v, f = self.vertices, self.faces
        with open(path, "w") as fh:
            if str(path).lower().endswith(".ply"):
                fh.write(f"ply\nformat ascii 1.0\nelement vertex {len(v)}\nproperty double x\nproperty double y\nproperty double z\n"
                         f"element face {len(f)}\nproperty list uchar int vertex_indices\nend_header\n")
                fh.writelines(f"{p[0]!r} {p[1]!r} {p[2]!r}\n" for p in v.tolist())
                fh.writelines(f"3 {t[0]} {t[1]} {t[2]}\n" for t in f.tolist())
            else:
                fh.writelines(f"v {p[0]!r} {p[1]!r} {p[2]!r}\n" for p in v.tolist())
                fh.writelines(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n" for t in f.tolist())


def extract_mesh_MESHUDF(df_values, normals, device=None, smooth_borders=False, **kwargs):
    """Drop-in for the numerical part of extract_mesh_MESHUDF (src/render_mc.py:103-134): clamp, MeshUDF marching cubes with
    spacing 2 / (N - 1), avg_thresh 1.05, max_thresh 1.75, shift by -1 — through the C++ mesher (marching_cubes.udf_mc_lewiner,
    identical arrays to the reference's Cython module).  Returns (TriangleSoup, None) where the reference returns
    (pred_mesh, trimesh.Trimesh): its trimesh clean-up loops and the border smoothing (:136-199) are third-party mesh hygiene and
    are not rebuilt (smooth_borders=True raises)."""
    if smooth_borders:
        raise NotImplementedError("border smoothing / trimesh clean-up are downstream mesh hygiene (SURVEY.md 2 #9), not part of this path")
    from .marching_cubes import meshudf_from_fields
    verts, faces = meshudf_from_fields(df_values, normals)
    return TriangleSoup(verts, faces), None


def cap_triangles(ndf, grad, resolution, threshold=0.008, device=None):
    """CAP-UDF marching cubes on the device (dudf_cap_mesh; src/render_mc.py:201-256): ndf (N,N,N), grad (N,N,N,3) as numpy arrays
    or tensors -> (T,3,3) float64 CUDA tensor of triangles in the reference's cell order."""
    import ctypes

    from . import _lib
    dev = torch.device(device) if device is not None else (ndf.device if torch.is_tensor(ndf) and ndf.is_cuda else torch.device("cuda:0"))
    df = torch.as_tensor(ndf).to(device=dev, dtype=torch.float32).contiguous()
    g = torch.as_tensor(grad).to(device=dev, dtype=torch.float32).contiguous()
    N = int(resolution)
    if tuple(df.shape) != (N, N, N) or tuple(g.shape) != (N, N, N, 3):
        raise ValueError(f"extract_mesh_CAP: expected ndf {(N, N, N)} and grad {(N, N, N, 3)}, got {tuple(df.shape)} and {tuple(g.shape)}")
    L = _lib.lib()
    h = _cap_context(dev)
    n = ctypes.c_int64(0)
    with torch.cuda.device(dev):
        _lib.check(L.dudf_cap_mesh(h, df.data_ptr(), g.data_ptr(), N, float(threshold), None, 0, ctypes.byref(n), _lib.current_stream()),
                   "dudf_cap_mesh")
        tris = torch.empty(int(n.value), 3, 3, device=dev, dtype=torch.float64)
        _lib.check(L.dudf_cap_mesh(h, df.data_ptr(), g.data_ptr(), N, float(threshold), _lib.ptr(tris), int(n.value), ctypes.byref(n),
                                   _lib.current_stream()), "dudf_cap_mesh")
    return tris


_CAP_CTX = {}


def _cap_context(dev):
    import ctypes

    from . import _lib
    key = (dev.type, dev.index)
    if key not in _CAP_CTX:
        h = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().dudf_create(8, 30.0, 30.0, ctypes.byref(h)), "dudf_create")
        _CAP_CTX[key] = h
    return _CAP_CTX[key]


def extract_mesh_CAP(ndf, grad, resolution):
    """Drop-in for src/render_mc.py:201-256: same arguments (numpy arrays as generate_mc.py:35 passes them, or tensors), returns an
    object with .vertices / .faces / .export() like the trimesh the reference builds with process=False.  Raises ValueError when no
    cell is triangulated (the reference fails in np.concatenate of an empty list)."""
    tris = cap_triangles(ndf, grad, resolution)
    if tris.shape[0] == 0:
        raise ValueError("need at least one array to concatenate")
    v = tris.reshape(-1, 3).cpu().numpy()
    f = np.arange(v.shape[0], dtype=np.int64).reshape(-1, 3)
    return TriangleSoup(v, f)
