"""Drop-in for the geometry of /root/reference/src/preprocess_mesh.py: `normalizeMesh` (:5-15) and the surface cloud of
`preprocessMesh` (:29-40, Open3D `mesh.sample_points_uniformly(n, use_triangle_normal=True)`), on arrays instead of Open3D
objects; the samples are drawn on the device by `dudf_mesh_sample_surface`.  File IO beyond a plain `.obj` reader stays with
the caller (Open3D / trimesh are third-party consumers)."""
import numpy as np
import torch

from . import _lib


def read_obj(path):
    """Vertices (n, 3) float64 and triangles (m, 3) int64 of a Wavefront .obj (polygons are fanned)."""
    V, F = [], []
    with open(path) as fh:
        for line in fh:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                V.append([float(t) for t in p[1:4]])
            elif p[0] == "f":
                idx = [int(t.split("/")[0]) for t in p[1:]]
                idx = [i - 1 if i > 0 else len(V) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    F.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(V, np.float64), np.asarray(F, np.int64)


def normalizeMesh(vertices):
    """Returns (normalised vertices, 4x4 transform S @ T): centre = mean of the vertices, scale 1 / (1.1 max|coord|)."""
    V = np.asarray(vertices, np.float64)
    c = V.mean(0)
    T = np.eye(4)
    T[:3, 3] = -c
    Vc = V - c
    m = np.abs(Vc).max()
    S = np.eye(4)
    S[:3, :3] *= 1.0 / (m + m * 0.1)
    return Vc * S[0, 0], S @ T


def sample_points_uniformly(vertices, faces, number_of_points, device, seed=0, draws=None):
    """Area-weighted surface samples with triangle normals -> (points (n, 3), normals (n, 3)) fp32 CUDA tensors.
    `draws` (n, 3) = (u_triangle, r1, r2) in [0, 1) replaces the Philox draws (parity tests)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("diffudf_b200.preprocess_mesh.sample_points_uniformly draws on a CUDA (sm_100) device; there is no CPU fallback")
    V = np.asarray(vertices, np.float64)
    F = np.asarray(faces, np.int64)
    tri = V[F]
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    cdf = torch.from_numpy((np.cumsum(area) / area.sum()).astype(np.float32)).to(dev)
    T = torch.from_numpy(tri.astype(np.float32)).to(dev).contiguous()
    n = int(number_of_points)
    pts = torch.empty(n, 3, device=dev, dtype=torch.float32)
    nrm = torch.empty(n, 3, device=dev, dtype=torch.float32)
    dr = None if draws is None else torch.as_tensor(draws).to(device=dev, dtype=torch.float32).contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dudf_mesh_sample_surface(T.data_ptr(), cdf.data_ptr(), T.shape[0], n, int(seed) & (2 ** 64 - 1), _lib.ptr(dr),
                                                       pts.data_ptr(), nrm.data_ptr(), _lib.current_stream()), "dudf_mesh_sample_surface")
    return pts, nrm
