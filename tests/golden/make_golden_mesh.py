"""Fixture for the mesh half of the batch sampler:

    python tests/golden/make_golden_mesh.py

data/beetle/beetle.obj of the reference (a DATA file) -> normalised as src/preprocess_mesh.py:5-15 -> the triangle list as
fp32 (2 053 x 3 x 3), plus the fp64 oracle's answers (oracle/mesh_oracle.py) on seeded queries: 4 000 point-triangle
distances (uniform domain points and points 1e-2 / 1e-4 off the surface) and 2 000 area-weighted surface samples for given
draws.  Open3D is absent (parity with its sign / random stream unpinned, see oracle/mesh_oracle.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import mesh_oracle as M  # noqa: E402


def main():
    V, F = M.read_obj("/root/reference/data/beetle/beetle.obj")
    V = M.normalize_vertices(V)
    tri = V[F].astype(np.float32)
    rng = np.random.default_rng(7)
    draws = rng.uniform(size=(2000, 3)).astype(np.float32)
    pts, nrm, t = M.sample_surface(tri, draws)
    q_far = rng.uniform(-1, 1, (2000, 3))
    q_near = pts[:1000] + nrm[:1000] * rng.normal(0, 0.01, (1000, 1))
    q_tiny = pts[1000:] + nrm[1000:] * rng.normal(0, 1e-4, (1000, 1))
    q = np.concatenate([q_far, q_near, q_tiny]).astype(np.float32)
    d = M.point_triangle_distance(q, tri)
    path = os.path.join(HERE, "beetle_mesh.npz")
    np.savez_compressed(path, tri=tri, draws=draws, surf_pts=pts, surf_nrm=nrm, surf_tri=t, q=q, d=d)
    print("wrote", path, os.path.getsize(path), "faces", len(F), "d range", d.min(), d.max())


if __name__ == "__main__":
    main()
