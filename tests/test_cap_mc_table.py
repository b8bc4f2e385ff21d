"""The marching-cubes case table of the CAP-UDF mesher (tools/gen_mc_table.py -> oracle/mc_table.py, csrc/dudf_mc_table.h) and the
CPU restatement of extract_mesh_CAP (oracle/cap_mc.py; reference src/render_mc.py:201-256).  PyMCubes' own table is not
available (parity unpinned), so the table is held to the properties any correct marching-cubes table has."""
import importlib.util
import os
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_mc_table", os.path.join(ROOT, "tools", "gen_mc_table.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_committed_tables_are_the_generators_output():
    from oracle import mc_table
    g = _gen()
    table = g.build()
    assert [list(map(tuple, t)) for t in mc_table.TRIS] == [list(t) for t in table]
    assert [tuple(e) for e in mc_table.EDGE_CORNERS] == g.EDGES
    hdr = open(os.path.join(ROOT, "diffudf_b200", "csrc", "dudf_mc_table.h")).read()
    for m in (1, 37, 105, 254):
        flat = [e for tri in table[m] for e in tri] + [15] * (3 * mc_table.MAX_TRIS - 3 * len(table[m]))
        assert "{" + ", ".join(map(str, flat)) + "}," + f"  // {m}\n" in hdr


def test_every_case_uses_exactly_the_sign_changing_edges():
    from oracle.mc_table import EDGE_CORNERS, MAX_TRIS, TRIS
    assert len(TRIS) == 256 and MAX_TRIS == 5 and TRIS[0] == [] and TRIS[255] == []
    for m in range(256):
        crossing = {e for e, (a, b) in enumerate(EDGE_CORNERS) if ((m >> a) & 1) != ((m >> b) & 1)}
        used = {e for tri in TRIS[m] for e in tri}
        assert used == crossing, m
        # the polygons of a case are fans over closed loops: n crossing edges in k loops give n - 2k triangles
        assert all(len(set(tri)) == 3 for tri in TRIS[m])
        # complementary configurations cut the same edges
        assert {e for tri in TRIS[255 - m] for e in tri} == crossing


def _mesh_of_signed_volume(vol):
    from oracle.cap_mc import cell_triangles
    tris = []
    n = vol.shape[0]
    for i in range(n - 1):
        for j in range(n - 1):
            for k in range(n - 1):
                res = vol[i:i + 2, j:j + 2, k:k + 2].astype(np.float64)
                if res.min() < 0 <= res.max():
                    tris.append(cell_triangles(res) + np.array([i, j, k], dtype=np.float64))
    return np.concatenate(tris)


def test_mesh_of_a_consistently_signed_field_is_closed_and_oriented():
    """Every undirected triangle edge that is not on the volume boundary is shared by exactly two triangles, once in each direction
    (vertices on a shared cube edge are computed from the same two corner values, so they match bit for bit)."""
    rng = np.random.default_rng(7)
    n = 14
    x = np.stack(np.meshgrid(*[np.linspace(-1, 1, n)] * 3, indexing="ij"), -1)
    for trial in range(3):
        a, ph = rng.normal(size=(4, 3)) * 2.5, rng.uniform(0, 6.28, 4)
        vol = sum(np.sin(x @ a[q] + ph[q]) for q in range(4)) + 0.1           # plenty of ambiguous faces
        tris = _mesh_of_signed_volume(vol)
        assert len(tris) > 500
        directed = Counter()
        for t in tris:
            for q in range(3):
                directed[(tuple(t[q]), tuple(t[(q + 1) % 3]))] += 1
        for (p, q), c in directed.items():
            assert c == 1, "a directed edge appears twice: inconsistent orientation"
            on_boundary = any(p[d] == q[d] and p[d] in (0.0, float(n - 1)) for d in range(3))
            if not on_boundary:
                assert directed.get((q, p), 0) == 1, "open edge inside the volume"


def test_oracle_cap_on_an_analytic_sphere():
    from oracle.cap_mc import extract_mesh_CAP
    N = 40
    x = np.stack(np.meshgrid(*[np.linspace(-1, 1, N, dtype=np.float32)] * 3, indexing="ij"), -1)
    r = np.linalg.norm(x, axis=-1)
    ndf = np.abs(r - 0.5).astype(np.float32)
    grad = (-np.sign(r - 0.5)[..., None] * x / np.maximum(r, 1e-9)[..., None]).astype(np.float32)   # -normalize(grad |r - 0.5|)
    # the reference's fixed 0.008 threshold is meant for voxels of that size (N >= 256): at N = 40 (voxel 0.051) it skips most
    # surface cells, exactly like the reference would; a threshold of one voxel keeps them all
    assert 0 < extract_mesh_CAP(ndf, grad, N).shape[0] < extract_mesh_CAP(ndf, grad, N, threshold=0.06).shape[0]
    tris = extract_mesh_CAP(ndf, grad, N, threshold=0.06)
    assert tris.shape[0] > 1000
    rr = np.linalg.norm(tris.reshape(-1, 3), axis=1)
    assert np.abs(rr - 0.5).max() < 2e-3                        # linear interpolation of |r - 0.5| is exact along rays, ~h^2 off them
    # the area of the mesh approaches the sphere's
    e1, e2 = tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    area = 0.5 * np.linalg.norm(np.cross(e1, e2), axis=1).sum()
    assert abs(area - 4 * np.pi * 0.25) < 0.02 * 4 * np.pi * 0.25
    # threshold: with a threshold below the smallest corner distance of every surface cell nothing is triangulated
    assert extract_mesh_CAP(ndf, grad, N, threshold=-1.0).shape[0] == 0
