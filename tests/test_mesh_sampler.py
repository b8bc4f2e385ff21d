"""Mesh half of the batch sampler (src/dataset.py:14-70, src/preprocess_mesh.py:5-15,29-40).
CPU: the fp64 oracle (oracle/mesh_oracle.py) against the committed beetle fixture and against a dense sampling of the
triangles.  GPU: dudf_mesh_distance / dudf_sample_batch_mesh / dudf_mesh_sample_surface against the oracle.
Open3D is absent: its sign convention / random stream are unpinned (the oracle header says so)."""
import numpy as np
import pytest

from oracle import mesh_oracle as M


def test_oracle_distance_against_dense_sampling():
    rng = np.random.default_rng(0)
    tri = rng.normal(size=(6, 3, 3))
    P = rng.normal(size=(40, 3)) * 1.5
    d = M.point_triangle_distance(P, tri)
    u = np.linspace(0, 1, 300)
    U, V = np.meshgrid(u, u)
    m = U + V <= 1
    s, t = U[m], V[m]
    best = np.full(len(P), np.inf)
    for k in range(len(tri)):
        pts = tri[k, 0] + s[:, None] * (tri[k, 1] - tri[k, 0]) + t[:, None] * (tri[k, 2] - tri[k, 0])
        best = np.minimum(best, np.linalg.norm(P[:, None] - pts[None], axis=-1).min(1))
    assert (d <= best + 1e-12).all() and np.abs(d - best).max() < 1e-4


def test_oracle_reproduces_fixture(golden):
    g = golden("beetle_mesh.npz")
    d = M.point_triangle_distance(g["q"][:500], g["tri"])
    assert np.allclose(d, g["d"][:500], rtol=1e-12, atol=1e-15)
    pts, nrm, t = M.sample_surface(g["tri"], g["draws"][:200])
    assert np.array_equal(t, g["surf_tri"][:200]) and np.allclose(pts, g["surf_pts"][:200], atol=1e-15)
    # the samples lie on their triangles, the normals are unit and orthogonal to the triangle edges
    assert np.abs(M.point_triangle_distance(pts, g["tri"])).max() < 1e-7
    tri = g["tri"][t].astype(np.float64)
    assert np.abs(np.einsum("nk,nk->n", nrm, tri[:, 1] - tri[:, 0])).max() < 1e-7 and np.allclose(np.linalg.norm(nrm, axis=1), 1)


def test_normalize_matches_preprocess_formula():
    V = np.random.default_rng(1).normal(size=(50, 3)) * [3, 1, 2] + [5, -2, 1]
    Vn = M.normalize_vertices(V)
    assert np.allclose(Vn.mean(0), 0, atol=1e-12) and np.isclose(np.abs(Vn).max(), 1 / 1.1)


@pytest.mark.gpu
def test_mesh_distance_matches_oracle(golden):
    import torch
    from diffudf_b200.dataset import meshDistance
    g = golden("beetle_mesh.npz")
    q = torch.from_numpy(g["q"]).cuda()
    d = meshDistance(q, torch.from_numpy(g["tri"]).cuda()).cpu().numpy().astype(np.float64)
    ref = g["d"]
    err = np.abs(d - ref)
    print(f"mesh distance: max abs err {err.max():.2e}, max rel err (d > 1e-3) {np.max(err[ref > 1e-3] / ref[ref > 1e-3]):.2e}")
    # fp32 coordinates: absolute 1e-6 everywhere (distances down to 1e-7 next to the surface), relative 1e-5 away from it
    assert err.max() < 1e-6
    assert np.max(err[ref > 1e-3] / ref[ref > 1e-3]) < 1e-5
    for n in (0, 1, 255, 257):          # ragged query counts; one triangle; (vertices, faces) form
        dd = meshDistance(q[:n], torch.from_numpy(g["tri"]).cuda()).cpu().numpy()
        assert dd.shape == (n,) and np.allclose(dd, ref[:n], atol=1e-6)
    one = meshDistance(q[:100], g["tri"][:1]).cpu().numpy()
    assert np.allclose(one, M.point_triangle_distance(g["q"][:100], g["tri"][:1]), atol=1e-6)


@pytest.mark.gpu
def test_mesh_surface_sampler_matches_oracle(golden):
    import torch
    from diffudf_b200.preprocess_mesh import normalizeMesh, sample_points_uniformly
    g = golden("beetle_mesh.npz")
    tri = g["tri"].astype(np.float64)
    V = tri.reshape(-1, 3)
    F = np.arange(len(V)).reshape(-1, 3)
    pts, nrm = sample_points_uniformly(V, F, len(g["draws"]), "cuda:0", draws=g["draws"])
    pts, nrm = pts.cpu().numpy(), nrm.cpu().numpy()
    same = np.linalg.norm(pts - g["surf_pts"], axis=1) < 1e-6          # a draw within fp32 rounding of a CDF step may pick the neighbour
    assert same.mean() > 0.995
    assert np.abs(nrm[same] - g["surf_nrm"][same]).max() < 1e-5
    assert M.point_triangle_distance(pts, tri).max() < 1e-6             # every sample lies on the mesh
    p2, n2 = sample_points_uniformly(V, F, 20000, "cuda:0", seed=3)    # Philox draws: on the mesh, area-weighted
    assert M.point_triangle_distance(p2[:2000].cpu().numpy(), tri).max() < 1e-6
    Vn, T = normalizeMesh(V * 3 + 1)
    assert np.isclose(np.abs(Vn).max(), 1 / 1.1) and np.allclose((np.c_[V * 3 + 1, np.ones(len(V))] @ T.T)[:, :3], Vn)


@pytest.mark.gpu
def test_mesh_batch_matches_oracle(golden):
    import torch
    from diffudf_b200.dataset import PointCloud, sampleTrainingData
    g = golden("beetle_mesh.npz")
    rng = np.random.default_rng(11)
    n_on, n_off = 333, 666
    draws = dict(on_idx=rng.integers(0, 2000, n_on), far=rng.uniform(-1, 1, (n_off // 2, 3)).astype(np.float32),
                 near_idx=rng.integers(0, n_on, n_off - n_off // 2), near_off=rng.normal(0, 0.01, n_off - n_off // 2).astype(np.float32))
    sp, sn = g["surf_pts"].astype(np.float32), g["surf_nrm"].astype(np.float32)
    x, n, d = sampleTrainingData(torch.from_numpy(sp).cuda(), torch.from_numpy(sn).cuda(), n_on, n_off, g["tri"], draws=draws)
    xr, nr, dr = M.sample_training_data(sp, sn, g["tri"], n_on, n_off, draws["on_idx"], draws["far"], draws["near_idx"], draws["near_off"])
    assert x.shape == (1, n_on + n_off, 3) and d.shape == (1, n_on + n_off, 1)
    assert np.abs(x[0].cpu().numpy() - xr).max() < 1e-6 and np.abs(n[0].cpu().numpy() - nr).max() < 1e-6
    assert np.abs(d[0, :, 0].cpu().numpy() - dr).max() < 2e-6
    ds = PointCloud(sp, sn, 3000, [0.333, 0.666], 2, "cuda:0", seed=1, triangles=g["tri"])
    batches = list(ds)
    assert len(batches) == 2 and not ds.onlyPCloud
    xb, nb, db = batches[0]
    off = xb[0, ds.samplesOnSurface:].cpu().numpy()
    assert np.abs(db[0, ds.samplesOnSurface:, 0].cpu().numpy() - M.point_triangle_distance(off, g["tri"])).max() < 2e-6
    assert float(db[0, :ds.samplesOnSurface].abs().max()) == 0.0
