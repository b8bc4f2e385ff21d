"""Host-side logic that needs no GPU: module construction / state_dict contract, schedule, batch layout,
sharding, inverses, error behaviour."""
import numpy as np
import pytest
import torch


def test_siren_module_contract():
    from diffudf_b200 import SIREN
    torch.manual_seed(123)
    m = SIREN(3, 1, [256] * 8, w0=30)
    sd = m.state_dict()
    assert list(sd.keys()) == [f"net.{i}.0.{k}" for i in range(9) for k in ("weight", "bias")]
    assert sd["net.0.0.weight"].shape == (256, 3) and sd["net.8.0.weight"].shape == (1, 256)
    assert sum(v.numel() for v in sd.values()) == 461825
    assert float(sd["net.0.0.weight"].abs().max()) <= 1 / 3 + 1e-7
    assert float(sd["net.3.0.weight"].abs().max()) <= np.sqrt(6 / 256) / 30 + 1e-7
    assert m.ww == 30 and SIREN(3, 1, [256], w0=30, ww=20).ww == 20


def test_init_matches_reference_rng_stream(oracle):
    """Same seed -> same weights as the reference constructor (fixture weights_init.npz was made by it)."""
    import os
    from conftest import GOLDEN
    from diffudf_b200 import SIREN
    torch.manual_seed(123)
    m = SIREN(3, 1, [256] * 8, w0=30)
    ref = oracle.load_params(os.path.join(GOLDEN, "weights_init.npz"))
    for i, (W, b) in enumerate(ref):
        assert np.array_equal(m.net[i][0].weight.detach().numpy(), W)
        assert np.array_equal(m.net[i][0].bias.detach().numpy(), b)


def test_unsupported_configs_raise_without_fallback():
    from diffudf_b200 import SIREN, gradient
    for bad in ([128], [256, 128], []):
        with pytest.raises(ValueError):
            SIREN(3, 1, bad)
    with pytest.raises(ValueError):
        SIREN(2, 1, [256])
    with pytest.raises(ValueError):
        SIREN(3, 1, [256], activation="relu")
    m = SIREN(3, 1, [256, 256])
    with pytest.raises(RuntimeError):
        m(torch.zeros(5, 3))                       # CPU: the product has no CPU path
    x = torch.zeros(3, 3, requires_grad=True)
    with pytest.raises(RuntimeError):
        gradient(x.sum(-1, keepdim=True), x)


def test_lr_schedule_matches_reference_formula():
    from diffudf_b200.train import lr_for_epoch
    cfg = dict(epochs=3000, s1_epochs=2000, warmup_epochs=1000, warmup_lr=1e-4, lr_s1=1e-5, lr_s2=1e-7)
    assert lr_for_epoch(0, **cfg) == 1e-4 and lr_for_epoch(999, **cfg) == 1e-4
    assert lr_for_epoch(1000, **cfg) == 1e-5 and lr_for_epoch(1999, **cfg) == 1e-5
    for e in (2000, 2500, 2999):
        assert lr_for_epoch(e, **cfg) == 0.5 * (np.cos(e / 1000 * np.pi) + 1) * 1e-7


def test_inverses_match_reference_semantics():
    from diffudf_b200.inverses import inverse, inverse_torch
    f = np.array([0.0, 0.004, 0.01, 0.02, 0.5], np.float32)
    out = inverse("tanh", f, 100)
    assert np.allclose(out, [0.0, np.sqrt(0.004 / 100), 0.01, 0.02, 0.5])
    assert np.allclose(inverse("siren", np.array([-1.0, 0.0, 0.3]), 100), [0.01, 0.01, 0.3])
    assert np.allclose(inverse("squared", np.array([4.0, -1.0]), 100, 0.01), [0.2, 0.001])
    assert torch.allclose(inverse_torch("tanh", torch.from_numpy(f), 100.0), torch.from_numpy(out))
    with np.errstate(invalid="ignore"):
        assert np.isnan(inverse("tanh", np.array([-0.001]), 100, 0)[0])      # render_pc.py:51 relies on this NaN


def test_synthetic_batch_layout():
    from diffudf_b200 import synthetic
    shape = synthetic.make_shape(0)
    rng = np.random.default_rng(0)
    sp, sn = shape.sample_surface(5000, rng)
    assert np.max(shape.udf(sp)) < 1e-6 and np.allclose(np.linalg.norm(sn, axis=1), 1, atol=1e-5)
    x, n, d = synthetic.make_batch(shape, sp, sn, 30000, (0.333, 0.666), rng)
    assert x.shape == (1, 29970, 3) and n.shape == (1, 29970, 3) and d.shape == (1, 29970, 1)
    assert np.all(d[0, :9990, 0] == 0) and np.all(n[0, 9990:] == 0) and np.all(d[0, 9990:, 0] > 0)
    assert x.dtype == np.float32 and np.abs(x[0, 9990:19980]).max() <= 1


def test_shard_batch_keeps_groups_balanced():
    from diffudf_b200.parallel import shard_batch, shard_range
    P, n_on, n_far = 29970, 9990, 9990
    x = np.arange(P * 3, dtype=np.float32).reshape(P, 3)
    nrm = np.zeros((P, 3), np.float32)
    d = np.concatenate([np.zeros(n_on), np.ones(P - n_on)]).astype(np.float32)
    seen = []
    for r in range(8):
        xr, nr, dr, on_r = shard_batch(x, nrm, d, n_on, n_far, r, 8)
        assert np.all(dr[:on_r] == 0) and np.all(dr[on_r:] == 1)
        assert abs(on_r - n_on / 8) <= 1 and abs(xr.shape[0] - P / 8) <= 3
        seen.append(xr[:, 0])
    assert np.array_equal(np.sort(np.concatenate(seen)), x[:, 0])
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_eigh3_restatement_matches_numpy():
    """The Jacobi solver used on the device is restated here in numpy to pin its ordering / triangle conventions."""
    rng = np.random.default_rng(0)
    A = rng.normal(size=(200, 3, 3))
    A = A + np.swapaxes(A, 1, 2)
    from oracle import dudf_oracle as O
    lam, V = O.eig_top(A)
    assert np.all(np.diff(lam, axis=1) >= 0)
    recon = np.einsum("pik,pk,pjk->pij", V, lam, V)
    assert np.max(np.abs(recon - A)) < 1e-12
    B = A.copy()
    B[:, 0, 1] += 5.0                                  # upper triangle must be ignored (LAPACK 'L')
    lam2, _ = O.eig_top(B)
    assert np.allclose(lam, lam2)


def test_triangle_soup_export_roundtrip(tmp_path):
    """What extract_mesh_CAP returns in place of trimesh.Trimesh(..., process=False): .vertices / .faces / .export (render_mc.py:253)."""
    import numpy as np
    from diffudf_b200.render_mc import TriangleSoup
    v = np.random.default_rng(0).normal(size=(6, 3))
    f = np.arange(6, dtype=np.int64).reshape(2, 3)
    m = TriangleSoup(v, f)
    for name in ("m.obj", "m.ply"):
        p = tmp_path / name
        m.export(str(p))
        txt = p.read_text().splitlines()
        if name.endswith(".obj"):
            vs = np.array([[float(t) for t in l.split()[1:]] for l in txt if l.startswith("v ")])
            fs = np.array([[int(t) - 1 for t in l.split()[1:]] for l in txt if l.startswith("f ")])
        else:
            h = txt.index("end_header")
            assert f"element vertex {len(v)}" in txt and f"element face {len(f)}" in txt
            vs = np.array([[float(t) for t in l.split()] for l in txt[h + 1:h + 1 + len(v)]])
            fs = np.array([[int(t) for t in l.split()[1:]] for l in txt[h + 1 + len(v):]])
        assert np.array_equal(vs, v) and np.array_equal(fs, f)          # repr() round-trips float64 exactly


def test_cloud_index_buffer_sizes():
    """dudf_cloud_index_bytes (host arithmetic only): header + points padded to whole 32-leaf blocks + one box pair per node of every
    level up to the first level with at most 32 boxes; out-of-range sizes answer -1."""
    from diffudf_b200 import _lib
    L = _lib.lib()
    assert L.dudf_cloud_index_bytes(0) == -1 and L.dudf_cloud_index_bytes(-5) == -1 and L.dudf_cloud_index_bytes((1 << 30) + 1) == -1
    def expect(n):
        pts = ((n + 31) // 32 + 31) // 32 * 1024
        total, count = 64 + pts * 16, n
        while True:
            count = (count + 31) // 32
            total += (count + 31) // 32 * 32 * 32
            if count <= 32:
                return total
    prev = 0
    for n in (1, 31, 32, 33, 1024, 1025, 32768, 32769, 200000, 1 << 20, (1 << 20) + 1, 1 << 30):
        b = L.dudf_cloud_index_bytes(n)
        assert b == expect(n), (n, b, expect(n))
        assert b >= prev and b % 16 == 0
        prev = b
