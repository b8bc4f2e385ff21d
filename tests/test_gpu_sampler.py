"""Device-side point-cloud batch sampler against outputs of the unmodified reference
(`sampleTrainingDataPC`, `shortestDistance`, src/dataset.py:72-131; fixture tests/golden/sampler_pc.npz) and the oracle.
Tolerances: gathered rows are bit-exact; displaced rows differ by one fp32 rounding of p + n*off (2e-7); distances come
from the reference's own expansion |x|^2 - 2 p.x + |p|^2 evaluated in fp32: 1e-6 absolute on d^2."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(golden):
    g = golden("sampler_pc.npz")
    return g, torch.from_numpy(g["surf_pts"]).cuda(), torch.from_numpy(g["surf_nrm"]).cuda()


def test_batch_matches_reference_on_its_own_draws(golden):
    from diffudf_b200.dataset import sampleTrainingDataPC
    g, X, N = _cloud(golden)
    n_on, n_off = int(g["n_on"]), int(g["n_off"])
    draws = dict(on_idx=g["on_idx"], far=g["far"].astype(np.float32), near_idx=g["near_idx"], near_off=g["near_off"][:, 0])
    c, n, s = sampleTrainingDataPC(X, N, n_on, n_off, draws=draws)
    assert c.shape == (1, n_on + n_off, 3) and n.shape == c.shape and s.shape == (1, n_on + n_off, 1)
    c, n, s = c.cpu().numpy(), n.cpu().numpy(), s.cpu().numpy()
    n_far = n_off // 2
    assert np.array_equal(c[0, :n_on + n_far], g["coords"][0, :n_on + n_far])          # gathered rows / supplied points: exact
    assert np.abs(c - g["coords"]).max() < 2e-7
    assert np.array_equal(n, g["normals"])
    assert np.array_equal(s[0, :n_on], g["sdf"][0, :n_on]) and np.array_equal(s[0, n_on + n_far:], g["sdf"][0, n_on + n_far:])
    assert np.abs(s[0, n_on:n_on + n_far, 0] ** 2 - g["sdf"][0, n_on:n_on + n_far, 0] ** 2).max() < 1e-6


@pytest.mark.parametrize("nq,nx", [(500, 6000), (1, 1), (1025, 1023), (3000, 2049)])
def test_shortest_distance(nq, nx, golden, oracle):
    from diffudf_b200.dataset import shortestDistance
    g = golden("sampler_pc.npz")
    P = g["sd_queries"] if nq == 500 else np.random.default_rng(nq).uniform(-1, 1, (nq, 3))
    X = g["surf_pts"][:nx].astype(np.float64)
    ref = g["sd64"] if (nq, nx) == (500, 6000) else oracle.shortest_distance(P, X)
    brute = np.sqrt(((P[:, None, :] - X[None, :, :]) ** 2).sum(-1).min(1))
    assert np.abs(ref - brute).max() < 1e-7                                              # the expansion equals the definition in fp64
    out = shortestDistance(torch.from_numpy(P.astype(np.float32)).cuda(), torch.from_numpy(X.astype(np.float32)).cuda()).cpu().numpy()
    assert np.abs(out ** 2 - ref ** 2).max() < 1e-6
    assert shortestDistance(torch.zeros(0, 3).cuda(), torch.from_numpy(X.astype(np.float32)).cuda()).shape == (0,)


def _true_min_distance(P, X):
    """fp64 minimum distance by blocks on the device (the definition, src/dataset.py:72-78 without the expansion)."""
    P64, X64 = P.double(), X.double()
    out = torch.empty(P.shape[0], device=P.device, dtype=torch.float64)
    for i in range(0, P.shape[0], 2048):
        out[i:i + 2048] = torch.cdist(P64[i:i + 2048], X64, compute_mode="donot_use_mm_for_euclid_dist").min(1).values
    return out


@pytest.mark.parametrize("nx", [1, 31, 33, 1024, 1025, 40000, 200000])
def test_cloud_index_is_exact(nx):
    """The box hierarchy prunes with monotone fp32 bounds: it must return the fp32 difference-form distance of the TRUE nearest
    point (relative 1e-6 of the fp64 minimum), for queries inside, on and far outside the cloud, and at every padding edge."""
    from diffudf_b200.dataset import CloudIndex, shortestDistance
    g = torch.Generator(device="cuda").manual_seed(nx)
    u = torch.randn(nx, 3, device="cuda", generator=g)
    X = 0.6 * u / u.norm(dim=1, keepdim=True) + 0.002 * torch.randn(nx, 3, device="cuda", generator=g)     # a noisy sphere: a surface
    if nx >= 1024:
        X[: nx // 4] = torch.rand(nx // 4, 3, device="cuda", generator=g) * 0.5 - 1.0                          # plus a filled corner
    P = torch.cat([torch.rand(3000, 3, device="cuda", generator=g) * 2 - 1,                                    # the sampler's far rows
                   X[torch.randint(0, nx, (500,), device="cuda", generator=g)],                                # ON the cloud
                   X[torch.randint(0, nx, (500,), device="cuda", generator=g)] + 1e-3 * torch.randn(500, 3, device="cuda", generator=g),
                   torch.randn(300, 3, device="cuda", generator=g) * 5.0])                                     # far outside the boxes
    idx = CloudIndex(X)
    d = idx.distance(P)
    ref = _true_min_distance(P, X)
    assert float((d.double() - ref).abs().max()) <= 1e-6 * float(ref.max()) + 1e-7
    assert float(((d.double() - ref).abs() / ref.clamp_min(1e-3)).max()) < 2e-6
    assert float(d[3000:3500].max()) == 0.0
    assert torch.equal(shortestDistance(P, idx), d)
    if 2048 <= nx <= 16 * P.shape[0]:                       # the un-indexed entry point builds a temporary index of its own
        assert torch.equal(shortestDistance(P, X), d)
    assert idx.distance(torch.zeros(0, 3, device="cuda")).shape == (0,)


def test_cloud_index_degenerate_clouds():
    from diffudf_b200.dataset import CloudIndex
    g = torch.Generator(device="cuda").manual_seed(0)
    P = torch.rand(2000, 3, device="cuda", generator=g) * 2 - 1
    same = torch.full((5000, 3), 0.25, device="cuda")                                       # one point 5 000 times
    assert torch.allclose(CloudIndex(same).distance(P), (P - 0.25).norm(dim=1), rtol=1e-6, atol=0)
    line = torch.zeros(7000, 3, device="cuda")
    line[:, 0] = torch.linspace(-1, 1, 7000, device="cuda")                                 # boxes of zero volume
    d = CloudIndex(line).distance(P)
    assert float((d.double() - _true_min_distance(P, line)).abs().max()) < 1e-6


def test_indexed_batch_equals_scanned_batch(golden):
    """PointCloud's batches (far rows through the prebuilt index) against the brute-force scan of the same draws."""
    from diffudf_b200.dataset import CloudIndex, sampleTrainingDataPC
    g, X, N = _cloud(golden)
    idx = CloudIndex(X)
    a = sampleTrainingDataPC(X, N, 999, 1998, seed=11, batch_index=2, index=idx)
    b = sampleTrainingDataPC(X[:2047], N[:2047], 999, 1998, seed=11, batch_index=2)          # < 2 048 points: the tiled scan
    c = sampleTrainingDataPC(X[:2047], N[:2047], 999, 1998, seed=11, batch_index=2, index=CloudIndex(X[:2047]))
    assert torch.equal(b[0], c[0]) and torch.equal(b[1], c[1])
    assert float((b[2] ** 2 - c[2] ** 2).abs().max()) < 1e-6                               # the scan's expansion carries 2e-7 on d^2
    far = a[0][0, 999:999 + 999]
    assert float((a[2][0, 999:999 + 999, 0].double() - _true_min_distance(far, X)).abs().max()) < 1e-6


def test_device_draws_are_valid_deterministic_and_distinct(golden):
    from diffudf_b200.dataset import sampleTrainingDataPC, shortestDistance
    g, X, N = _cloud(golden)
    n_on, n_off = 9990, 19980
    a = sampleTrainingDataPC(X, N, n_on, n_off, seed=7, batch_index=3)
    b = sampleTrainingDataPC(X, N, n_on, n_off, seed=7, batch_index=3)
    c = sampleTrainingDataPC(X, N, n_on, n_off, seed=7, batch_index=4)
    assert all(torch.equal(u, v) for u, v in zip(a, b))
    assert not torch.equal(a[0], c[0])
    coords, normals, sdf = (t[0] for t in a)
    n_far = n_off // 2
    on, far, near = coords[:n_on], coords[n_on:n_on + n_far], coords[n_on + n_far:]
    assert float(shortestDistance(on, X).max()) == 0.0                                    # on rows ARE cloud rows
    assert torch.allclose(normals[:n_on].norm(dim=1), torch.ones(n_on, device="cuda"), atol=1e-4)
    assert float(normals[n_on:].abs().max()) == 0.0 and float(sdf[:n_on].abs().max()) == 0.0
    assert float(far.min()) >= -1.0 and float(far.max()) <= 1.0
    assert float(far.mean(0).abs().max()) < 0.03 and abs(float(far.std()) - 2 / np.sqrt(12)) < 0.01   # uniform on [-1, 1]^3
    off = sdf[n_on + n_far:, 0]
    assert abs(float(off.mean()) - 0.01 * np.sqrt(2 / np.pi)) < 3e-4                       # E|N(0, sigma)| = sigma sqrt(2/pi)
    assert float(shortestDistance(near, X).max()) <= float(off.max()) + 1e-6               # displaced by |off| from a cloud point
    d_far = sdf[n_on:n_on + n_far, 0]
    assert torch.allclose(d_far, shortestDistance(far, X))
    idx_hist = torch.unique(on, dim=0).shape[0]
    assert idx_hist > 0.7 * min(n_on, X.shape[0]) * (1 - np.exp(-n_on / X.shape[0]))       # draws spread over the cloud


def test_training_loop_on_the_device_dataset(golden, weights):
    from diffudf_b200 import SIREN
    from diffudf_b200.dataset import PointCloud
    from diffudf_b200.train import train_model_tanh
    g = golden("sampler_pc.npz")
    ds = PointCloud(g["surf_pts"], g["surf_nrm"], 3000, [0.333, 0.666], 3, "cuda:0", seed=1)
    m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
    m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(weights["init"])
                       for k, v in (("weight", W), ("bias", b))})
    cfg = dict(epochs=3, s1_epochs=2, warmup_epochs=1, warmup_lr=1e-4, lr_s1=1e-5, lr_s2=1e-7, loss_s1_weights=[1e4, 1e4, 1e4, 1e3],
               loss_s2_weights=[1e5, 1e5], alpha=100.0, precision="tc16")
    losses, best, _ = train_model_tanh(ds, m, torch.device("cuda:0"), cfg)
    assert ds.batches_drawn == 9 and best is not None
    assert all(np.isfinite(v).all() for v in losses.values())
    assert losses["sdf_off_surf"][1] < losses["sdf_off_surf"][0]                           # it learns
