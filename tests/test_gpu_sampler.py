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
