"""More drop-in coverage of the field-query drivers: gradient-descent refinement of hits (render_st.py:163-172),
value-only grids / gt_mode 'siren' (render_mc.py:314-358, inverses.py:21-22) and evaluate() edge cases."""
import numpy as np
import pytest
import torch

from conftest import rel_max

pytestmark = pytest.mark.gpu


def test_grad_descent_matches_oracle_steps(oracle, weights, cuda_models):
    from diffudf_b200 import render_st
    rng = np.random.default_rng(3)
    t0 = rng.uniform(-0.6, 0.6, (500, 3))
    mask = rng.uniform(size=500) < 0.7
    ref = t0.copy()
    for _ in range(2):                                         # oracle restatement of the loop body
        j = oracle.siren_jet(weights["trained"], ref[mask].astype(np.float32), 1, dtype=np.float32)
        g = j["g"].astype(np.float64)
        g /= np.linalg.norm(g, axis=1, keepdims=True)
        steps = oracle.inverse("tanh", np.abs(j["f"].astype(np.float32))[:, None], 100.0)
        ref[mask] -= g * steps
    render_st.grad_descent(cuda_models["trained"], t0, mask, {"gt_mode": "tanh", "alpha": 100.0}, {"gd_steps": 2}, torch.device("cuda:0"))
    assert np.max(np.abs(t0 - ref)) < 1e-4
    assert np.array_equal(t0[~mask], ref[~mask])


def test_value_grid_and_siren_mode(oracle, weights, cuda_models):
    from diffudf_b200.render_mc import extract_fields, grid_values
    m = cuda_models["init"]
    N = 10
    f = grid_values(m, N).cpu().numpy()
    ref = oracle.siren_jet(weights["init"], oracle.grid_coords(N), 0, dtype=np.float32)["f"].reshape(N, N, N)
    assert f.shape == (N, N, N) and rel_max(f, ref) < 2e-5
    df, vecs = extract_fields(m, torch.Tensor([[]]), N, "siren", torch.device("cuda:0"), 100.0)
    df_ref, vecs_ref = oracle.extract_fields(weights["init"], N, "siren", 100.0)
    assert rel_max(df.cpu().numpy(), df_ref) < 2e-5 and np.max(np.abs(vecs.cpu().numpy() - vecs_ref)) < 2e-4


def test_evaluate_edge_cases(cuda_models):
    from diffudf_b200 import evaluate
    m = cuda_models["init"]
    out = evaluate(m, np.zeros((0, 3), np.float32), device=torch.device("cuda:0"))
    assert out.shape == (0, 1)
    x = np.random.default_rng(0).uniform(-1, 1, (4097, 3)).astype(np.float32)       # one past the reference chunk size
    g32 = np.zeros((4097, 3), np.float32)                                            # caller array of another dtype is filled too
    f = evaluate(m, x, device=torch.device("cuda:0"), gradients=g32)
    g64 = np.zeros((4097, 3))
    f2 = evaluate(m, torch.from_numpy(x).cuda(), device=torch.device("cuda:0"), gradients=g64, max_batch=100000)
    assert np.array_equal(f, f2) and np.allclose(g32, g64, rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        evaluate(m, x, latent_vec=torch.zeros(1, 8), device=torch.device("cuda:0"))
