"""Size-independent properties at BASELINE.json's full sizes (the oracle cannot run there in seconds):
* a query is a pure per-point map: splitting, offsetting or permuting the batch changes nothing, bit for bit;
* the dense grid query equals the point query on the explicit coordinates;
* the tensor-core path stays within its tolerance of the fp32 path on a 2 M-point sample of the 512^3 grid;
* sphere tracing on 1024 x 1024 rays and NDF projection of 2 M points terminate with sane invariants."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "tc16"])
def test_query_is_a_pure_per_point_map(precision, cuda_models):
    eng = cuda_models["trained"]._engine_synced()
    g = torch.Generator(device="cuda").manual_seed(0)
    P = 2_000_003 if precision == "tc16" else 300_007
    x = torch.rand(P, 3, device="cuda", generator=g) * 2 - 1
    f, gr, H, _ = eng.query(x, 2, precision)
    cut = P // 3 + 5
    f1, g1, H1, _ = eng.query(x[:cut].contiguous(), 2, precision)
    f2, g2, H2, _ = eng.query(x[cut:].contiguous(), 2, precision)
    assert torch.equal(f, torch.cat([f1, f2])) and torch.equal(gr, torch.cat([g1, g2])) and torch.equal(H, torch.cat([H1, H2]))
    perm = torch.randperm(P, device="cuda", generator=g)
    fp, gp, _, _ = eng.query(x[perm].contiguous(), 1, precision)
    f1o, g1o, _, _ = eng.query(x, 1, precision)
    assert torch.equal(fp, f1o[perm]) and torch.equal(gp, g1o[perm])
    assert bool(torch.isfinite(H).all())
    assert float((H - H.transpose(1, 2)).abs().max()) == 0.0            # symmetric by construction


def test_grid_512_tc16_matches_fp32_on_a_sample_and_point_queries(oracle, cuda_models):
    m = cuda_models["trained"]
    eng = m._engine_synced()
    N = 512
    total = N ** 3
    df, vecs, _ = eng.query_grid(N, 0, total, "tc16", 3, 100.0)           # the full 134 M-point grid, config 3
    assert df.shape == (total,) and vecs.shape == (total, 3)
    assert bool(torch.isfinite(df).all()) and bool((df >= 0).all())
    nrm = torch.linalg.norm(vecs, dim=1)
    assert float((nrm - 1).abs().max()) < 1e-4
    # a contiguous 2 M-point slab recomputed in fp32, and the same slab as explicit points in tc16 (bit-identical)
    first, cnt = 77 * N * N + 12345, 2_000_000
    df32, v32, _ = eng.query_grid(N, first, cnt, "fp32", 3, 100.0)
    err = float((df[first:first + cnt] - df32).abs().max() / df32.abs().max())
    assert err < 4e-3, err
    cosang = (vecs[first:first + cnt] * v32).sum(1)
    assert float(cosang.median()) > 0.99999 and float((cosang < 0.99).float().mean()) < 2e-3
    idx = torch.arange(first, first + 300_000, device="cuda")
    from diffudf_b200.render_st import grid_points
    f_pts, g_pts, _, _ = eng.query(grid_points(N, idx, "cuda"), 1, "tc16", flags=3, alpha=100.0)
    assert torch.equal(f_pts, df[first:first + 300_000]) and torch.equal(g_pts, vecs[first:first + 300_000])


def test_sphere_tracing_1024_and_projection_2m(cuda_models):
    from diffudf_b200 import render_st
    from diffudf_b200.render_pc import Sampler
    m = cuda_models["trained"]
    m.precision = "tc16"
    try:
        R = 1024
        cam = np.array([0.8939, 0.7, 2.86]) * 0.45
        u, v = np.meshgrid(np.linspace(-0.6, 0.6, R), np.linspace(-0.6, 0.6, R))
        d = np.stack([u.ravel(), v.ravel(), -np.ones(R * R)], 1)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        fwd = -cam / np.linalg.norm(cam)
        right = np.cross(fwd, [0, 1.0, 0]); right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        rays = d[:, :1] * right + d[:, 1:2] * up - d[:, 2:3] * fwd
        t0 = np.tile(cam, (R * R, 1)) + rays * 0.35
        mask = np.ones(R * R, dtype=bool)
        hits = render_st.propagate_rays(m, rays, t0, mask, {"gt_mode": "tanh", "alpha": 100.0},
                                        {"surface_threshold": 0.004, "max_iterations": 100}, torch.device("cuda:0"))
        assert hits.shape == (R * R,) and 0.02 < hits.mean() < 0.98
        assert np.all(np.abs(t0[hits]) < 1.0)
        attrs = render_st.hit_attributes(m, torch.from_numpy(t0[hits]).cuda(), torch.from_numpy(rays[hits]).cuda(), "mean")
        n = attrs["normals"]
        assert float((torch.linalg.norm(n, dim=1) - 1).abs().max()) < 1e-4
        assert float((n * torch.from_numpy(rays[hits]).cuda().float()).sum(1).max()) <= 1e-6      # sign-fixed against the rays
        assert bool(torch.isfinite(attrs["mean"]).all())
        # config 5: 2 M seeds, 3 projection steps
        s = Sampler(decoder=m, device="cuda:0")
        g = torch.Generator(device="cuda").manual_seed(1)
        seeds = (torch.rand(2_000_000, 3, device="cuda", generator=g, dtype=torch.float64) * 2 - 1)
        pts, steps, grad, H = s.project(seeds, "tanh", 100.0, 3)
        ok = torch.isfinite(steps) & (steps < 0.007) & ((pts.abs() <= 1).all(1))
        assert 0.05 < float(ok.float().mean()) <= 1.0
        f_final, _, _, _ = m._engine_synced().query(pts[ok].float().contiguous(), 0, "fp32")
        assert float(f_final.abs().median()) < 5e-3                   # accepted points sit on the zero level set
    finally:
        m.precision = "fp32"
        for p in m.parameters():
            p.requires_grad_(True)
