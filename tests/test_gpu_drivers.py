"""dudf_march_rays / dudf_project_points (the device-resident loops of src/render_st.py:136-172 and src/render_pc.py:43-53) against
a step-by-step restatement of the same loops on torch tensors around dudf_query_points — the array expressions of the reference,
one query per iteration — on identical weights and rays.  Every ray is independent, so hit masks, surviving rays and the float64
positions must agree exactly; the golden-vector tests of the drivers against the reference itself are in test_gpu_render*.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(weights, tag="trained", precision="fp32"):
    from diffudf_b200 import SIREN
    m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
    m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(weights[tag]) for k, v in (("weight", W), ("bias", b))})
    m = m.cuda()
    m.precision = precision
    return m


def _inverse(gt_mode, f, alpha, min_step):
    from diffudf_b200.inverses import inverse_torch
    return inverse_torch(gt_mode, f, alpha, min_step=min_step)


def _march_restated(model, rays_d, t0_d, idx, gt_mode, alpha, thr, max_it):
    eng = model._engine_synced()
    hits = torch.zeros(t0_d.shape[0], dtype=torch.bool, device=t0_d.device)
    it = nq = 0
    while idx.numel() > 0 and it < max_it:
        x = t0_d[idx].to(torch.float32).contiguous()
        f, _, _, _ = eng.query(x, 0, model.precision)
        nq += x.shape[0]
        steps = _inverse(gt_mode, f.abs(), alpha, 0.01)
        pos = t0_d[idx] + rays_d[idx] * steps.to(torch.float64)[:, None]
        t0_d[idx] = pos
        below = (f < thr) if gt_mode == "siren" else (steps.abs() < thr)
        inside = ((pos > -1).all(dim=1)) & ((pos < 1).all(dim=1))
        hits[idx] |= below & inside
        idx = idx[(~below) & inside]
        it += 1
    return hits, idx, nq


def _rays(n, seed):
    rng = np.random.default_rng(seed)
    org = np.array([0.8939, 0.7, 2.86]) * 0.6
    tgt = rng.uniform(-0.6, 0.6, (n, 3))
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    t0 = org + d * rng.uniform(0.3, 0.9, (n, 1))
    return torch.from_numpy(d).cuda(), torch.from_numpy(t0).cuda()


@pytest.mark.parametrize("precision", ["fp32", "tc16"])
@pytest.mark.parametrize("gt_mode,thr,max_it", [("tanh", 0.004, 100), ("tanh", 0.004, 7), ("siren", 0.01, 40), ("squared", 0.02, 25)])
def test_march_rays_equals_the_restated_loop(precision, gt_mode, thr, max_it, weights):
    from diffudf_b200 import render_st
    m = _model(weights, precision=precision)
    rays, t0 = _rays(5003, 3)
    idx = torch.nonzero(torch.from_numpy(np.random.default_rng(1).uniform(size=5003) < 0.9).cuda()).reshape(-1)
    ta, tb = t0.clone(), t0.clone()
    h0, i0, q0 = _march_restated(m, rays, ta, idx.clone(), gt_mode, 100.0, thr, max_it)
    h1, i1, q1 = render_st._march(m, rays, tb, idx.clone(), gt_mode, 100.0, thr, max_it)
    assert q0 == q1 and q0 > 0
    assert torch.equal(h0, h1)
    assert torch.equal(i0, i1)
    assert torch.equal(ta, tb)
    untouched = torch.ones(5003, dtype=torch.bool, device="cuda")
    untouched[idx] = False
    assert torch.equal(tb[untouched], t0[untouched])


def test_march_rays_degenerate_inputs(weights):
    m = _model(weights)
    eng = m._engine_synced()
    rays, t0 = _rays(64, 0)
    none = torch.zeros(64, dtype=torch.uint8, device="cuda")
    hit = torch.zeros(64, dtype=torch.uint8, device="cuda")
    before = t0.clone()
    assert eng.march_rays(t0, rays, none, hit, "tanh", 100.0, 0.004, 100) == 0          # no active ray: nothing moves
    assert torch.equal(t0, before) and int(hit.sum()) == 0 and int(none.sum()) == 0
    e = torch.zeros(0, 3, dtype=torch.float64, device="cuda")
    z = torch.zeros(0, dtype=torch.uint8, device="cuda")
    assert eng.march_rays(e, e.clone(), z, z.clone(), "tanh", 100.0, 0.004, 100) == 0  # empty ray set
    allr = torch.ones(64, dtype=torch.uint8, device="cuda")
    assert eng.march_rays(t0, rays, allr, hit, "tanh", 100.0, 0.004, 0) == 0            # max_it = 0: every ray still marching
    assert int(allr.sum()) == 64
    with pytest.raises(KeyError):
        eng.march_rays(t0, rays, allr, hit, "nope", 100.0, 0.004, 10)
    with pytest.raises(RuntimeError):
        eng.march_rays(t0.float(), rays, allr, hit, "tanh", 100.0, 0.004, 10)


@pytest.mark.parametrize("precision", ["fp32", "tc16"])
@pytest.mark.parametrize("gt_mode,num_steps", [("tanh", 3), ("tanh", 1), ("siren", 4), ("squared", 2)])
def test_project_points_equals_the_restated_loop(precision, gt_mode, num_steps, weights):
    from diffudf_b200.render_pc import Sampler
    m = _model(weights, precision=precision)
    eng = m._engine_synced()
    smp = Sampler(decoder=m, device=0)
    x0 = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, (4097, 3))).cuda()
    s = x0.clone()
    for step in range(num_steps):                    # the reference's float64 array arithmetic on the widened fp32 results
        f, g, H, _ = eng.query(s.to(torch.float32).contiguous(), 2 if step == num_steps - 1 else 1, precision)
        st = _inverse(gt_mode, f.to(torch.float64), 100.0, 0)
        g64 = g.to(torch.float64)
        gn = g64 / torch.sqrt((g64 * g64).sum(dim=1, keepdim=True))
        s = s - st[:, None] * gn
    s1, st1, g1, H1 = smp.project(x0, gt_mode, 100.0, num_steps)
    ok = torch.isfinite(s).all(dim=1)
    assert torch.equal(ok, torch.isfinite(s1).all(dim=1))           # negative values give NaN steps in both (no abs(), like the reference)
    # |g| may round differently by one fp32 ulp, which moves a point by 1e-9 per step and (tensor-core path) can flip fp16
    # roundings downstream: exact for a single step, rounding-level agreement otherwise
    if num_steps == 1:
        assert float((s[ok] - s1[ok]).abs().max()) <= 1e-15
        assert torch.equal(st[ok], st1[ok]) and torch.equal(g[ok], g1[ok]) and torch.equal(H[ok], H1[ok])
    else:
        tol = 1e-5 if precision == "fp32" else 2e-3
        assert float((s[ok] - s1[ok]).abs().max()) <= tol
        assert float((st[ok] - st1[ok]).abs().max()) <= tol
        assert float((g[ok] - g1[ok]).abs().max()) <= tol * float(g[ok].abs().max())
        assert float((H[ok] - H1[ok]).abs().max()) <= tol * float(H[ok].abs().max())
    assert torch.equal(x0, torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, (4097, 3))).cuda())   # input not aliased
