"""Parity of the fused training step: loss terms, parameter gradients, Adam — against the fp64 oracle and the
reference's own numbers (golden fixtures).  fp32 path: terms 2e-5 relative; gradients 1e-3 in the max measure
per tensor (fp32 accumulation over hundreds of rows and 1/(eigen-gap) amplification in the Hessian term)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_max

pytestmark = pytest.mark.gpu

MODES = [("s1", [1e4, 1e4, 1e4, 1e3]), ("s1_nohess", [1e4, 1e4, 0, 1e3]), ("s2", [1e5, 1e5]), ("siren", [3e3, 1e2, 1e2, 5e1])]


def _loss(model, mode, x, n, d, w, alpha=100.0):
    import diffudf_b200 as D
    gt = {"normals": torch.from_numpy(n).cuda(), "sdf": torch.from_numpy(d).cuda()}
    xi = torch.from_numpy(x).cuda()
    if mode.startswith("s1"):
        return D.loss_s1(model, xi, gt, w, alpha)
    if mode == "s2":
        return D.loss_s2(model, xi, gt, w, alpha)
    return D.loss_siren(model, xi, gt, w)


@pytest.mark.parametrize("tag", ["init", "trained"])
@pytest.mark.parametrize("mode,w", MODES)
def test_loss_terms_and_gradients(tag, mode, w, golden, oracle, weights, cuda_models):
    Ld = golden(f"losses_{tag}.npz")
    m = cuda_models[tag]
    for p in m.parameters():
        p.requires_grad_(True)
        p.grad = None
    loss = _loss(m, mode, Ld["x"], Ld["normals"], Ld["d"], w)
    terms_ref, grads_ref = oracle.train_grads(weights[tag], Ld["x"], Ld["normals"], Ld["d"], mode.split("_")[0], w, 100.0)
    assert list(loss.keys()) == list(terms_ref.keys())
    for k, v in loss.items():
        ref = float(terms_ref[k])
        got = float(v)
        assert abs(got - ref) <= 5e-5 * max(abs(ref), 1e-2), (k, got, ref)
        ref32 = float(Ld[f"{mode}_32_{k}"][0])
        assert abs(got - ref32) <= 2e-3 * max(abs(ref32), 1e-2), (k, got, ref32)       # the reference's own fp32 run
    total = 0
    for v in loss.values():
        total = total + v
    total.backward()
    worst = {}
    for i in range(len(grads_ref)):
        gW = m.net[i][0].weight.grad.cpu().numpy()
        gb = m.net[i][0].bias.grad.cpu().numpy()
        worst[f"W{i}"] = rel_max(gW, grads_ref[i][0].reshape(gW.shape))
        worst[f"b{i}"] = rel_max(gb, grads_ref[i][1].reshape(gb.shape))
        assert rel_max(gW.reshape(-1)[::37], Ld[f"{mode}_gWsub{i}"]) < 2e-3        # reference autograd (fp64 run)
    print(f"{tag} {mode}: worst grad errors {max(worst.values()):.2e}")
    assert max(worst.values()) < 1e-3, worst


def test_loss_general_row_order(golden, oracle, weights, cuda_models):
    """On-surface rows need not be a prefix: shuffled rows give the same terms (reference sums are order-free)."""
    Ld = golden("losses_trained.npz")
    perm = np.random.default_rng(0).permutation(Ld["x"].shape[1])
    x, n, d = Ld["x"][:, perm], Ld["normals"][:, perm], Ld["d"][:, perm]
    w = [1e4, 1e4, 1e4, 1e3]
    a = _loss(cuda_models["trained"], "s1", Ld["x"], Ld["normals"], Ld["d"], w)
    b = _loss(cuda_models["trained"], "s1", x, n, d, w)
    for k in a:
        assert abs(float(a[k]) - float(b[k])) <= 1e-5 * max(abs(float(a[k])), 1e-2)


def test_autograd_upstream_weights(golden, oracle, weights, cuda_models):
    """d(sum_k c_k term_k)/d(params) for arbitrary c: the backward honours the upstream gradients."""
    Ld = golden("losses_init.npz")
    m = cuda_models["init"]
    for p in m.parameters():
        p.requires_grad_(True)
        p.grad = None
    c = [0.5, 2.0, 0.0, 3.0]
    loss = _loss(m, "s1", Ld["x"], Ld["normals"], Ld["d"], [1e4, 1e4, 1e4, 1e3])
    tot = sum(ci * v for ci, v in zip(c, loss.values()))
    tot.backward()
    _, gref = oracle.train_grads(weights["init"], Ld["x"], Ld["normals"], Ld["d"], "s1", [1e4, 1e4, 1e4, 1e3], 100.0, upstream=c)
    for i in (0, 3, 8):
        gW = m.net[i][0].weight.grad.cpu().numpy()
        assert rel_max(gW, gref[i][0].reshape(gW.shape)) < 1e-3


def test_dropin_operator_training_path(golden, oracle, weights, cuda_models):
    """The un-fused drop-in route: model() -> gradient() -> torch ops -> backward() reaches the same parameter grads."""
    from diffudf_b200 import gradient
    Ld = golden("losses_init.npz")
    m = cuda_models["init"]
    for p in m.parameters():
        p.requires_grad_(True)
        p.grad = None
    x = torch.from_numpy(Ld["x"]).cuda()
    out = m(x)
    g = gradient(out["model_out"], out["model_in"])
    val = ((g.norm(dim=-1) - 1.0) ** 2).mean() + out["model_out"].abs().mean()
    val.backward()
    P = Ld["x"].shape[1]
    params = weights["init"]
    jet = oracle.siren_jet(params, Ld["x"], 1, keep=True)
    gn = np.linalg.norm(jet["g"], axis=-1)
    gbar = (2 * (gn - 1) / P)[:, None] * jet["g"] / gn[:, None]
    fbar = np.sign(jet["f"]) / P
    gref = oracle.reverse_sweep(params, Ld["x"].reshape(-1, 3).astype(np.float64), jet, fbar, gbar, None)
    for i in (0, 4, 8):
        gW = m.net[i][0].weight.grad.cpu().numpy()
        assert rel_max(gW, gref[i][0].reshape(gW.shape)) < 1e-3


def test_fused_trainer_trajectory(golden, oracle, weights):
    """7 optimisation steps (4 x s1 @1e-4, 3 x s2 @1e-7) through FusedTrainer vs the reference's recorded losses and
    vs the oracle's Adam on the first step (weights after one step are a sign-like update: compare loosely)."""
    from diffudf_b200 import SIREN
    from diffudf_b200.train import FusedTrainer
    T = golden("trajectory_init.npz")
    m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
    m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(weights["init"]) for k, v in (("weight", W), ("bias", b))})
    m = m.cuda()
    tr = FusedTrainer(m)
    for step in range(7):
        mode, w, lr = ("s1", [1e4, 1e4, 1e4, 1e3], 1e-4) if step < 4 else ("s2", [1e5, 1e5], 1e-7)
        x = torch.from_numpy(T["x"][step][0]).cuda()
        n = torch.from_numpy(T["normals"][step][0]).cuda()
        d = torch.from_numpy(T["d"][step][0, :, 0]).cuda()
        n_on = int((T["d"][step][0, :, 0] == 0).sum())
        terms = tr.step(mode, x, n, d, n_on, w, 100.0, lr).cpu().numpy()
        ref = T[f"loss{step}"]
        # fp32 Adam trajectories from the SIREN init are chaotic (test_oracle_golden.py shows the fp64 oracle drifting
        # 5 % from the reference's fp32 run by step 4): tight on the first two steps, sanity band afterwards.
        rtol = 1e-4 if step == 0 else (3e-3 if step == 1 else 0.35)
        assert np.allclose(terms[: len(ref)], ref, rtol=rtol, atol=1e-3), (step, terms, ref)
    sd = m.state_dict()
    assert all(torch.isfinite(v).all() for v in sd.values())


@pytest.mark.parametrize("prec", ["fp32", "tcx3"])
def test_loss_s2_on_surface_prefix_equals_all_rows(prec, golden, weights):
    """loss_s2 reads f only where d == 0 (src/loss_functions.py:106-121): with gt['n_on'] the network is evaluated on the
    leading on-surface rows only — same terms and the same parameter gradient as evaluating every row."""
    import diffudf_b200 as D
    from diffudf_b200 import SIREN
    Ld = golden("losses_trained.npz")
    x, d = Ld["x"].reshape(-1, 3), Ld["d"].reshape(-1)
    order = np.argsort(d != 0, kind="stable")                             # [on | off] layout
    x, d = np.ascontiguousarray(x[order]), np.ascontiguousarray(d[order])
    n_on = int((d == 0).sum())
    assert 0 < n_on < len(d)
    out = []
    for extra in ({}, {"n_on": n_on}):
        m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
        m.train_precision = prec
        m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(weights["trained"])
                           for k, v in (("weight", W), ("bias", b))})
        m = m.cuda()
        gt = {"sdf": torch.from_numpy(d).cuda(), **extra}
        loss = D.loss_s2(m, torch.from_numpy(x).cuda(), gt, [1e5, 1e5], 100.0)
        sum(loss.values()).backward()
        out.append(({k: float(v) for k, v in loss.items()}, [p.grad.clone() for p in m.parameters()]))
    (ta, ga), (tb, gb) = out
    for k in ta:
        assert abs(ta[k] - tb[k]) <= 1e-6 * max(abs(ta[k]), 1e-3), (k, ta[k], tb[k])
    for a, b in zip(ga, gb):
        assert float((a - b).abs().max()) <= 2e-6 * float(a.abs().max()) + 1e-12


def test_adam_kernel_matches_torch():
    from diffudf_b200.engine import adam_step
    torch.manual_seed(0)
    p = torch.randn(100003, device="cuda")
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for t in range(1, 6):
        g = torch.randn_like(p) * (10.0 ** (t - 3))
        ref.grad = g.clone()
        opt.step()
        adam_step(p, g, m, v, 1e-3, t)
    assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-7)
