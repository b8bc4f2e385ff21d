"""BASELINE configs[0] (beetle, train_cfg.json, short run) against the run of the unmodified reference
recorded in tests/golden/beetle_traj.npz (tests/golden/make_golden_beetle.py): 4 warm-up steps of loss_s1 at lr 1e-4, 4 at
lr_s1 = 1e-5, 4 steps of loss_s2 with the cosine learning rate, 2 997-row batches drawn by the reference's own sampler from
10 000 surface samples of the normalised mesh."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(weights):
    from diffudf_b200 import SIREN
    m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
    m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(weights["init"]) for k, v in (("weight", W), ("bias", b))})
    return m.cuda()


@pytest.mark.parametrize("precision", ["fp32", "tc16"])
def test_beetle_schedule_matches_reference_run(precision, golden, weights):
    from diffudf_b200.train import FusedTrainer, lr_for_epoch
    T = golden("beetle_traj.npz")
    m = _model(weights)
    tr = FusedTrainer(m, precision=precision)
    n_on = int(T["n_on"])
    for e in range(12):
        lr = lr_for_epoch(e, 12, 8, 4, 1e-4, 1e-5, 1e-7)
        assert abs(lr - float(T["lr"][e])) <= 1e-12
        mode, w = ("s1", [1e4, 1e4, 1e4, 1e3]) if e < 8 else ("s2", [1e5, 1e5])
        x, n, d = (torch.from_numpy(np.ascontiguousarray(T[k][e])).cuda() for k in ("x", "normals", "d"))
        terms = tr.step(mode, x, n, d, n_on, w, 100.0, lr).cpu().numpy()
        ref = T[f"loss{e}"]
        # the first step sees identical weights: parity tolerance of the arithmetic (fp32 1e-4, tensor-core 2e-3 on these
        # 2 997-row means).  Afterwards Adam's sign-like first update (every parameter moves by +-lr, the sign of a near-zero
        # gradient is rounding noise) makes fp32 trajectories from the SIREN init diverge: 0.7 % in the Hessian term at step 1
        # (tests/test_gpu_losses.py::test_fused_trainer_trajectory): sanity band.
        # From step 2 on the run is a different sample of a chaotic trajectory (Adam's first updates move every parameter by +-lr
        # and the sign of a near-zero gradient is rounding noise; the fp32 step also accumulates with float atomics).  Measured
        # spread between OUR OWN fp32 runs on identical inputs: on-surface term of step 4 = 61 / 138 / 167 / 173, Hessian term of
        # steps 4-7 = 1.0x ... 1.95x the reference's, total within 10 %.  So from step 2 on: the total within 35 %, every term
        # within a factor of 3 (+ 5 % of the total for the small ones); steps 0 and 1 are the parity checks.
        if e < 8:
            if e < 2:
                rtol = (1e-4 if precision == "fp32" else 2e-3) if e == 0 else 2e-2
                assert np.allclose(terms[: len(ref)], ref, rtol=rtol, atol=1e-3), (e, terms, ref)
            else:
                tot = float(ref.sum())
                assert abs(float(terms[: len(ref)].sum()) - tot) <= 0.35 * tot, (e, terms, ref)
                assert np.all(terms[: len(ref)] <= 3.0 * ref + 0.05 * tot) and np.all(terms[: len(ref)] >= ref / 3.0 - 0.05 * tot), (e, terms, ref)
        else:
            # loss_s2 after 8 diverged steps: the spread of the on-surface predictions is comparable, their signed mean
            # (|mean| is the first term) is a cancellation of values of either sign and is only bounded by the spread
            assert abs(terms[1] - ref[1]) <= 0.35 * ref[1] and 0.0 <= terms[0] <= 3.0 * ref[1], (e, terms, ref)
    assert all(torch.isfinite(v).all() for v in m.state_dict().values())
    # size of the whole 12-step update, per tensor, against the reference's
    ws, bs = m._weights_biases()
    for i, (W, b) in enumerate(weights["init"]):
        for key, cur, ini in ((f"dnorm_W{i}", ws[i], W), (f"dnorm_b{i}", bs[i], b)):
            got = float(np.linalg.norm(cur.detach().cpu().numpy().astype(np.float64) - ini.astype(np.float64).reshape(tuple(cur.shape))))
            want = float(T[key][0])
            if cur.numel() >= 256:            # a scalar's 12-step walk of +-lr steps is not a statistic
                assert abs(got - want) <= 0.15 * want, (key, got, want)


def test_device_sampler_reproduces_the_reference_distances_on_the_beetle_cloud(golden):
    """The far rows of the recorded batches carry nearest-cloud-point distances computed by the reference in fp64."""
    from diffudf_b200.dataset import shortestDistance
    T = golden("beetle_traj.npz")
    X = torch.from_numpy(T["cloud_pts"]).cuda()
    n_on, n_far = int(T["n_on"]), int(T["n_off"]) // 2
    for e in (0, 5, 11):
        far = torch.from_numpy(np.ascontiguousarray(T["x"][e][n_on:n_on + n_far])).cuda()
        want = T["d"][e][n_on:n_on + n_far].astype(np.float64)
        got = shortestDistance(far, X).cpu().numpy().astype(np.float64)
        assert np.abs(got ** 2 - want ** 2).max() < 1e-6
        on = torch.from_numpy(np.ascontiguousarray(T["x"][e][:n_on])).cuda()
        assert float(shortestDistance(on, X).max()) == 0.0


@pytest.mark.parametrize("precision", ["fp32", "tcx3"])
def test_beetle_full_size_schedule_teacher_forced(precision, golden, weights):
    """BASELINE configs[0] AT FULL SIZE: the normalised beetle mesh (tests/golden/beetle_mesh.npz), 100 000
    area-weighted surface samples, 29 970-row batches [9 990 on | 9 990 far | 9 990 near] drawn by the device sampler in its mesh
    mode, and train_cfg.json's schedule compressed to 20 + 20 + 20 steps (warm-up lr 1e-4, lr_s1 1e-5, loss_s2 with the cosine
    lr; train.py:167-191).  TEACHER = the reference's own operation sequence (oracle/autograd_port.py: autograd double-backward,
    torch.linalg.eigh, backward, torch.optim.Adam; pinned to the unmodified reference on the CPU by tests/test_oracle_golden.py) run
    in float64 on the same GPU.  Before EVERY step the trainer under test is reset to the teacher's weights and Adam moments, so
    each of the 60 steps is a parity check of loss terms, parameter gradient and parameter update — not a sample of a chaotic
    trajectory."""
    from diffudf_b200.dataset import PointCloud
    from diffudf_b200.preprocess_mesh import sample_points_uniformly
    from diffudf_b200.train import FusedTrainer, lr_for_epoch
    from oracle import autograd_port as AP
    dev = torch.device("cuda:0")
    tri = golden("beetle_mesh.npz")["tri"].astype(np.float32)
    V = tri.reshape(-1, 3)
    F = np.arange(V.shape[0]).reshape(-1, 3)
    pts, nrm = sample_points_uniformly(V, F, 100000, dev, seed=7)
    ds = PointCloud(pts, nrm, 30000, [0.333, 0.666], 60, dev, seed=11, triangles=tri)
    n_on = ds.samplesOnSurface
    m = _model(weights)
    tr = FusedTrainer(m, precision=precision)
    teacher = AP.make_params(weights["init"], dtype=torch.float64, device="cuda:0")
    opt = AP.make_optimizer(teacher, 1e-4)
    flat_of = lambda get: torch.cat([get(t).reshape(-1) for pair in teacher for t in pair])
    # loss terms / parameter gradient (rel-L2 of the flat gradient) / parameter update (rel-L2).  fp32: the gradient is accumulated with
    # float atomics over ~700 partial sums per element (measured 2e-5 ... 1.3e-4); tcx3: single-pass reverse sweep (measured 2e-4 ... 2e-3)
    tol_t, tol_g, tol_u = {"fp32": (1e-4, 3e-4, 2e-3), "tcx3": (5e-4, 2e-3, 3e-2)}[precision]
    worst = {"terms": 0.0, "grad": 0.0, "update": 0.0}
    for e, (x, n, d) in enumerate(ds):
        lr = lr_for_epoch(e, 60, 40, 20, 1e-4, 1e-5, 1e-7)
        mode, w = ("s1", [1e4, 1e4, 1e4, 1e3]) if e < 40 else ("s2", [1e5, 1e5])
        assert x.shape == (1, 29970, 3) and int((d == 0).sum()) == n_on == 9990
        # ---- teacher forcing: weights and Adam state of the teacher into the trainer under test
        w_t = flat_of(lambda t: t.detach())
        tr.flat.copy_(w_t.float())
        if e == 0:
            tr.m.zero_(); tr.v.zero_()
        else:
            tr.m.copy_(flat_of(lambda t: opt.state[t]["exp_avg"]).float())
            tr.v.copy_(flat_of(lambda t: opt.state[t]["exp_avg_sq"]).float())
        tr.t = e
        terms = tr.step(mode, x[0], n[0], d[0, :, 0], n_on, w, 100.0, lr).cpu().numpy()
        for g in opt.param_groups:
            g["lr"] = lr
        # yardstick: the reference's OWN arithmetic (fp32) at the same weights against the float64 teacher.  After a few steps some
        # on-surface rows have nearly degenerate Hessian eigenvalues (the alignment term's gradient carries 1 / (lambda_2 - lambda_j))
        # and residuals that sit on the kink of |.|: there the gradient is ill-conditioned and ANY fp32 evaluation — the reference's
        # included — is 1e-3-class off the float64 one (measured 3e-3 at step 2).  The bar for the path under test is therefore
        # max(arithmetic tolerance, 3 x the reference-fp32 error of that step).
        p32 = [(W.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)) for W, b in teacher]
        l32 = (AP.loss_s1 if mode == "s1" else AP.loss_s2)(p32, x, n, d, w, 100.0)
        sum(l32.values()).backward()
        g32 = torch.cat([t.grad.reshape(-1) for pair in p32 for t in pair]).double()
        ref = AP.train_step(teacher, opt, x.double(), n.double(), d.double(), mode, w, 100.0)
        ref_t = np.array(list(ref.values()))
        # loss_s2's first term is |mean f| over the on-surface rows — a cancellation of values of either sign whose natural scale is
        # their spread (the second term, w std f): both terms are measured against the larger of the two
        scale = np.maximum(np.abs(ref_t), 1e-2) if mode == "s1" else np.full(len(ref_t), max(float(np.abs(ref_t).max()), 1e-2))
        e_t = float(np.max(np.abs(terms[: len(ref_t)] - ref_t) / scale))
        g_ref = flat_of(lambda t: t.grad)
        e_ref = float(torch.linalg.norm(g32 - g_ref) / torch.linalg.norm(g_ref))
        e_g = float(torch.linalg.norm(tr.grad.double() - g_ref) / torch.linalg.norm(g_ref))
        du_ref = flat_of(lambda t: t.detach()) - w_t
        du = tr.flat.double() - w_t.float().double()
        e_u = float(torch.linalg.norm(du - du_ref) / torch.linalg.norm(du_ref))
        # fp32 parameters cannot hold an update finer than their spacing: in the loss_s2 stage (lr <= 1e-7) an update is a few ulps of
        # the weight, and that quantisation — the reference's fp32 parameters have it too — is the floor of this comparison
        w32 = w_t.float()
        ulp = (torch.nextafter(w32.abs(), torch.full_like(w32, float("inf"))) - w32.abs()).double()
        e_u_floor = float(torch.linalg.norm(ulp) / torch.linalg.norm(du_ref))
        worst = {"terms": max(worst["terms"], e_t), "grad": max(worst["grad"], e_g), "update": max(worst["update"], e_u if e > 0 else 0.0),
                 "grad_reference_fp32": max(worst.get("grad_reference_fp32", 0.0), e_ref),
                 "grad_over_reference_fp32": max(worst.get("grad_over_reference_fp32", 0.0), e_g / max(e_ref, 1e-12))}
        # the very first Adam update is lr * sign(gradient): where a gradient is rounding noise its sign is too, so step 0's
        # update is held to a looser bar than the steps with teacher-forced moments
        assert e_t < tol_t and e_g < max(tol_g, 3 * e_ref) and e_u < (0.3 if e == 0 else max(tol_u, 10 * e_ref)) + e_u_floor, \
            (precision, e, mode, e_t, e_g, e_ref, e_u, terms, ref_t)
    print(f"beetle full size, {precision}: worst over 60 teacher-forced steps: {worst}")
