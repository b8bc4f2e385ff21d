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
