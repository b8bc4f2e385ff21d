"""The public training loops (train_model_tanh / train_model_siren, BatchFeeder, FusedTrainer) against the drop-in
route the reference's own loop takes (loss_fn -> backward -> torch.optim.Adam.step, train.py:195-222)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class TinyDataset:
    """Same iteration protocol as the reference's PointCloud dataset (src/dataset.py:134-185)."""

    def __init__(self, batches, n_on):
        self.batches = batches
        self.batchesPerEpoch = len(batches)
        self.samplesOnSurface = n_on

    def __iter__(self):
        for x, n, d in self.batches:
            yield torch.from_numpy(x), torch.from_numpy(n), torch.from_numpy(d)


def _make(weights, n_batches=2, rows=1500):
    from diffudf_b200 import SIREN, synthetic
    shape = synthetic.make_shape(0)
    sp, sn = shape.sample_surface(20000, np.random.default_rng(0))
    batches = [synthetic.make_batch(shape, sp, sn, rows, (0.333, 0.666), np.random.default_rng(40 + i)) for i in range(n_batches)]
    m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
    m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(weights["trained"])
                       for k, v in (("weight", W), ("bias", b))})
    return m.cuda(), batches, int(rows * 0.333)


def test_fused_loop_matches_dropin_adam_loop(weights):
    import diffudf_b200 as D
    from diffudf_b200.train import train_model_tanh
    cfg = dict(epochs=3, s1_epochs=2, warmup_epochs=1, warmup_lr=1e-5, lr_s1=1e-6, lr_s2=1e-7, loss_s1_weights=[1e4, 1e4, 1e4, 1e3],
               loss_s2_weights=[1e5, 1e5], alpha=100.0)
    m1, batches, n_on = _make(weights)
    losses, best, seconds = train_model_tanh(TinyDataset(batches, n_on), m1, torch.device("cuda:0"), cfg)
    assert set(losses) == {"sdf_on_surf", "sdf_off_surf", "hessian_constraint", "grad_constraint", "std_on_surf"}
    assert all(len(v) == 3 for v in losses.values()) and best is not None and seconds > 0
    # the reference-style loop on the drop-in surface
    m2, _, _ = _make(weights)
    opt = torch.optim.Adam(m2.parameters(), lr=1e-5)
    ref = {k: [0.0] * 3 for k in losses}
    for epoch in range(3):
        lr = 1e-5 if epoch < 1 else (1e-6 if epoch < 2 else 0.5 * (np.cos(epoch / 1 * np.pi) + 1) * 1e-7)
        for g in opt.param_groups:
            g["lr"] = lr
        for x, n, d in batches:
            opt.zero_grad()
            gt = {"normals": torch.from_numpy(n).cuda(), "sdf": torch.from_numpy(d).cuda()}
            loss = (D.loss_s1(m2, torch.from_numpy(x).cuda(), gt, cfg["loss_s1_weights"], 100.0) if epoch < 2
                    else D.loss_s2(m2, torch.from_numpy(x).cuda(), gt, cfg["loss_s2_weights"], 100.0))
            tot = 0
            for k, v in loss.items():
                tot = tot + v
                ref[k][epoch] += float(v.detach())
            tot.backward()
            opt.step()
    for k in losses:
        assert np.allclose(losses[k], ref[k], rtol=2e-3, atol=1e-3), (k, losses[k], ref[k])
    for p1, p2 in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(p1, p2, rtol=0, atol=3e-5)          # 5 Adam steps of <= 1e-5 each


def test_siren_loop_and_tc16_precision_run(weights):
    from diffudf_b200.train import train_model_siren
    m, batches, n_on = _make(weights, n_batches=1)
    cfg = dict(epochs=2, loss_weights=[3e3, 1e2, 1e2, 5e1], lr=1e-5, precision="tc16")
    losses, best, _ = train_model_siren(TinyDataset(batches, n_on), m, torch.device("cuda:0"), cfg)
    assert set(losses) == {"sdf_on_surf", "sdf_off_surf", "normal_constraint", "grad_constraint"}
    assert all(np.isfinite(v).all() for v in losses.values())
    assert all(torch.isfinite(p).all() for p in m.parameters())


def test_batch_feeder_preserves_order_and_content():
    from diffudf_b200.train import BatchFeeder
    feeder = BatchFeeder(torch.device("cuda:0"))
    batches = [(torch.full((7, 3), float(i)).pin_memory(), torch.full((7, 3), float(-i)).pin_memory(), torch.full((7,), float(10 * i)).pin_memory())
               for i in range(9)]
    seen = []
    for x, n, d in feeder.feed(batches):
        seen.append((float(x[3, 1]), float(n[0, 2]), float(d[6])))
    assert seen == [(float(i), float(-i), float(10 * i)) for i in range(9)]
    assert list(feeder.feed([])) == []


def test_fused_step_skips_the_update_when_the_loss_scale_is_outgrown(weights):
    """the fused tc16 step scales its fp16 adjoints with the PREVIOUS step's seed magnitude.  When the seeds
    outgrow that scale (here: the recorded magnitude is shrunk 100x, i.e. the scale is 64-128x too large and adjoints would
    saturate), the guard flag trips, Adam leaves parameters and moments untouched, the skip is counted, and the next step —
    whose scale comes from the skipped step's own magnitude — updates normally again."""
    from diffudf_b200.train import FusedTrainer
    m, batches, n_on = _make(weights, n_batches=1, rows=3000)
    x, n, d = (torch.from_numpy(a).cuda() for a in (batches[0][0][0], batches[0][1][0], batches[0][2][0, :, 0]))
    w = [1e4, 1e4, 1e4, 1e3]
    tr = FusedTrainer(m, precision="tc16")
    tr.step("s1", x, n, d, n_on, w, 100.0, 1e-6)          # three-kernel route: measures the scale
    tr.step("s1", x, n, d, n_on, w, 100.0, 1e-6)          # fused launch, scale one step stale but adequate
    assert tr.core.last_fused is not None and tr.skipped_steps() == 0
    tr.core.amax[tr.core.amax_slot] *= 1e-2               # pretend the previous step's seeds were 100x smaller
    before, m_before = tr.flat.clone(), tr.m.clone()
    tr.step("s1", x, n, d, n_on, w, 100.0, 1e-6)
    assert tr.skipped_steps() == 1
    assert torch.equal(tr.flat, before) and torch.equal(tr.m, m_before)
    tr.step("s1", x, n, d, n_on, w, 100.0, 1e-6)
    assert tr.skipped_steps() == 1 and not torch.equal(tr.flat, before)
    assert bool(torch.isfinite(tr.flat).all())


def test_adam_step_peers_with_one_rank_equals_adam_step():
    """dudf_adam_step_peers (gradient sum over peer-mapped buffers fused into Adam) with a single 'peer' — this process's own buffer —
    is dudf_adam_step bit for bit, including the tail that is not a multiple of 4 and the guard flag behind the gradient."""
    import ctypes
    from diffudf_b200.engine import adam_step, adam_step_peers
    torch.manual_seed(0)
    n = 4099
    g_all = torch.randn(n + 1, device="cuda")
    g_all[-1] = 0.0
    p1, m1, v1 = torch.randn(n, device="cuda"), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    p2, m2, v2 = p1.clone(), m1.clone(), v1.clone()
    ptrs = (ctypes.c_void_p * 1)(g_all.data_ptr())
    gsum = torch.zeros(n, device="cuda")
    skipped = torch.zeros(1, device="cuda", dtype=torch.int64)
    for t in (1, 2, 3):
        adam_step(p1, g_all[:n], m1, v1, 1e-3, t)
        adam_step_peers(p2, ptrs, 1, m2, v2, 1e-3, t, guarded=True, skipped=skipped, g_sum_out=gsum)
    assert torch.equal(p1, p2) and torch.equal(m1, m2) and torch.equal(v1, v2) and torch.equal(gsum, g_all[:n]) and int(skipped) == 0
    # two "ranks" that are the same buffer: the sum is 2 g, in rank order
    ptrs2 = (ctypes.c_void_p * 2)(g_all.data_ptr(), g_all.data_ptr())
    adam_step_peers(p2, ptrs2, 2, m2, v2, 1e-3, 4, g_sum_out=gsum)
    assert torch.equal(gsum, g_all[:n] + g_all[:n])
    # a raised guard flag skips the update on "every rank" and is counted
    g_all[-1] = 1.0
    before = p2.clone()
    adam_step_peers(p2, ptrs, 1, m2, v2, 1e-3, 5, guarded=True, skipped=skipped)
    assert torch.equal(p2, before) and int(skipped) == 1


@pytest.mark.parametrize("prec", ["fp32", "tcx3"])
def test_graph_replayed_steps_equal_eager_steps(prec, golden, weights):
    """FusedTrainer(graph=True): the step captured as a CUDA graph (learning rate and Adam's step count on the device,
    dudf_adam_step_dev) against the eager launches — same loss terms, same parameters after s1 and s2 steps with a changing
    learning rate and an eager step in between (the device count is re-synchronised)."""
    from diffudf_b200 import SIREN
    from diffudf_b200.train import FusedTrainer
    T = golden("trajectory_init.npz")
    runs = []
    for graph in (False, True):
        m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
        m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(weights["init"])
                           for k, v in (("weight", W), ("bias", b))})
        tr = FusedTrainer(m.cuda(), precision=prec, graph=graph)
        out = []
        # s1 and s2 alternate at the start: each signature is captured AFTER a step of the other shape evicted its cached workspaces
        sched = [("s1", 1e-5), ("s2", 1e-7), ("s1", 1e-5), ("s1", 2e-6), ("s2", 1e-7), ("s2", 1e-7), ("s2", 5e-8), ("s1", 2e-6), ("s2", 5e-8)]
        for step, (mode, lr) in enumerate(sched):
            b = step % 7
            x = torch.from_numpy(T["x"][b][0]).cuda()
            n = torch.from_numpy(T["normals"][b][0]).cuda()
            d = torch.from_numpy(T["d"][b][0, :, 0]).cuda()
            n_on = int((T["d"][b][0, :, 0] == 0).sum())
            w = [1e4, 1e4, 1e4, 1e3] if mode == "s1" else [1e5, 1e5]
            out.append(tr.step(mode, x, n, d, n_on, w, 100.0, lr).cpu().numpy())
        runs.append((out, tr.flat.clone(), tr.t))
        if graph:
            assert sum(isinstance(g, dict) for g in tr._graphs.values()) == 2           # s1 and s2 were captured
            assert int(tr._adam_state[2:4].view(torch.int64).item()) == tr.t == len(sched)
    (ta, pa, _), (tb, pb, _) = runs
    for a, b in zip(ta, tb):
        assert np.allclose(a, b, rtol=2e-4, atol=1e-7), (a, b)
    # Same arithmetic (bias corrections in double on the device instead of on the host).  Gradients are accumulated with atomics, so
    # two runs differ in the last bits, and Adam's first updates are sign-like (lr * g / |g|): an element whose gradient is ~0 may
    # move by 2 lr in either run.  A wrong step count or learning rate would shift EVERY element by O(lr) = 1e-5.
    diff = (pa - pb).abs()
    assert float(diff.max()) < 2.5e-5 and float(diff.mean()) < 2e-7, (float(diff.max()), float(diff.mean()))
