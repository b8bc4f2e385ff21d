"""Errors and rates of the three arithmetics (fp32 CUDA cores, tc16 single-pass, tcx3 split) side by side: queries against the
fp64 oracle, grid / point throughput, the training step per route.  Diagnostics (imports oracle/ as the checker)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffudf_b200 import SIREN, synthetic  # noqa: E402
from diffudf_b200.train import FusedTrainer  # noqa: E402
from oracle import dudf_oracle as O  # noqa: E402

F = {0: 919552, 1: 3673600, 2: 9181696}
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rel_max(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / np.max(np.abs(b)))


def load(tag):
    params = O.load_params(os.path.join(G, f"weights_{tag}.npz"))
    m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
    m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(params) for k, v in (("weight", W), ("bias", b))})
    return params, m.cuda()


def timed(fn, reps):
    for _ in range(3):
        fn()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        fn()
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / reps


what = sys.argv[1:] or ["err", "rates", "train"]
if "err" in what:
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (20000, 3)).astype(np.float32)
    for tag in ("init", "trained"):
        params, m = load(tag)
        ref = O.siren_jet(params, x, 2)
        eng = m._engine_synced()
        xt = torch.from_numpy(x).cuda()
        for prec in ("fp32", "tc16", "tcx3"):
            f, g, H, _ = eng.query(xt, 2, prec)
            print(f"err {tag:8s} {prec}: f {rel_max(f.cpu().numpy(), ref['f']):.2e}  g {rel_max(g.cpu().numpy(), ref['g']):.2e}  "
                  f"H {rel_max(H.cpu().numpy(), ref['H']):.2e}", flush=True)
if "rates" in what:
    params, m = load("trained")
    eng = m._engine_synced()
    for prec in ("tc16", "tcx3"):
        for order in (0, 1, 2):
            for P in (1 << 20, 1 << 14):
                xq = torch.rand(P, 3, device="cuda") * 2 - 1
                ms = timed(lambda: eng.query(xq, order, prec), 10)
                print(f"rate {prec} order {order} P {P:8d}: {ms:8.4f} ms {P / ms / 1e3:8.1f} M q/s {P * F[order] / ms / 1e9:7.1f} TFLOP/s", flush=True)
        N = 256
        df = torch.empty(N ** 3, device="cuda")
        vecs = torch.empty(N ** 3, 3, device="cuda")
        ms = timed(lambda: eng.query_grid(N, 0, N ** 3, prec, flags=3, alpha=100.0, out=(df, vecs)), 3)
        print(f"rate {prec} grid {N}^3 (f, grad): {ms:8.3f} ms {N ** 3 / ms / 1e3:8.1f} M q/s {N ** 3 * F[1] / ms / 1e9:7.1f} TFLOP/s", flush=True)
if "grid" in what:          # one line per process: used by the probe loop over DUDF_TCX_* environment settings
    params, m = load("trained")
    eng = m._engine_synced()
    N = 256
    df = torch.empty(N ** 3, device="cuda")
    vecs = torch.empty(N ** 3, 3, device="cuda")
    ms = timed(lambda: eng.query_grid(N, 0, N ** 3, "tcx3", flags=3, alpha=100.0, out=(df, vecs)), 3)
    xq = torch.rand(1 << 20, 3, device="cuda") * 2 - 1
    ms0 = timed(lambda: eng.query(xq, 0, "tcx3"), 10)
    ms2 = timed(lambda: eng.query(xq, 2, "tcx3"), 5)
    env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("DUDF_TCX"))
    print(f"probe [{env}] grid256 (f,grad) {ms:8.2f} ms {N ** 3 / ms / 1e3:7.1f} M q/s | value 1M {ms0:6.3f} ms {(1 << 20) / ms0 / 1e3:7.1f} M q/s | "
          f"hess 1M {ms2:6.2f} ms {(1 << 20) / ms2 / 1e3:6.1f} M q/s", flush=True)
if "train" in what:
    shape = synthetic.make_shape(0)
    sp, sn = shape.sample_surface(200000, np.random.default_rng(0))
    x, n, d = synthetic.make_batch(shape, sp, sn, 30000, (0.333, 0.666), np.random.default_rng(5))
    xd, nd, dd = torch.from_numpy(x[0]).cuda(), torch.from_numpy(n[0]).cuda(), torch.from_numpy(d[0, :, 0]).cuda()
    n_on = int((d == 0).sum())
    w = [1e4, 1e4, 1e4, 1e3]
    for prec, fused in (("tc16", True), ("tc16", False), ("tcx3", False)):
        torch.manual_seed(123)
        m = SIREN(3, 1, [256] * 8, w0=30).cuda()
        tr = FusedTrainer(m, precision=prec, fused=fused)
        ms = timed(lambda: tr.step("s1", xd, nd, dd, n_on, w, 100.0, 1e-5), 20)
        print(f"train {prec} fused={fused}: {ms:.4f} ms/step {xd.shape[0] / ms / 1e3:.2f} M pts/s", flush=True)
if "ref" in what:
    from oracle import autograd_port as AP
    shape = synthetic.make_shape(0)
    sp, sn = shape.sample_surface(200000, np.random.default_rng(0))
    x, n, d = synthetic.make_batch(shape, sp, sn, 30000, (0.333, 0.666), np.random.default_rng(5))
    params_np = O.init_params()
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        params = AP.make_params(params_np, device="cuda:0")
        opt = AP.make_optimizer(params, 1e-5)
        xt, nt, dt = (torch.from_numpy(a).cuda() for a in (x, n, d))
        for mode, w in (("s1", [1e4, 1e4, 1e4, 1e3]), ("s1", [1e4, 1e4, 0, 1e3]), ("s2", [1e5, 1e5])):
            for _ in range(2):
                AP.train_step(params, opt, xt, nt, dt, mode, w, 100.0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                AP.train_step(params, opt, xt, nt, dt, mode, w, 100.0)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / 5 * 1e3
            print(f"reference eager B200 tf32={tf32} {mode} w2={w[2] if len(w) > 2 else '-'}: {ms:.2f} ms/step {xt.shape[1] / ms / 1e3:.3f} M pts/s", flush=True)
        xq = np.random.default_rng(0).uniform(-1, 1, (1 << 18, 3)).astype(np.float32)
        for wg, wh in ((True, False), (True, True)):
            AP.evaluate(params, xq[:8192], wg, wh)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            AP.evaluate(params, xq, wg, wh)
            torch.cuda.synchronize()
            s = time.perf_counter() - t0
            print(f"reference eager B200 tf32={tf32} evaluate grad={wg} hess={wh}: {xq.shape[0] / s / 1e6:.3f} M q/s", flush=True)
