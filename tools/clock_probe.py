"""SM clock and board power while the tensor-core grid query / the training step run back to back for a few seconds
(explains the tensor-pipe ceiling the query kernel sees: the board power-limits long tensor-heavy kernels)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ALPHA, LR, W_S1, ClockSampler, make_batches  # noqa: E402
from diffudf_b200 import SIREN  # noqa: E402
from diffudf_b200.train import FusedTrainer  # noqa: E402

torch.manual_seed(123)
m = SIREN(3, 1, [256] * 8, w0=30).cuda()
eng = m._engine_synced(2)
N = 256
df = torch.empty(N ** 3, device="cuda")
vecs = torch.empty(N ** 3, 3, device="cuda")


def run(name, fn, seconds=3.0):
    fn()
    torch.cuda.synchronize()
    cs = ClockSampler(0)
    cs.start()
    t0, n = time.time(), 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    while time.time() - t0 < seconds:
        fn()
        n += 1
        if n % 8 == 0:
            torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    raw = list(cs.lines)
    info = cs.stop()
    pw = [float(l.split(",")[2]) for l in raw if len(l.split(",")) >= 3 and l.split(",")[2].strip().replace(".", "").isdigit()]
    print(f"{name:28s}: {s.elapsed_time(e) / n:8.3f} ms/iter  clocks {info}  power W median {np.median(pw) if pw else None}", flush=True)


run("grid 256^3 tc16 (f, grad)", lambda: eng.query_grid(N, 0, N ** 3, "tc16", 3, 100.0, want_vecs=True, want_hess=False, out=(df, vecs)))
tr = FusedTrainer(m, precision="tc16")
batches = [tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x[0], n[0], d[0, :, 0])) for x, n, d in make_batches(2, 0)]
state = {"i": 0}


def step():
    state["i"] += 1
    tr.step("s1", *batches[state["i"] % 2], 9990, W_S1, ALPHA, LR)


run("fused train step tc16", step)
