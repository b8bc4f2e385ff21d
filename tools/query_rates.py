"""Throughput of dudf_query_points per jet order and batch size (tensor-core path): where the sphere-tracing time goes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffudf_b200 import SIREN  # noqa: E402

torch.manual_seed(123)
m = SIREN(3, 1, [256] * 8, w0=30).cuda()
eng = m._engine_synced()
F = {0: 919552, 1: 3673600, 2: 9181696}
for order in (0, 1, 2):
    for P in (1 << 20, 1 << 17, 1 << 14, 1 << 11):
        x = torch.rand(P, 3, device="cuda") * 2 - 1
        for _ in range(3):
            eng.query(x, order, "tc16")
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        torch.cuda.synchronize()
        s.record()
        for _ in range(reps):
            eng.query(x, order, "tc16")
        t.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(t) / reps
        print(f"order {order}  P {P:8d}: {ms:8.4f} ms  {P / ms / 1e3:8.1f} M queries/s  {P * F[order] / ms / 1e9:7.1f} TFLOP/s", flush=True)
