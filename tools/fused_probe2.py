import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from bench import ALPHA, LR, W_S1, make_batches
from diffudf_b200 import SIREN
from diffudf_b200.train import FusedTrainer
batches = [tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x[0], n[0], d[0, :, 0])) for x, n, d in make_batches(4, 0)]
for base in (3, 11, 27):
    for name, dbg in (("full", 0), ("no weights", 4), ("no images", 16), ("no scratch", 32), ("no weights+images", 20), ("no weights+images+scratch", 52), ("no MMA", 2), ("nothing (epilogue only)", 54)):
        torch.manual_seed(123)
        model = SIREN(3, 1, [256] * 8, w0=30).cuda()
        tr = FusedTrainer(model, precision="tc16", fused=True)
        tr.core.fused_flags = base | 4 | (dbg << 8)      # 4: no wgrad inside the step -> times the fused launch (+ adam etc.)
        try:
            for i in range(4):
                tr.step("s1", *batches[i % 4], 9990, W_S1, ALPHA, LR)
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); s.record()
            for i in range(20):
                tr.step("s1", *batches[i % 4], 9990, W_S1, ALPHA, LR)
            t.record(); torch.cuda.synchronize()
            print(f"flags {base:2d} {name:28s}: {s.elapsed_time(t) / 20:7.4f} ms/step", flush=True)
        except Exception as e:
            print(base, name, "failed", e)
