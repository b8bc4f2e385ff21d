import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from bench import ALPHA, LR, W_S1, make_batches
from diffudf_b200 import SIREN
from diffudf_b200.train import FusedTrainer
batches = [tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x[0], n[0], d[0, :, 0])) for x, n, d in make_batches(4, 0)]
for rep in range(3):
    for flags in ([int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]):
        torch.manual_seed(123)
        tr = FusedTrainer(SIREN(3, 1, [256] * 8, w0=30).cuda(), precision="tc16", fused=True)
        tr.core.fused_flags = flags
        for i in range(6):
            tr.step("s1", *batches[i % 4], 9990, W_S1, ALPHA, LR)
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record()
        for i in range(40):
            tr.step("s1", *batches[i % 4], 9990, W_S1, ALPHA, LR)
        t.record(); torch.cuda.synchronize()
        print(f"rep {rep} flags {flags}: {s.elapsed_time(t) / 40:7.4f} ms/step", flush=True)
