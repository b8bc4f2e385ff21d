"""Three eager loss_s2 steps (for an ncu launch list: which of the small kernels make up the step?)
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches.csv python tools/s2_kernels.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ALPHA, LR, make_batches  # noqa: E402
from diffudf_b200 import SIREN  # noqa: E402
from diffudf_b200.train import FusedTrainer  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "s2"
torch.manual_seed(123)
tr = FusedTrainer(SIREN(3, 1, [256] * 8, w0=30).cuda(), precision="tcx3")
x, n, d = make_batches(1, 0)[0]
x, n, d = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x[0], n[0], d[0, :, 0]))
w = [1e5, 1e5] if mode == "s2" else [3e3, 1e2, 1e2, 5e1]
for _ in range(3):
    tr.step(mode, x, n, d, 9990, w, ALPHA, LR)
torch.cuda.synchronize()
print("done")
