"""Short driver for ncu captures: a few fused training steps and one grid query per precision.
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 3 -o gpurun_out/prof \
        python tools/profile_step.py [steps] [gridN] [train precision: tcx3|tc16|fp32] [grid precisions, comma separated]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ALPHA, LR, W_S1, make_batches  # noqa: E402
from diffudf_b200 import SIREN  # noqa: E402
from diffudf_b200.train import FusedTrainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
gridN = int(sys.argv[2]) if len(sys.argv) > 2 else 128
prec = sys.argv[3] if len(sys.argv) > 3 else "tcx3"
grid_precs = (sys.argv[4] if len(sys.argv) > 4 else "tcx3").split(",")
torch.manual_seed(123)
model = SIREN(3, 1, [256] * 8, w0=30).cuda()
tr = FusedTrainer(model, precision=prec, fused=os.environ.get("DUDF_FUSED", "1") != "0")
tr.core.fused_flags = int(os.environ.get("DUDF_FUSED_FLAGS", "2"))
x, n, d = make_batches(1, 0)[0]
x, n, d = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x[0], n[0], d[0, :, 0]))
for _ in range(steps):
    tr.step("s1", x, n, d, 9990, W_S1, ALPHA, LR)
if gridN > 0:
    eng = model._engine_synced()
    for p in grid_precs:
        eng.query_grid(gridN, 0, gridN ** 3, p, 3, ALPHA)
torch.cuda.synchronize()
print("profile_step done")
