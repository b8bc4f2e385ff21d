"""Times the fused tensor-core training step under its tuning flags (DUDF_FUSED_* bits) against the three-kernel route."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ALPHA, LR, W_S1, make_batches  # noqa: E402
from diffudf_b200 import SIREN  # noqa: E402
from diffudf_b200.train import FusedTrainer  # noqa: E402

batches = [tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x[0], n[0], d[0, :, 0])) for x, n, d in make_batches(4, 0)]
for name, fused, flags, w in (("unfused", False, 0, W_S1), ("fused16 flags=11", True, 11, W_S1), ("fused16 flags=11, no hessian term", True, 11, [1e4, 1e4, 0, 1e3]),
                              ("fused16, no images/scratch/weights/MMA (diag.)", True, 11 | (54 << 8), W_S1), ("fused flags=0", True, 0, W_S1), ("fused flags=1", True, 1, W_S1),
                              ("fused flags=2", True, 2, W_S1), ("fused flags=3", True, 3, W_S1), ("fused flags=3, no weight streaming (diagnostic)", True, 3 | (4 << 8), W_S1),
                              ("fused flags=3, no MMA (diagnostic)", True, 3 | (2 << 8), W_S1),
                              ("fused flags=3, no image copies (diagnostic)", True, 3 | (16 << 8), W_S1),
                              ("fused flags=3, no scratch traffic (diagnostic)", True, 3 | (32 << 8), W_S1),
                              ("fused flags=3, no images, no scratch (diagnostic)", True, 3 | (48 << 8), W_S1),
                              ("fused flags=3, no images/scratch/weights (diag.)", True, 3 | (52 << 8), W_S1),
                              ("fused flags=3, no images/scratch/weights/MMA", True, 3 | (54 << 8), W_S1),
                              ("unfused, no hessian term", False, 0, [1e4, 1e4, 0, 1e3]), ("fused flags=3, no hessian term", True, 3, [1e4, 1e4, 0, 1e3])):
    torch.manual_seed(123)
    model = SIREN(3, 1, [256] * 8, w0=30).cuda()
    tr = FusedTrainer(model, precision="tc16", fused=fused)
    tr.core.fused_flags = flags
    for i in range(4):
        tr.step("s1", *batches[i % 4], 9990, w, ALPHA, LR)
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for i in range(20):
        tr.step("s1", *batches[i % 4], 9990, w, ALPHA, LR)
    t.record()
    torch.cuda.synchronize()
    print(f"{name:52s}: {s.elapsed_time(t) / 20:7.4f} ms/step", flush=True)
