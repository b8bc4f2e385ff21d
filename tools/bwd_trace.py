"""Event timeline of CTA 0 of the single-pass reverse sweep (tt_backward_kernel) inside a conforming (tcx3) training step:
MMA warp and epilogue warp 0, clock stamps per sub-tile and layer.
    python tools/bwd_trace.py [first pair to print] [pairs]"""
import os
import statistics as st
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ALPHA, LR, W_S1, make_batches  # noqa: E402
from diffudf_b200 import SIREN, _lib  # noqa: E402
from diffudf_b200.train import FusedTrainer  # noqa: E402

p0 = int(sys.argv[1]) if len(sys.argv) > 1 else 2
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(123)
model = SIREN(3, 1, [256] * 8, w0=30).cuda()
tr = FusedTrainer(model, precision="tcx3")
x, n, d = make_batches(1, 0)[0]
x, n, d = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x[0], n[0], d[0, :, 0]))
for _ in range(2):
    tr.step("s1", x, n, d, 9990, W_S1, ALPHA, LR)
L = _lib.lib()
buf = torch.zeros(2 * 8192, dtype=torch.int64, device="cuda")
# trace only the reverse sweep: run the forward untraced, then switch the trace on for the backward of the same step
core = tr.core
terms = core.forward("s1", x, n, d, 9990, W_S1, ALPHA, None, None)
torch.cuda.synchronize()
L.dudf_debug_set_trace(buf.data_ptr())
tr.grad_all.zero_()
core.backward(None, tr.gW, tr.gB)
torch.cuda.synchronize()
L.dudf_debug_set_trace(None)
raw = buf.cpu().numpy().astype("uint64")
NAMES = {1: "MMA  act_ready seen  s=", 2: "MMA  group issued    2+s=", 14: "EPI  pair start nch=", 40: "EPI  wait acc s=0 layer ", 41: "EPI  wait acc s=1 layer ",
         30: "EPI  start    s=0 layer ", 31: "EPI  start    s=1 layer ", 32: "EPI  done     s=0 layer ", 33: "EPI  done     s=1 layer "}
ev = []
for region in (0, 1):
    for v in raw[region * 8192:(region + 1) * 8192]:
        v = int(v)
        if v == 0:
            break
        ev.append((v & 0xFFFFFFFFFFFF, v >> 56, (v >> 48) & 0xFF))
ev.sort()
starts = [t for t, tag, _ in ev if tag == 14]
lo, hi = starts[p0], starts[p0 + npairs]
print(f"pairs {p0}..{p0 + npairs - 1}: {hi - lo} clocks ({(hi - lo) / npairs:.0f} per pair)")
prev = lo
work, wait = [], []
tw = ts = None
for t, tag, aux in ev:
    if lo <= t < hi:
        print(f"{t - lo:8d}  (+{t - prev:6d})  {NAMES.get(tag, str(tag))}{aux}")
        prev = t
    if tag in (40, 41):
        tw = t
    elif tag in (30, 31) and tw is not None:
        wait.append(t - tw); ts = t
    elif tag in (32, 33) and ts is not None:
        work.append(t - ts)
if work:
    print(f"reverse-sweep epilogue per sub-tile.layer: work median {st.median(work):.0f} clk, wait-for-accumulator median {st.median(wait):.0f} clk")
