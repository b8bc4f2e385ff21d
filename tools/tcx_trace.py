"""Event timeline of CTA 0 of the split-precision (tcx3) chain kernel: MMA warp and epilogue warp 0, clock stamps per
layer half.  Where does a tile·layer's time go (accumulator wait / epilogue work / hand-off)?
    python tools/tcx_trace.py [order 0|1|2] [first tile to print] [tiles]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffudf_b200 import SIREN, _lib  # noqa: E402

train = len(sys.argv) > 1 and sys.argv[1] == "train"        # python tools/tcx_trace.py train [first tile] [tiles]: the training forward
if train:
    sys.argv[1] = "2"
order = int(sys.argv[1]) if len(sys.argv) > 1 else 1
p0 = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ntiles = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.manual_seed(123)
m = SIREN(3, 1, [256] * 8, w0=30).cuda()
eng = m._engine_synced(2)
x = torch.rand(1 << 19, 3, device="cuda") * 2 - 1
buf = torch.zeros(2 * 8192, dtype=torch.int64, device="cuda")
L = _lib.lib()
if train:
    import numpy as np
    from bench import ALPHA, W_S1, make_batches
    from diffudf_b200.train import FusedTrainer
    tr = FusedTrainer(m, precision="tcx3")
    bx, bn, bd = make_batches(1, 0)[0]
    bx, bn, bd = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (bx[0], bn[0], bd[0, :, 0]))
    tr.core.forward("s1", bx, bn, bd, 9990, W_S1, ALPHA, None, None)
    torch.cuda.synchronize()
    L.dudf_debug_set_trace(buf.data_ptr())
    tr.core.forward("s1", bx, bn, bd, 9990, W_S1, ALPHA, None, None)
else:
    eng.query(x, order, "tcx3")
    L.dudf_debug_set_trace(buf.data_ptr())
    eng.query(x, order, "tcx3")
torch.cuda.synchronize()
L.dudf_debug_set_trace(None)
raw = buf.cpu().numpy().astype("uint64")
NAMES = {1: "MMA  act_ready seen kh=", 2: "MMA  group issued  kh*2+h=", 14: "EPI  tile start nch=", 15: "EPI  tile end", 20: "EPI  layers done",
         40: "EPI  wait acc h=0 layer ", 41: "EPI  wait acc h=1 layer ", 10: "EPI  start    h=0 layer ", 11: "EPI  start    h=1 layer ",
         12: "EPI  stored   h=0 layer ", 13: "EPI  stored   h=1 layer ", 30: "EPI  arrived  h=0 layer ", 31: "EPI  arrived  h=1 layer "}
ev = []
for region in (0, 1):
    for v in raw[region * 8192:(region + 1) * 8192]:
        v = int(v)
        if v == 0:
            break
        ev.append((v & 0xFFFFFFFFFFFF, v >> 56, (v >> 48) & 0xFF))
ev.sort()
starts = [t for t, tag, _ in ev if tag == 14]
lo, hi = starts[p0], starts[p0 + ntiles]
print(f"order {order}: tiles {p0}..{p0 + ntiles - 1}: {hi - lo} clocks total ({(hi - lo) / ntiles:.0f} per tile)")
prev = lo
# summary: per-half epilogue work (start -> stored) and waits
work, wait, hand = [], [], []
t_wait = t_start = None
for t, tag, aux in ev:
    if lo <= t < hi:
        print(f"{t - lo:8d}  (+{t - prev:6d})  {NAMES.get(tag, str(tag))}{aux}")
        prev = t
    if tag in (40, 41):
        t_wait = t
    elif tag in (10, 11) and t_wait is not None:
        wait.append(t - t_wait); t_start = t
    elif tag in (12, 13) and t_start is not None:
        work.append(t - t_start); t_st = t
    elif tag in (30, 31) and t_start is not None:
        hand.append(t - t_st)
import statistics as st
if work:
    print(f"epilogue half: work median {st.median(work):.0f} clk, wait-for-accumulator median {st.median(wait):.0f}, fence+arrive median {st.median(hand):.0f}")
