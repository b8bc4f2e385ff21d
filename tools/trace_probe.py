"""Event timeline of CTA 0 of the tensor-core query kernel (MMA warp and epilogue warp 0): where a sub-tile pair's time goes.
    [DUDF_TC_CLUSTER=1] [DUDF_TC_REUSE=0] python tools/trace_probe.py [first pair to print] [pairs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffudf_b200 import SIREN, _lib  # noqa: E402

p0 = int(sys.argv[1]) if len(sys.argv) > 1 else 3
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(123)
m = SIREN(3, 1, [256] * 8, w0=30).cuda()
eng = m._engine_synced(2)
N = 128
buf = torch.zeros(2 * 8192, dtype=torch.int64, device="cuda")
eng.query_grid(N, 0, N ** 3, "tc16", 3, 100.0)
L = _lib.lib()
L.dudf_debug_set_trace(buf.data_ptr())
eng.query_grid(N, 0, N ** 3, "tc16", 3, 100.0)
torch.cuda.synchronize()
L.dudf_debug_set_trace(None)
raw = buf.cpu().numpy().astype("uint64")
NAMES = {1: "MMA  act_ready seen  s=", 2: "MMA  group issued    hs=", 10: "EPI  start s=0 layer ", 11: "EPI  start s=1 layer ",
         12: "EPI  done  s=0 layer ", 13: "EPI  done  s=1 layer ", 14: "EPI  pair start"}
ev = []
for region, who in ((0, "mma"), (1, "epi")):
    for v in raw[region * 8192:(region + 1) * 8192]:
        v = int(v)
        if v == 0:
            break
        ev.append((v & 0xFFFFFFFFFFFF, v >> 56, (v >> 48) & 0xFF))
ev.sort()
starts = [t for t, tag, _ in ev if tag == 14]
lo, hi = starts[p0], starts[p0 + npairs]
print(f"pair {p0}..{p0 + npairs - 1}: {hi - lo} clocks total ({(hi - lo) / npairs:.0f} per pair)")
prev = lo
for t, tag, aux in ev:
    if lo <= t < hi:
        print(f"{t - lo:8d}  (+{t - prev:6d})  {NAMES.get(tag, str(tag))}{aux if tag != 14 else ''}")
        prev = t
