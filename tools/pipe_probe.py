"""Pipeline diagnostics of the tcgen05 query kernel: time the grid query with (a) everything, (b) MMAs skipped
(epilogue + protocol only), (c) epilogue math skipped (MMAs + weight streaming + protocol only).  Results of (b)/(c) are
numerically meaningless; only their durations matter."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffudf_b200 import SIREN  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(123)
m = SIREN(3, 1, [256] * 8, w0=30).cuda()
eng = m._engine_synced()
cnt = N ** 3
df = torch.empty(cnt, device="cuda")
vecs = torch.empty(cnt, 3, device="cuda")
H = None
for order_name, want_h in (("f+grad (4 ch)", False), ("f+grad+hess (10 ch)", True)):
    for name, dbg in (("full", 0), ("full, accumulators prefetched a group ahead", 64 << 8), ("no MMA", 2 << 8), ("no epilogue math", 1 << 8), ("neither", 3 << 8),
                      ("no weight streaming", 4 << 8), ("no weights, no epilogue math", 5 << 8),
                      ("no weights, no MMA", 6 << 8), ("no weights, neither", 7 << 8),
                      ("protocol only (no first/output layer either)", 15 << 8), ("full minus first/output layer", 8 << 8)):
        flags = 3 | dbg
        for rep in range(2):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            eng.query_grid(N, 0, cnt, "tc16", flags, 100.0, want_vecs=True, want_hess=want_h, out=(df, vecs))
            t.record()
            torch.cuda.synchronize()
        ms = s.elapsed_time(t)
        print(f"{order_name:22s} {name:46s}: {ms:8.3f} ms  {cnt / ms / 1e3:8.1f} M queries/s", flush=True)
