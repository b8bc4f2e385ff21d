"""Times dudf_cap_mesh on 512^3 fields of an analytic two-sphere shape (both calls of the Python wrapper and the classification alone)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffudf_b200 import _lib  # noqa: E402
from diffudf_b200.render_mc import _cap_context, cap_triangles  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
# analytic field: two spheres (distance |r - r0|, negated normalised gradient), about as many surface cells as a trained shape
ax = torch.linspace(-1, 1, N, device="cuda")
X = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1)
r1 = torch.linalg.norm(X - torch.tensor([0.2, 0.1, 0.0], device="cuda"), dim=-1)
r2 = torch.linalg.norm(X + torch.tensor([0.4, 0.3, 0.2], device="cuda"), dim=-1)
d1, d2 = r1 - 0.55, r2 - 0.3
use1 = d1.abs() < d2.abs()
u = torch.where(use1, d1.abs(), d2.abs()).contiguous()
c = torch.where(use1[..., None], X - torch.tensor([0.2, 0.1, 0.0], device="cuda"), X + torch.tensor([0.4, 0.3, 0.2], device="cuda"))
g = (-torch.sign(torch.where(use1, d1, d2))[..., None] * c / torch.linalg.norm(c, dim=-1, keepdim=True).clamp_min(1e-9)).contiguous()
del X, r1, r2, d1, d2, use1, c
L = _lib.lib()
h = _cap_context(torch.device("cuda:0"))
n = ctypes.c_int64(0)


def timed(fn, reps=5):
    fn()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / reps


ms_cls = timed(lambda: _lib.check(L.dudf_cap_mesh(h, u.data_ptr(), g.data_ptr(), N, 0.008, None, 0, ctypes.byref(n), _lib.current_stream()), "cap"))
tris = torch.empty(int(n.value), 3, 3, device="cuda", dtype=torch.float64)
ms_emit = timed(lambda: _lib.check(L.dudf_cap_mesh(h, u.data_ptr(), g.data_ptr(), N, 0.008, tris.data_ptr(), int(n.value), ctypes.byref(n),
                                                   _lib.current_stream()), "cap"))
ms_all = timed(lambda: cap_triangles(u, g, N))
byt = 4 * N ** 3 + 2 * (N - 1) ** 3 + 72 * int(n.value)
print(f"N {N}: {int(n.value)} triangles; classify + count + scan + read-back {ms_cls:.3f} ms, emit {ms_emit:.3f} ms, wrapper {ms_all:.3f} ms; "
      f"algorithmic {byt / 1e9:.3f} GB -> {byt / (ms_cls + ms_emit) / 1e6:.0f} GB/s")
