"""Active rays per sphere-tracing iteration of the bench's 1024 x 1024 frame, and the wall time of the native loop."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffudf_b200 import SIREN, render_st  # noqa: E402
from diffudf_b200.inverses import inverse_torch  # noqa: E402

torch.manual_seed(123)
m = SIREN(3, 1, [256] * 8, w0=30).cuda()
m.precision = "tc16"
R = 1024
cam = np.array([0.8939, 0.7, 2.86]) * 0.45
u, v = np.meshgrid(np.linspace(-0.6, 0.6, R), np.linspace(-0.6, 0.6, R))
d = np.stack([u.ravel(), v.ravel(), -np.ones(R * R)], 1)
d /= np.linalg.norm(d, axis=1, keepdims=True)
fwd = -cam / np.linalg.norm(cam)
right = np.cross(fwd, [0, 1.0, 0]); right /= np.linalg.norm(right)
up = np.cross(right, fwd)
rays = d[:, :1] * right + d[:, 1:2] * up - d[:, 2:3] * fwd
start = np.tile(cam, (R * R, 1)) + rays * 0.35
rays_d = torch.from_numpy(rays).cuda()
t0_d = torch.from_numpy(start).cuda()
idx = torch.arange(R * R, device="cuda")
eng = m._engine_synced()
t = t0_d.clone()
counts = []
ii = idx
for it in range(100):
    if ii.numel() == 0:
        break
    counts.append(int(ii.numel()))
    f, _, _, _ = eng.query(t[ii].float().contiguous(), 0, "tc16")
    st = inverse_torch("tanh", f.abs(), 100.0)
    pos = t[ii] + rays_d[ii] * st.double()[:, None]
    t[ii] = pos
    below = st.abs() < 0.004
    inside = ((pos > -1).all(1)) & ((pos < 1).all(1))
    ii = ii[(~below) & inside]
print("iterations", len(counts), "queries", sum(counts))
print("active per iteration:", counts)
for rep in range(3):
    tt = t0_d.clone()
    torch.cuda.synchronize()
    a = time.perf_counter()
    hits, _, nq = render_st._march(m, rays_d, tt, idx, "tanh", 100.0, 0.004, 100)
    torch.cuda.synchronize()
    print(f"native march: {(time.perf_counter() - a) * 1e3:.2f} ms, {nq} queries")
