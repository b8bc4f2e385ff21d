"""Where the time of a whole 1024 x 1024 sphere-tracing frame (create_projectional_image, mean-curvature shading) goes."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffudf_b200 import SIREN, render_st  # noqa: E402

W = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "weights_trained.npz"))
m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(W[f"{'W' if k == 'weight' else 'b'}{i}"]) for i in range(9) for k in ("weight", "bias")})
m = m.cuda()
m.precision = sys.argv[1] if len(sys.argv) > 1 else "tcx3"
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
net_cfg = {"gt_mode": "tanh", "alpha": 100.0}
cfg = {"surface_threshold": 0.004, "max_iterations": 100, "gd_steps": 0, "height": R, "width": R, "light_position": [1, 2.38206, 10],
       "camera_position": cam.tolist(), "shininess": -1, "plot_curvatures": "mean", "curv_low_bound": 5, "curv_high_bound": 95,
       "reflection_method": "blinn-phong", "alpha1": 0.2, "alpha2": 0.2}
dev = torch.device("cuda:0")


def tick(label, t):
    torch.cuda.synchronize()
    now = time.perf_counter()
    print(f"  {label:34s} {1e3 * (now - t):8.2f} ms")
    return now


for rep in range(2):
    print("rep", rep)
    t0_np, mask = start.copy(), np.ones(R * R, dtype=bool)
    torch.cuda.synchronize()
    t = time.perf_counter()
    hits = render_st.propagate_rays(m, rays, t0_np, mask, net_cfg, cfg, dev)
    t = tick("propagate_rays (numpy in / out)", t)
    pts = torch.from_numpy(np.ascontiguousarray(t0_np[hits])).to(dev)
    rd = torch.from_numpy(np.ascontiguousarray(rays[hits])).to(dev)
    t = tick("hit points to the device", t)
    att = render_st.hit_attributes(m, pts, rd, curvature="mean")
    t = tick(f"hit_attributes ({pts.shape[0]} hits)", t)
    curv = att["mean"].double().cpu().numpy()[:, None]
    curv = np.clip(curv, np.percentile(curv, 5), np.percentile(curv, 95))
    curv -= curv.min(); curv /= curv.max()
    col = render_st._rdylbu(curv.squeeze(1))
    t = tick("percentile clip + colour map (host)", t)
    img = render_st.phong_shading(cfg["light_position"], -1, hits, t0_np, att["normals"].double(), color_map=col)
    t = tick("phong_shading (numpy in / out)", t)
