"""Golden fixture for BASELINE configs[0] ("beetle ... short run on CPU"), produced by the UNMODIFIED
reference modules on CPU:

    python tests/golden/make_golden_beetle.py

data/beetle/beetle.obj -> normalised as src/preprocess_mesh.py:5-15 (centre = vertex mean, scale 1/(1.1 max|coord|)) ->
10 000 area-weighted surface samples with triangle normals (numpy seed 0; Open3D's sample_points_uniformly is absent, the
sampling is restated) -> 12 batches of 2 997 rows from the reference's own sampleTrainingDataPC (numpy / torch seed 123 as
train.py:292-295) -> the schedule of train.py:167-191 compressed to 4 warm-up steps (loss_s1, lr 1e-4), 4 steps at lr_s1 = 1e-5,
4 steps of loss_s2 with the cosine lr, torch.optim.Adam, reference SIREN / loss_s1 / loss_s2.  Stored: the cloud, the batches,
the loss terms of every step, the parameter update (norm per tensor + a slice)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import import_reference, state_to_npz  # noqa: E402
from make_golden_sampler import import_dataset  # noqa: E402


def read_obj(path):
    V, F = [], []
    for line in open(path):
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            V.append([float(t) for t in p[1:4]])
        elif p[0] == "f":
            idx = [int(t.split("/")[0]) - 1 for t in p[1:]]
            for k in range(1, len(idx) - 1):
                F.append([idx[0], idx[k], idx[k + 1]])
    return np.array(V, np.float64), np.array(F, np.int64)


def main():
    ref = import_reference()
    ds = import_dataset()
    V, F = read_obj("/root/reference/data/beetle/beetle.obj")
    V = V - V.mean(0)
    V = V / (1.1 * np.abs(V).max())
    tri = V[F]
    cr = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    area = 0.5 * np.linalg.norm(cr, axis=1)
    rng = np.random.default_rng(0)
    t = rng.choice(len(F), size=10000, p=area / area.sum())
    r1, r2 = np.sqrt(rng.uniform(size=10000)), rng.uniform(size=10000)
    pts = ((1 - r1)[:, None] * tri[t, 0] + (r1 * (1 - r2))[:, None] * tri[t, 1] + (r1 * r2)[:, None] * tri[t, 2]).astype(np.float32)
    nrm = (cr[t] / np.maximum(np.linalg.norm(cr[t], axis=1, keepdims=True), 1e-30)).astype(np.float32)
    X = torch.from_numpy(pts.astype(np.float64))
    N = torch.from_numpy(nrm.astype(np.float64))
    X.get_device = lambda: "cpu"
    torch.manual_seed(123)
    np.random.seed(123)
    m = ref.model.SIREN(3, 1, [256] * 8, w0=30)
    init = {k: np.array(v, copy=True) for k, v in state_to_npz(m.state_dict()).items()}      # a copy: the arrays alias the live parameters
    n_on, n_off = int(3000 * 0.333), int(3000 * 0.666)
    E, S1, WU = 12, 8, 4
    opt = torch.optim.Adam(lr=1e-4, params=m.parameters())
    out = {"cloud_pts": pts, "cloud_nrm": nrm, "n_on": n_on, "n_off": n_off}
    xs, ns, dd, lrs = [], [], [], []
    for e in range(E):
        x, n, d = ds.sampleTrainingDataPC(X, N, n_on, n_off)
        if e >= S1:
            lr = 0.5 * (np.cos(e / (E - S1) * np.pi) + 1) * 1e-7
        elif e >= WU:
            lr = 1e-5
        else:
            lr = 1e-4
        for g in opt.param_groups:
            g["lr"] = lr
        opt.zero_grad()
        gt = {"normals": n, "sdf": d}
        loss = ref.lf.loss_s1(m, x, gt, [1e4, 1e4, 1e4, 1e3], 100) if e < S1 else ref.lf.loss_s2(m, x, gt, [1e5, 1e5], 100)
        sum(loss.values()).backward()
        opt.step()
        out[f"loss{e}"] = np.array([float(v) for v in loss.values()])
        xs.append(x.numpy()[0]); ns.append(n.numpy()[0]); dd.append(d.numpy()[0, :, 0]); lrs.append(lr)
        print(e, lr, out[f"loss{e}"])
    out["x"], out["normals"], out["d"], out["lr"] = np.stack(xs), np.stack(ns), np.stack(dd), np.array(lrs)
    fin = state_to_npz(m.state_dict())
    for k in fin:
        dlt = fin[k].astype(np.float64) - init[k].astype(np.float64)
        out["delta_" + k] = dlt.astype(np.float32)[..., :8].copy() if dlt.ndim == 2 else dlt.astype(np.float32)[:8].copy()
        out["dnorm_" + k] = np.array([np.linalg.norm(dlt)])
    path = os.path.join(HERE, "beetle_traj.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
