"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference modules
from /root/reference on CPU (this only works in the build container; the fixtures travel).

    python tests/golden/make_golden.py [--train-steps 300]

The reference holds no golden vectors of its own, so these outputs of the
reference itself are what pins the oracle (oracle/dudf_oracle.py) and, through it, the CUDA path.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    """Import the reference's hot-path modules read-only, stubbing the absent third-party
    packages they import at module top."""
    np.bool8 = np.bool_
    for name in ["open3d", "open3d.core", "trimesh", "mcubes", "skimage", "skimage.measure",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "_marching_cubes_lewiner", "tqdm"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["skimage.measure"].marching_cubes = None
    sys.modules["_marching_cubes_lewiner"].udf_mc_lewiner = None
    sys.modules["tqdm"].tqdm = lambda it, *a, **k: it
    sys.modules["open3d"].core = sys.modules["open3d.core"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    sys.path.insert(0, REF)
    import src.model as model
    import src.diff_operators as dif
    import src.loss_functions as lf
    import src.evaluate as ev
    import src.inverses as inv
    import src.render_mc as rmc
    import src.render_st as rst
    import src.render_pc as rpc
    return types.SimpleNamespace(model=model, dif=dif, lf=lf, ev=ev, inv=inv, rmc=rmc, rst=rst, rpc=rpc)


def state_to_npz(sd):
    out = {}
    n = len([k for k in sd if k.endswith(".0.weight")])
    for i in range(n):
        out[f"W{i}"] = sd[f"net.{i}.0.weight"].detach().cpu().numpy()
        out[f"b{i}"] = sd[f"net.{i}.0.bias"].detach().cpu().numpy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train-steps", type=int, default=300)
    ap.add_argument("--train-rows", type=int, default=6000)
    args = ap.parse_args()
    ref = import_reference()
    from diffudf_b200 import synthetic
    dev = torch.device("cpu")
    torch.set_num_threads(os.cpu_count())

    # ---- weights: seed 123 init exactly as train.py:292-317 ----
    torch.manual_seed(123)
    np.random.seed(123)
    net = ref.model.SIREN(3, 1, [256] * 8, w0=30)
    np.savez(os.path.join(HERE, "weights_init.npz"), **state_to_npz(net.state_dict()))

    # ---- synthetic shape + a short CPU training with the reference losses ----
    shape = synthetic.make_shape(0)
    rng = np.random.default_rng(0)
    surf_p, surf_n = shape.sample_surface(200000, rng)
    alpha = 100.0
    tnet = ref.model.SIREN(3, 1, [256] * 8, w0=30)
    tnet.load_state_dict(net.state_dict())
    opt = torch.optim.Adam(lr=1e-4, params=tnet.parameters())
    brng = np.random.default_rng(1)
    w_s1 = [1e4, 1e4, 1e4, 1e3]
    for step in range(args.train_steps):
        x, nrm, d = synthetic.make_batch(shape, surf_p, surf_n, args.train_rows, (0.333, 0.666), brng)
        opt.zero_grad()
        loss = ref.lf.loss_s1(tnet, torch.from_numpy(x), {"normals": torch.from_numpy(nrm), "sdf": torch.from_numpy(d)}, w_s1, alpha)
        tot = sum(loss.values())
        tot.backward()
        opt.step()
        if step % 20 == 0:
            print("train", step, {k: float(v) for k, v in loss.items()}, flush=True)
    np.savez(os.path.join(HERE, "weights_trained.npz"), **state_to_npz(tnet.state_dict()))

    for tag, model in (("init", net), ("trained", tnet)):
        model.eval()
        # ---- jets: f, grad, hessian by the reference operators, fp32 and fp64 ----
        prng = np.random.default_rng(7)
        sp, _ = shape.sample_surface(128, prng)
        pts = np.concatenate([prng.uniform(-1, 1, (384, 3)).astype(np.float32), sp], 0)
        out = {"x": pts}
        for dt, name in ((torch.float32, "32"), (torch.float64, "64")):
            m = ref.model.SIREN(3, 1, [256] * 8, w0=30).to(dt)
            m.load_state_dict({k: v.to(dt) for k, v in model.state_dict().items()})
            o = m(torch.from_numpy(pts).to(dt).unsqueeze(0))
            xin, y = o["model_in"], o["model_out"]
            g = ref.dif.gradient(y, xin)
            H = ref.dif.hessian(y.squeeze(-1), xin)
            out["f" + name] = y.detach().numpy()[0, :, 0]
            out["g" + name] = g.detach().numpy()[0]
            out["H" + name] = H.detach().numpy()[0]
        np.savez(os.path.join(HERE, f"jets_{tag}.npz"), **out)

        # ---- losses + parameter gradients on a small batch ----
        brng2 = np.random.default_rng(11)
        x, nrm, d = synthetic.make_batch(shape, surf_p, surf_n, 600, (0.333, 0.666), brng2)
        out = {"x": x, "normals": nrm, "d": d}
        for dt, name in ((torch.float32, "32"), (torch.float64, "64")):
            for mode in ("s1", "s1_nohess", "s2", "siren"):
                m = ref.model.SIREN(3, 1, [256] * 8, w0=30).to(dt)
                m.load_state_dict({k: v.to(dt) for k, v in model.state_dict().items()})
                gt = {"normals": torch.from_numpy(nrm).to(dt), "sdf": torch.from_numpy(d).to(dt)}
                xi = torch.from_numpy(x).to(dt)
                if mode == "s1":
                    loss = ref.lf.loss_s1(m, xi, gt, [1e4, 1e4, 1e4, 1e3], alpha)
                elif mode == "s1_nohess":
                    loss = ref.lf.loss_s1(m, xi, gt, [1e4, 1e4, 0, 1e3], alpha)
                elif mode == "s2":
                    loss = ref.lf.loss_s2(m, xi, gt, [1e5, 1e5], alpha)
                else:
                    loss = ref.lf.loss_siren(m, xi, gt, [3e3, 1e2, 1e2, 5e1])
                tot = 0
                for k, v in loss.items():
                    tot = tot + v
                    out[f"{mode}_{name}_{k}"] = np.asarray(v.detach().numpy()).reshape(-1)[:1]
                tot.backward()
                if name == "64":      # fp64 parameter gradients: per-tensor norm + strided subsample
                    for i in range(9):
                        gW = m.net[i][0].weight.grad.numpy().reshape(-1)
                        gb = m.net[i][0].bias.grad.numpy().reshape(-1)
                        out[f"{mode}_gWnorm{i}"] = np.array([np.linalg.norm(gW)])
                        out[f"{mode}_gbnorm{i}"] = np.array([np.linalg.norm(gb)])
                        out[f"{mode}_gWsub{i}"] = gW[::37].copy()
                        out[f"{mode}_gbsub{i}"] = gb[::5].copy()
        np.savez_compressed(os.path.join(HERE, f"losses_{tag}.npz"), **out)

        # ---- evaluate(): chunked query, crosses the 4096 chunk boundary ----
        prng = np.random.default_rng(13)
        samples = prng.uniform(-1, 1, (5000, 3)).astype(np.float32)
        grads = np.zeros((5000, 3))
        hess = np.zeros((5000, 3, 3))
        f = ref.ev.evaluate(model, torch.from_numpy(samples), device=dev, gradients=grads, hessians=hess)
        np.savez_compressed(os.path.join(HERE, f"evaluate_{tag}.npz"), x=samples, f=f, g=grads, H=hess)

        # ---- extract_fields N=12 ----
        df, vecs = ref.rmc.extract_fields(model, torch.Tensor([[]]), 12, "tanh", dev, alpha)
        np.savez_compressed(os.path.join(HERE, f"fields_{tag}.npz"), df=df.numpy(), vecs=vecs.numpy())

    # ---- sphere tracing + normals + curvature on the trained net (24x24 rays) ----
    model = tnet
    R = 24
    cam = np.array([0.8939, 0.7, 2.86]) * 0.45
    u, v = np.meshgrid(np.linspace(-0.5, 0.5, R), np.linspace(-0.5, 0.5, R))
    dirs = np.stack([u.ravel(), v.ravel(), -np.ones(R * R)], 1)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    fwd = -cam / np.linalg.norm(cam)
    right = np.cross(fwd, [0, 1.0, 0]); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    rays = dirs[:, :1] * right + dirs[:, 1:2] * up - dirs[:, 2:3] * fwd
    t0 = np.tile(cam, (R * R, 1)).astype(np.float64)
    # march to the unit box first so that every start lies inside (-1,1)^3
    t0 = t0 + rays * 0.35
    mask = np.ones(R * R, dtype=bool)
    netcfg = {"gt_mode": "tanh", "alpha": alpha}
    rcfg = {"max_iterations": 100, "surface_threshold": 0.004}
    t0_in = t0.copy()
    hits = ref.rst.propagate_rays(model, rays, t0, mask.copy(), netcfg, rcfg, dev)
    out = {"rays": rays, "t0_in": t0_in, "t0_out": t0, "hits": hits}
    for dt, name in ((torch.float32, "32"), (torch.float64, "64")):
        m = ref.model.SIREN(3, 1, [256] * 8, w0=30).to(dt)
        m.load_state_dict({k: v.to(dt) for k, v in model.state_dict().items()})
        xin = torch.from_numpy(t0[hits]).to(dt).unsqueeze(0)
        o = m(xin)
        n, pcd = ref.rst.compute_normals_and_cd(o["model_in"], o["model_out"])
        mean = ref.rst.compute_curvature(o["model_in"], n, "mean", dev)
        out["n" + name] = n.detach().numpy()[0]
        out["dirs" + name] = pcd.numpy()[0]
        out["mean" + name] = mean.numpy().reshape(-1)
        # gaussian curvature: restate the 4x4 determinant with the reference jacobian
        J, _ = ref.dif.jacobian(n, o["model_in"])
        ext = torch.zeros((J.shape[1], 4, 4), dtype=dt)
        ext[:, :3, :3] = J[0]
        ext[:, :3, 3] = n[0]
        ext[:, 3, :3] = n[0]
        out["gauss" + name] = (-torch.linalg.det(ext)).detach().numpy()
        out["J" + name] = J.detach().numpy()[0]
    np.savez_compressed(os.path.join(HERE, "rays_trained.npz"), **out)

    # ---- NDF-style projection (render_pc.py:26-73) on 1500 seeds, one outer iteration ----
    class _S(ref.rpc.Sampler):
        def __init__(self, dec):
            self.decoder = dec
            self.features = 3
            self.device = dev
    smp = _S(model)
    np.random.seed(5)
    seeds_state = np.random.get_state()
    pts, nrm = smp.generate_point_cloud("tanh", alpha, num_steps=3, num_points=1500, surf_thresh=0.007, max_iter=1)
    np.random.set_state(seeds_state)
    seeds = np.random.uniform(-1, 1, (1500, 3))
    np.savez_compressed(os.path.join(HERE, "pc_trained.npz"), seeds=seeds, points=pts, normals=nrm)

    # ---- short optimisation trajectory: 4 x loss_s1 (lr 1e-4) then 3 x loss_s2 (lr 1e-7) ----
    m = ref.model.SIREN(3, 1, [256] * 8, w0=30)
    m.load_state_dict(net.state_dict())
    opt = torch.optim.Adam(lr=1e-4, params=m.parameters())
    brng3 = np.random.default_rng(21)
    traj = {}
    batches = []
    for step in range(7):
        x, nrm, d = synthetic.make_batch(shape, surf_p, surf_n, 900, (0.333, 0.666), brng3)
        batches.append((x, nrm, d))
        if step == 4:
            for gph in opt.param_groups:
                gph["lr"] = 1e-7
        opt.zero_grad()
        gt = {"normals": torch.from_numpy(nrm), "sdf": torch.from_numpy(d)}
        if step < 4:
            loss = ref.lf.loss_s1(m, torch.from_numpy(x), gt, w_s1, alpha)
        else:
            loss = ref.lf.loss_s2(m, torch.from_numpy(x), gt, [1e5, 1e5], alpha)
        tot = sum(loss.values())
        tot.backward()
        opt.step()
        traj[f"loss{step}"] = np.array([float(v) for v in loss.values()])
    traj["x"] = np.stack([b[0] for b in batches])
    traj["normals"] = np.stack([b[1] for b in batches])
    traj["d"] = np.stack([b[2] for b in batches])
    init = state_to_npz(net.state_dict())
    fin = state_to_npz(m.state_dict())
    for k in fin:
        traj["delta_" + k] = (fin[k].astype(np.float64) - init[k].astype(np.float64)).astype(np.float32)[..., :8].copy() \
            if fin[k].ndim == 2 else (fin[k].astype(np.float64) - init[k].astype(np.float64)).astype(np.float32)[:8].copy()
        traj["dnorm_" + k] = np.array([np.linalg.norm(fin[k].astype(np.float64) - init[k].astype(np.float64))])
    np.savez_compressed(os.path.join(HERE, "trajectory_init.npz"), **traj)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
