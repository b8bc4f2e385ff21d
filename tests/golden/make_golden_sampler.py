"""Golden fixture for the point-cloud batch sampler: runs the UNMODIFIED reference
`sampleTrainingDataPC` / `shortestDistance` (/root/reference/src/dataset.py:72-131) on CPU.

    python tests/golden/make_golden_sampler.py

Open3D is absent here, so the module's top-level imports are stubbed (the PC path does not touch Open3D); the reference
takes its device from `surface_pc.get_device()`, which is -1 on CPU tensors, so the fixture tensors carry an
instance-level `get_device` returning "cpu".  The random draws are re-created in the fixture (same numpy / torch seeds,
same call order as the reference) so that a restatement can be fed the identical draws."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_dataset():
    o3d = sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    o3c = sys.modules.setdefault("open3d.core", types.ModuleType("open3d.core"))
    o3c.Tensor = object
    o3d.core = o3c
    o3d.t = types.SimpleNamespace(geometry=types.SimpleNamespace(PointCloud=object))
    sys.path.insert(0, REF)
    import src.dataset as dataset
    return dataset


def main():
    from diffudf_b200 import synthetic
    ds = import_dataset()
    shape = synthetic.make_shape(0)
    pts, nrm = shape.sample_surface(6000, np.random.default_rng(1))
    pts, nrm = pts.astype(np.float32), nrm.astype(np.float32)      # the cloud as the device path sees it
    X = torch.from_numpy(pts.astype(np.float64))
    N = torch.from_numpy(nrm.astype(np.float64))
    X.get_device = lambda: "cpu"
    n_on, n_off = 333, 667
    np.random.seed(321)
    torch.manual_seed(321)
    coords, normals, sdf = ds.sampleTrainingDataPC(X, N, n_on, n_off)
    # the same draws, in the reference's call order (dataset.py:89, 98-101, 106, 110)
    np.random.seed(321)
    torch.manual_seed(321)
    on_idx = np.random.randint(0, X.shape[0], n_on)
    n_far = n_off // 2
    n_near = n_off - n_far
    far = np.random.uniform([-1, -1, -1], [1, 1, 1], (n_far, 3))
    near_idx = np.random.randint(0, n_on, n_near)
    near_off = torch.normal(0, 0.01, (n_near, 1)).numpy()
    # shortestDistance alone, fp64 and fp32
    P = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, (500, 3)))
    sd64 = ds.shortestDistance(P, X).numpy()
    sd32 = ds.shortestDistance(P.float(), X.float()).numpy()
    out = os.path.join(HERE, "sampler_pc.npz")
    np.savez_compressed(out, surf_pts=pts.astype(np.float32), surf_nrm=nrm.astype(np.float32), n_on=n_on, n_off=n_off,
                        on_idx=on_idx, far=far, near_idx=near_idx, near_off=near_off,
                        coords=coords.numpy(), normals=normals.numpy(), sdf=sdf.numpy(),
                        sd_queries=P.numpy(), sd64=sd64, sd32=sd32)
    print("wrote", out, os.path.getsize(out), "bytes; batch", tuple(coords.shape), tuple(normals.shape), tuple(sdf.shape))


if __name__ == "__main__":
    main()
