"""Golden vectors for the shading half of the sphere-tracing driver: the UNMODIFIED reference functions phong_shading and
ward_reflectance (/root/reference/src/render_st.py:174-245) on seeded random hit sets (build container only; the fixture travels).

    python tests/golden/make_golden_shading.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    import_reference()
    import src.render_st as rst
    rng = np.random.default_rng(11)
    R = 3000
    hits = rng.uniform(size=R) < 0.4
    H = int(hits.sum())
    samples = rng.uniform(-0.9, 0.9, (R, 3))
    nrm = rng.normal(size=(H, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pc1 = rng.normal(size=(H, 3))
    pc1 -= (pc1 * nrm).sum(1, keepdims=True) * nrm
    pc1 /= np.linalg.norm(pc1, axis=1, keepdims=True)
    pc2 = np.cross(nrm, pc1)
    cmap = rng.uniform(0, 1, (H, 3))
    light, cam = np.array([1.5, 1.0, 2.5]), np.array([0.8939, 0.7, 2.86])
    out = dict(hits=hits, samples=samples, normals=nrm, pc1=pc1, pc2=pc2, color_map=cmap, light=light, camera=cam)
    for name, cm in (("plain", None), ("cmap", cmap)):
        for sh in (0, 8, 40):
            out[f"phong_{name}_{sh}"] = rst.phong_shading(light, sh, hits, samples.copy(), nrm.copy(), color_map=None if cm is None else cm.copy())
        with np.errstate(all="ignore"):
            out[f"ward_{name}"] = rst.ward_reflectance(light, cam, hits, samples.copy(), nrm.copy(), alpha1=0.2, alpha2=0.5, pc1=pc1.copy(), pc2=pc2.copy(),
                                                       color_map=None if cm is None else cm.copy())
    np.savez_compressed(os.path.join(HERE, "shading.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
