"""the C++17 restatement of the reference's MeshUDF marching cubes (diffudf_b200/csrc/meshudf_mc.cpp,
diffudf_b200.marching_cubes.udf_mc_lewiner) against the reference's own Cython mesher compiled from its sources
(oracle/_ref, oracle/build_ref_mc.py): IDENTICAL vertex, face, normal and value arrays — same visit order, same float / double
arithmetic — on closed, open, noisy and deliberately unreliable fields, with steps, masks and non-cubic volumes.
The known answers of SURVEY 8c (analytic sphere: 64^3 -> 4 728 / 9 452, 128^3 -> 19 008 / 38 012) hold without the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref_mc  # noqa: E402

ref = build_ref_mc.load() if (build_ref_mc.built() or build_ref_mc.build()) else None
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref marching cubes not built (needs /root/reference once)")


def grid(shape):
    ax = [np.linspace(-1, 1, n) for n in shape]
    return np.stack(np.meshgrid(*ax, indexing="ij"), -1)


def fields(shape, kind, seed=0):
    """(df, vecs) float32: unsigned distance-like field and pseudo-normals, axes (z, y, x) like the reference's volumes"""
    P = grid(shape)
    rng = np.random.default_rng(seed)
    if kind == "disc":              # an OPEN surface: disc of radius 0.6 in the plane z = 0.03
        r = np.sqrt(P[..., 0] ** 2 + P[..., 1] ** 2)
        dz = P[..., 2] - 0.03
        dr = np.maximum(r - 0.6, 0)
        g = np.stack([np.where(dr > 0, dr * P[..., 0] / np.maximum(r, 1e-9), 0), np.where(dr > 0, dr * P[..., 1] / np.maximum(r, 1e-9), 0), dz], -1)
        g /= np.maximum(np.linalg.norm(g, axis=-1, keepdims=True), 1e-12)
        return np.sqrt(dr ** 2 + dz ** 2).astype(np.float32), (-g).astype(np.float32)
    if kind == "sphere":
        s = np.linalg.norm(P, axis=-1) - 0.5
    elif kind == "two_spheres":
        s = np.minimum(np.linalg.norm(P - [0.3, 0, 0], axis=-1) - 0.45, np.linalg.norm(P + [0.3, 0.1, 0], axis=-1) - 0.4)
    elif kind == "torus":
        s = np.sqrt((np.sqrt(P[..., 0] ** 2 + P[..., 2] ** 2) - 0.55) ** 2 + P[..., 1] ** 2) - 0.2
    elif kind == "noisy":
        s = np.linalg.norm(P, axis=-1) - 0.5 + 0.08 * np.sin(9 * P[..., 0]) * np.cos(7 * P[..., 1]) + 0.05 * np.sin(11 * P[..., 2] + 1)
    else:                            # "random": smooth random level set
        k, ph, a = rng.normal(size=(6, 3)) * 4, rng.uniform(0, 6, 6), rng.normal(size=6) * 0.15
        s = sum(a[i] * np.sin(P @ k[i] + ph[i]) for i in range(6)) + 0.02
    g = np.stack(np.gradient(s, *[2.0 / (n - 1) for n in shape]), -1)
    g /= np.maximum(np.linalg.norm(g, axis=-1, keepdims=True), 1e-12)
    vec = -g * np.sign(s)[..., None]
    if kind == "random":            # unreliable normals: exercises the unsure-case queue and the non-trivial Lewiner cases
        vec = vec + rng.normal(size=vec.shape) * 0.3
    return np.abs(s).astype(np.float32), vec.astype(np.float32)


def same(a, b):
    return all(x.shape == y.shape and x.dtype == y.dtype and np.array_equal(x, y) for x, y in zip(a, b))


CASES = [((40, 40, 40), "two_spheres"), ((48, 48, 48), "disc"), ((40, 40, 40), "torus"), ((56, 56, 56), "noisy"), ((40, 40, 40), "random"),
         ((33, 33, 33), "random"), ((64, 64, 64), "random"), ((24, 40, 31), "noisy"), ((31, 24, 40), "random")]


@needs_ref
@pytest.mark.parametrize("shape,kind", CASES)
@pytest.mark.parametrize("thresholds", [(1.05, 1.75), (2.0, 3.0)])
def test_port_equals_reference_mesher(shape, kind, thresholds):
    from diffudf_b200.marching_cubes import udf_mc_lewiner
    df, vecs = fields(shape, kind, seed=sum(shape))
    kw = dict(spacing=[2.0 / (shape[2] - 1)] * 3, avg_thresh=thresholds[0], max_thresh=thresholds[1])
    try:
        a = ref(df, vecs, **kw)
    except RuntimeError:
        with pytest.raises(RuntimeError):
            udf_mc_lewiner(df, vecs, **kw)
        return
    b = udf_mc_lewiner(df, vecs, **kw)
    assert a[0].shape[0] > 0 and same(a, b), (a[0].shape, a[1].shape, b[0].shape, b[1].shape)


@needs_ref
@pytest.mark.parametrize("step", [1, 2, 3])
def test_port_equals_reference_with_step_mask_and_ascent(step):
    from diffudf_b200.marching_cubes import udf_mc_lewiner
    df, vecs = fields((49, 49, 49), "noisy", seed=3)
    mask = np.zeros(df.shape, bool)
    mask[:, :30, :] = True
    for m in (None, mask):
        for direction in ("descent", "ascent"):
            kw = dict(spacing=(1., 1., 1.), gradient_direction=direction, step_size=step, avg_thresh=1.05 * step, max_thresh=1.75 * step, mask=m)
            try:
                a = ref(df, vecs, **kw)
            except RuntimeError:
                continue
            assert same(a, udf_mc_lewiner(df, vecs, **kw)), (step, m is not None, direction)


def test_known_answer_sphere_meshes():
    """analytic sphere UDF through extract_mesh_MESHUDF's call: 64^3 -> 4 728 / 9 452, 128^3 -> 19 008 / 38 012"""
    from diffudf_b200.marching_cubes import meshudf_from_fields
    for N, nv, nf in ((64, 4728, 9452), (128, 19008, 38012)):
        g = np.linspace(-1, 1, N, dtype=np.float32)
        P = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1)
        r = np.linalg.norm(P, axis=-1)
        v, f = meshudf_from_fields(np.abs(r - 0.5), -(P / np.maximum(r, 1e-9)[..., None]) * np.sign(r - 0.5)[..., None])
        assert v.shape == (nv, 3) and f.shape == (nf, 3)
        assert np.abs(np.linalg.norm(v, axis=1) - 0.5).max() < 2e-3
        assert f.min() == 0 and f.max() == nv - 1


def test_argument_checks_mirror_the_reference_wrapper():
    from diffudf_b200.marching_cubes import udf_mc_lewiner
    df, vecs = fields((8, 8, 8), "sphere")
    with pytest.raises(ValueError):
        udf_mc_lewiner(df[0], vecs)
    with pytest.raises(ValueError):
        udf_mc_lewiner(df[:1], vecs[:1])
    with pytest.raises(ValueError):
        udf_mc_lewiner(df, vecs, spacing=(1, 1))
    with pytest.raises(ValueError):
        udf_mc_lewiner(df, vecs, step_size=0)
    with pytest.raises(ValueError):
        udf_mc_lewiner(df, vecs, mask=np.ones((4, 4, 4), bool))
    with pytest.raises(RuntimeError):
        udf_mc_lewiner(np.ones((8, 8, 8), np.float32), vecs)          # no surface


@needs_ref
@pytest.mark.gpu
def test_port_equals_reference_mesher_on_the_trained_128_field(cuda_models):
    """the trained network's 128^3 fields from the CUDA path through both meshers: identical arrays"""
    import torch
    from diffudf_b200.marching_cubes import udf_mc_lewiner
    from diffudf_b200.render_mc import extract_fields
    m = cuda_models["trained"]
    m.precision = "fp32"
    N = 128
    df, vecs = extract_fields(m, None, N, "tanh", torch.device("cuda:0"), 100.0)
    df, vecs = df.cpu().numpy(), vecs.cpu().numpy()
    df[df < 0] = 0
    kw = dict(spacing=[2.0 / (N - 1)] * 3, avg_thresh=1.05, max_thresh=1.75)
    a, b = ref(df, vecs, **kw), udf_mc_lewiner(df, vecs, **kw)
    assert a[1].shape[0] > 40000 and same(a, b)
