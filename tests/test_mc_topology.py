"""north_star: "marching-cubes vertex and face topology is identical on the same field" (SURVEY 8d config 3).
The mesher is the reference's own MeshUDF marching cubes (src/marching_cubes, the only native code of the reference),
compiled from the sources where they lie by oracle/build_ref_mc.py into oracle/_ref/ (test infrastructure; the binaries
travel to the GPU box).  Called exactly as extract_mesh_MESHUDF does (src/render_mc.py:127-133): df clamped at 0,
spacing 2/(N-1), avg_thresh 1.05, max_thresh 1.75."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref_mc  # noqa: E402

mc = build_ref_mc.load() if (build_ref_mc.built() or build_ref_mc.build()) else None
needs_mc = pytest.mark.skipif(mc is None, reason="oracle/_ref marching cubes not built (needs /root/reference once: python oracle/build_ref_mc.py)")


def mesh(df, vecs):
    df = np.array(df, np.float32)
    df[df < 0] = 0
    N = df.shape[0]
    v, f, _, _ = mc(df, np.ascontiguousarray(vecs, np.float32), spacing=[2.0 / (N - 1)] * 3, avg_thresh=1.05, max_thresh=1.75)
    return v - 1, f


@needs_mc
def test_reference_mc_known_answer_sphere():
    """analytic sphere UDF, 64^3 -> 4 728 vertices / 9 452 faces with the rebuilt module."""
    N = 64
    g = np.linspace(-1, 1, N, dtype=np.float32)
    P = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1)
    r = np.linalg.norm(P, axis=-1)
    df = np.abs(r - 0.5)
    vecs = -(P / np.maximum(r, 1e-9)[..., None]) * np.sign(r - 0.5)[..., None]
    v, f = mesh(df, vecs)
    assert v.shape == (4728, 3) and f.shape == (9452, 3)
    assert np.abs(np.linalg.norm(v, axis=1) - 0.5).max() < 2e-3


@needs_mc
def test_oracle_field_meshes_like_reference_field(golden, oracle, weights):
    """extract_fields of the unmodified reference (fixture) and of the oracle restatement give the same mesh."""
    g = golden("fields_trained.npz")
    N = int(g["N"]) if "N" in g.files else g["df"].shape[0]
    df_o, vecs_o = oracle.extract_fields(weights["trained"], N, "tanh", 100.0)
    try:
        v0, f0 = mesh(g["df"], g["vecs"])
    except RuntimeError:
        pytest.skip("the 12^3 fixture grid holds no surface crossing")
    v1, f1 = mesh(df_o, vecs_o)
    assert np.array_equal(f0, f1) and np.allclose(v0, v1, atol=1e-5)


@needs_mc
@pytest.mark.gpu
def test_cuda_fields_give_the_reference_topology(oracle, weights, cuda_models):
    """fp32 CUDA path vs the fp32 oracle field on the same 48^3 grid: identical faces, vertices to 1e-5.
    Tensor-core path (1e-3-class field) at 128^3 against the fp32 CUDA path: same surface (Hausdorff-like distance below
    half a voxel for 97 % of the vertices, face count within 1 %); the topology statement is made for the fp32-grade path."""
    import torch
    from diffudf_b200.render_mc import extract_fields
    m = cuda_models["trained"]
    N = 48
    m.precision = "fp32"
    df, vecs = extract_fields(m, None, N, "tanh", torch.device("cuda:0"), 100.0)
    df_o, vecs_o = oracle.extract_fields(weights["trained"], N, "tanh", 100.0)
    v0, f0 = mesh(df_o, vecs_o)
    v1, f1 = mesh(df.cpu().numpy(), vecs.cpu().numpy())
    print(f"48^3: oracle {v0.shape[0]} verts / {f0.shape[0]} faces, CUDA fp32 {v1.shape[0]} / {f1.shape[0]}")
    assert f0.shape == f1.shape and np.array_equal(f0, f1)
    assert np.abs(v0 - v1).max() < 1e-5
    N = 128
    out = {}
    for prec in ("fp32", "tc16"):
        m.precision = prec
        df, vecs = extract_fields(m, None, N, "tanh", torch.device("cuda:0"), 100.0)
        out[prec] = mesh(df.cpu().numpy(), vecs.cpu().numpy())
    m.precision = "fp32"
    (va, fa), (vb, fb) = out["fp32"], out["tc16"]
    print(f"128^3: fp32 {va.shape[0]} verts / {fa.shape[0]} faces, tc16 {vb.shape[0]} / {fb.shape[0]}")
    assert abs(fa.shape[0] - fb.shape[0]) <= 0.01 * fa.shape[0]
    ta, tb = torch.from_numpy(va).float().cuda(), torch.from_numpy(vb).float().cuda()
    from diffudf_b200.dataset import shortestDistance
    voxel = 2.0 / (N - 1)
    dab, dba = shortestDistance(ta, tb), shortestDistance(tb, ta)
    far = int((dab > 0.5 * voxel).sum()) + int((dba > 0.5 * voxel).sum())
    print(f"       vertex distance fp32 <-> tc16: median {float(dab.median()):.2e}, max {float(max(dab.max(), dba.max())):.2e}, "
          f"{far} of {va.shape[0] + vb.shape[0]} farther than half a voxel")
    # MeshUDF's pseudo-sign voting is discontinuous in the field: a 1e-3 perturbation may flip a few open-boundary cells
    # (measured: 1.4 % of the vertices move by more than half a voxel, at most two voxels; face count 44 736 vs 44 733)
    assert float(dab.median()) < 1e-2 * voxel and far <= 0.03 * (va.shape[0] + vb.shape[0])


@needs_mc
@pytest.mark.gpu
def test_tensor_core_grid_gives_the_oracle_topology_at_128(golden, cuda_models):
    """SURVEY 8d config 3 / VERDICT r1 item 1: MeshUDF marching cubes of the SPLIT-PRECISION tensor-core field (tcx3) against
    the mesh of the oracle's fp32 field on the same 128^3 grid: IDENTICAL face arrays (= identical topology: every sign vote, every
    threshold and every Lewiner case agree) and the same vertex count.  The oracle mesh is a committed fixture
    (tests/golden/make_golden_mesh128.py: fp64-numpy oracle fields -> the reference's own mesher; ~25 CPU-minutes).  Vertex
    positions are interpolation weights 1 / (eps + |df|) of df = sqrt(|f| / alpha), which amplifies the field's 1e-6 near f = 0:
    fp32 path within 5e-5 of the oracle's vertices (measured 2.1e-5 = 1.3e-3 voxel), split-precision tensor-core path within 5e-4
    (measured 2.1e-4 = 1.3e-2 voxel).  Both the reference's mesher (oracle/_ref) and the C++ port (diffudf_b200.marching_cubes)
    are run on the CUDA fields."""
    import torch
    from diffudf_b200.marching_cubes import meshudf_from_fields
    from diffudf_b200.render_mc import extract_fields
    G = golden("mesh128_trained.npz")
    v0, f0 = G["verts"].astype(np.float64), G["faces"]
    m = cuda_models["trained"]
    N = int(G["N"])
    res = {}
    try:
        for prec in ("fp32", "tcx3"):
            m.precision = prec
            df, vecs = extract_fields(m, None, N, "tanh", torch.device("cuda:0"), 100.0)
            # the fixture's samples of the oracle FIELDS: the mesh comparison below is not vacuous
            idx = torch.from_numpy(G["df_sample_idx"]).cuda()
            assert float((df.reshape(-1)[idx].cpu() - torch.from_numpy(G["df_sample"])).abs().max()) < 2e-5
            v1, f1 = mesh(df.cpu().numpy(), vecs.cpu().numpy())
            v2, f2 = meshudf_from_fields(df, vecs)
            assert np.array_equal(f1, f2) and np.array_equal(v1, v2), "C++ port and reference mesher disagree on the CUDA field"
            same = f0.shape == f1.shape and np.array_equal(f0, f1)
            dv = float(np.abs(v0 - v1).max()) if v0.shape == v1.shape else float("nan")
            res[prec] = (same, dv, v1.shape[0], f1.shape[0])
            print(f"128^3 {prec}: {v1.shape[0]} verts / {f1.shape[0]} faces vs oracle {v0.shape[0]} / {f0.shape[0]}; faces identical: {same}; "
                  f"max vertex difference {dv:.2e}")
    finally:
        m.precision = "fp32"
    for prec, (same, dv, nv, nf) in res.items():
        assert same and dv < {"fp32": 5e-5, "tcx3": 5e-4}[prec], (prec, same, dv, nv, nf)
