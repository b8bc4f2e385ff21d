"""Parity of the field queries (f, grad f, Hessian, third derivatives) against the fp64 oracle and the
reference's own fp32 outputs (golden fixtures).  Tolerances (max|err| / max|ref|, BASELINE.md's measure):
fp32 CUDA-core path 1e-5 (north_star's fp32 tolerance; measured 2e-7..5e-6, the reference's own fp32 run sits at 2.6e-6
from fp64).  The single-pass fp16-operand tcgen05 path (tc16) is the NON-conforming fast mode: 1e-3 in relative L2 on
trained weights, 2.5e-3 at the SIREN init, 4e-3 in the max measure — documented, not the bar; the conforming tensor-core
mode is tcx3 (tests/test_gpu_tcx3.py)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_max

pytestmark = pytest.mark.gpu


def _query(model, x, order, precision):
    eng = model._engine_synced()
    xt = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
    f, g, H, T = eng.query(xt, order, precision)
    torch.cuda.synchronize()
    c = lambda t: None if t is None else t.cpu().numpy()
    return c(f), c(g), c(H), c(T)


@pytest.mark.parametrize("tag", ["init", "trained"])
@pytest.mark.parametrize("order", [0, 1, 2, 3])
def test_fp32_query_matches_oracle(tag, order, golden, oracle, weights, cuda_models):
    J = golden(f"jets_{tag}.npz")
    ref = oracle.siren_jet(weights[tag], J["x"], order)
    f, g, H, T = _query(cuda_models[tag], J["x"], order, "fp32")
    e = {"f": rel_max(f, ref["f"])}
    if order >= 1:
        e["g"] = rel_max(g, ref["g"])
    if order >= 2:
        e["H"] = rel_max(H, ref["H"])
    if order >= 3:
        Tr = ref["T"]
        idx = [(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 1, 1), (0, 1, 2), (0, 2, 2), (1, 1, 1), (1, 1, 2), (1, 2, 2), (2, 2, 2)]
        Tref = np.stack([Tr[:, a, b, c] for a, b, c in idx], 1)
        e["T"] = rel_max(T, Tref)
    print(f"fp32 {tag} order {order}: {e}")
    assert all(v < 1e-5 for v in e.values()), e


@pytest.mark.parametrize("tag", ["init", "trained"])
def test_fp32_query_matches_reference_fixture(tag, golden, cuda_models):
    J = golden(f"jets_{tag}.npz")
    f, g, H, _ = _query(cuda_models[tag], J["x"], 2, "fp32")
    assert rel_max(f, J["f32"]) < 1e-5 and rel_max(g, J["g32"]) < 1e-5 and rel_max(H, J["H32"]) < 1e-5


@pytest.mark.parametrize("tag", ["init", "trained"])
@pytest.mark.parametrize("order", [0, 1, 2])
def test_tc16_query_matches_oracle(tag, order, golden, oracle, weights, cuda_models):
    J = golden(f"jets_{tag}.npz")
    ref = oracle.siren_jet(weights[tag], J["x"], order)
    f, g, H, _ = _query(cuda_models[tag], J["x"], order, "tc16")
    e = {"f": (rel_max(f, ref["f"]), rel_l2(f, ref["f"]))}
    if order >= 1:
        e["g"] = (rel_max(g, ref["g"]), rel_l2(g, ref["g"]))
    if order >= 2:
        e["H"] = (rel_max(H, ref["H"]), rel_l2(H, ref["H"]))
    print(f"tc16 {tag} order {order}: (max-measure, rel-L2) {e}")
    # single-pass fp16 operands carry an 11-bit significand, the same as tf32: BASELINE.md's probe puts that class at
    # f 5.7e-4 / grad 1.6e-3 / hess 1.9e-3 at the SIREN init ("borderline at 1e-3"); trained weights sit below 1e-3.
    lim = 1e-3 if tag == "trained" else 2.5e-3
    for k, (emax, el2) in e.items():
        assert el2 < lim, (k, el2)
        assert emax < 4e-3, (k, emax)


@pytest.mark.parametrize("precision", ["fp32", "tc16"])
@pytest.mark.parametrize("P", [1, 7, 63, 64, 65, 257, 4096 + 33])
def test_ragged_sizes(P, precision, oracle, weights, cuda_models):
    rng = np.random.default_rng(P)
    x = rng.uniform(-1, 1, (P, 3)).astype(np.float32)
    ref = oracle.siren_jet(weights["trained"], x, 1)
    f, g, _, _ = _query(cuda_models["trained"], x, 1, precision)
    tol = 1e-5 if precision == "fp32" else 4e-3
    assert np.max(np.abs(f - ref["f"])) <= tol * max(np.max(np.abs(ref["f"])), 1e-3)
    assert np.max(np.abs(g - ref["g"])) <= tol * np.max(np.abs(ref["g"]))


def test_empty_query(cuda_models):
    eng = cuda_models["init"]._engine_synced()
    f, g, _, _ = eng.query(torch.empty(0, 3, device="cuda"), 1, "fp32")
    assert f.shape == (0,) and g.shape == (0, 3)


def test_shallow_network_and_state_dict_roundtrip(oracle):
    """Sampler's default config is 4 hidden layers (render_pc.py:12); keys must round-trip."""
    from diffudf_b200 import SIREN
    torch.manual_seed(3)
    m = SIREN(3, 1, [256] * 4, w0=30).cuda()
    sd = m.state_dict()
    assert list(sd.keys()) == [f"net.{i}.0.{k}" for i in range(5) for k in ("weight", "bias")]
    params = oracle.params_from_state_dict(sd)
    x = np.random.default_rng(0).uniform(-1, 1, (300, 3)).astype(np.float32)
    ref = oracle.siren_jet(params, x, 2)
    for prec, tol in (("fp32", 1e-5), ("tcx3", 2e-5), ("tc16", 4e-3)):
        f, g, H, _ = _query(m, x, 2, prec)
        assert rel_max(f, ref["f"]) < tol and rel_max(g, ref["g"]) < tol and rel_max(H, ref["H"]) < tol
    m2 = SIREN(3, 1, [256] * 4, w0=30, delay_init=True).cuda()
    m2.load_state_dict(sd)
    f2, _, _, _ = _query(m2, x, 0, "fp32")
    f1, _, _, _ = _query(m, x, 0, "fp32")
    assert np.array_equal(f1, f2)


@pytest.mark.parametrize("tag", ["init", "trained"])
def test_dropin_forward_gradient_hessian(tag, golden, cuda_models):
    """model(x) -> {'model_in','model_out'}; gradient/hessian with the reference's call pattern (src/evaluate.py:26-32)."""
    from diffudf_b200 import gradient, hessian
    J = golden(f"jets_{tag}.npz")
    m = cuda_models[tag]
    with torch.no_grad():
        pass
    x, y = m(torch.from_numpy(J["x"]).cuda().unsqueeze(0)).values()
    assert x.requires_grad and x.is_leaf and y.shape == (1, 512, 1)
    g = gradient(y, x)
    H = hessian(y, x)
    assert g.shape == (1, 512, 3) and H.shape == (1, 512, 3, 3)
    assert rel_max(y.detach().cpu().numpy()[0, :, 0], J["f32"]) < 1e-5
    assert rel_max(g.detach().cpu().numpy()[0], J["g32"]) < 1e-5
    assert rel_max(H.detach().cpu().numpy()[0], J["H32"]) < 1e-5


def test_unsupported_inputs_raise(cuda_models):
    from diffudf_b200 import SIREN, gradient
    with pytest.raises(ValueError):
        SIREN(3, 1, [128, 128])
    with pytest.raises(ValueError):
        SIREN(3, 1, [256], activation="relu")
    with pytest.raises(RuntimeError):
        cuda_models["init"](torch.zeros(4, 3))            # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        t = torch.zeros(4, 3, device="cuda", requires_grad=True)
        gradient(t.sum(-1, keepdim=True), t)              # not a SIREN forward: no autograd fallback
