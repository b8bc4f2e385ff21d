"""CPU oracle for the DUDF hot path — TEST INFRASTRUCTURE ONLY.

This file is a closed-form numpy restatement of the reference algorithm
(LIA-DiTella/DiffUDF).  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product path (``diffudf_b200``) never does and fails
loudly when the CUDA library is missing.

Parity status: PINNED.  The reference ships no golden vectors, so the
oracle is pinned against outputs of the unmodified reference modules imported from
``/root/reference`` in the build container (``tests/golden/make_golden.py`` -> fixtures
in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks them on every run).

What each function follows in the reference (paths relative to /root/reference):

* ``siren_jet``          src/model.py:116-135 (forward) + src/diff_operators.py:187-212
                         (gradient / hessian via autograd) restated as forward-mode jets.
* ``eig_top``            torch.linalg.eigh call sites: src/loss_functions.py:142-143,
                         src/render_st.py:59-60, src/render_mc.py:77-78.
* ``loss_s1/_s2/_siren`` src/loss_functions.py:123-155 / :106-121 / :82-104.
* ``train_grads``        train.py:195-222 (loss sum + backward) as an analytic reverse sweep.
* ``adam_step``          torch.optim.Adam defaults as constructed in train.py:334-337.
* ``lr_schedule``        train.py:174-191.
* ``inverse``            src/inverses.py:3-22.
* ``evaluate``           src/evaluate.py:5-37.
* ``grid_coords`` / ``extract_fields``   src/render_mc.py:20-101.
* ``propagate_rays``     src/render_st.py:136-161.
* ``normals_and_curvature``  src/render_st.py:42-62 (eigen-normal, mean / gaussian curvature).
* ``project_points``     src/render_pc.py:43-60.

All functions take ``params`` = list of (W, b) numpy arrays in nn.Linear layout
(W: [out, in]).  dtype of the computation follows ``dtype`` (float64 default).
"""
import numpy as np

SYM6 = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]


# --------------------------------------------------------------------------------------
# network container helpers
# --------------------------------------------------------------------------------------
def init_params(n_hidden=8, width=256, w0=30.0, seed=123):
    """SIREN init as src/model.py:7-19,94-113 using torch's RNG stream so that the
    weights are bit-identical to ``torch.manual_seed(seed); SIREN(3,1,[width]*n_hidden)``."""
    import torch
    torch.manual_seed(seed)
    dims = [3] + [width] * n_hidden + [1]
    lin = [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
    with torch.no_grad():
        lin[0].weight.uniform_(-1 / 3, 1 / 3)
        for l in lin[1:]:
            k = l.weight.size(-1)
            l.weight.uniform_(-np.sqrt(6 / k) / w0, np.sqrt(6 / k) / w0)
    return [(l.weight.detach().numpy().copy(), l.bias.detach().numpy().copy()) for l in lin]


def params_from_state_dict(sd):
    n = len([k for k in sd if k.endswith(".0.weight")])
    out = []
    for i in range(n):
        W = sd[f"net.{i}.0.weight"]
        b = sd[f"net.{i}.0.bias"]
        W = W.detach().cpu().numpy() if hasattr(W, "detach") else np.asarray(W)
        b = b.detach().cpu().numpy() if hasattr(b, "detach") else np.asarray(b)
        out.append((W, b))
    return out


def save_params(path, params):
    np.savez(path, **{f"W{i}": W for i, (W, _) in enumerate(params)},
             **{f"b{i}": b for i, (_, b) in enumerate(params)})


def load_params(path):
    z = np.load(path)
    n = len([k for k in z.files if k.startswith("W")])
    return [(z[f"W{i}"], z[f"b{i}"]) for i in range(n)]


# --------------------------------------------------------------------------------------
# forward-mode jets through the sine MLP
# --------------------------------------------------------------------------------------
def siren_jet(params, x, order=1, w0=30.0, dtype=np.float64, keep=False):
    """Value and input-space derivatives of f(x) up to ``order`` (0..3).

    Returns dict with 'f' (P,), 'g' (P,3), 'H' (P,3,3), 'T' (P,3,3,3) as available.
    With keep=True also returns the per-layer pre-activations (for the reverse sweep).
    """
    x = np.asarray(x, dtype=dtype).reshape(-1, 3)
    P = x.shape[0]
    w = dtype(w0)
    L = len(params) - 1                      # number of sine layers
    a = x                                     # (P, in)
    a1 = a2 = a3 = None
    stash = []
    for l in range(L):
        W = params[l][0].astype(dtype)
        b = params[l][1].astype(dtype)
        z = a @ W.T + b                       # (P, n)
        if order >= 1:
            if l == 0:
                z1 = np.broadcast_to(W.T[None, :, :], (P, 3, W.shape[0])).copy()   # (P,3,n): d_i z = W[:,i]
            else:
                z1 = a1 @ W.T
        if order >= 2:
            z2 = np.zeros((P, 3, 3, W.shape[0]), dtype) if l == 0 else a2 @ W.T
        if order >= 3:
            z3 = np.zeros((P, 3, 3, 3, W.shape[0]), dtype) if l == 0 else a3 @ W.T
        s = np.sin(w * z)
        c = np.cos(w * z)
        if keep:
            stash.append(dict(a=a, a1=a1, a2=a2, z=z, z1=z1 if order >= 1 else None,
                              z2=z2 if order >= 2 else None, s=s, c=c))
        a = s
        if order >= 1:
            a1 = w * c[:, None, :] * z1
        if order >= 2:
            a2 = (w * c[:, None, None, :] * z2
                  - w * w * s[:, None, None, :] * z1[:, :, None, :] * z1[:, None, :, :])
        if order >= 3:
            zi = z1[:, :, None, None, :]
            zj = z1[:, None, :, None, :]
            zk = z1[:, None, None, :, :]
            a3 = w * c[:, None, None, None, :] * z3
            a3 = a3 - w * w * s[:, None, None, None, :] * (
                z2[:, :, :, None, :] * zk          # z_ij z_k
                + z2[:, :, None, :, :] * zj        # z_ik z_j
                + z2[:, None, :, :, :] * zi)       # z_jk z_i
            a3 = a3 - w ** 3 * c[:, None, None, None, :] * zi * zj * zk
    W = params[L][0].astype(dtype)
    b = params[L][1].astype(dtype)
    out = {"f": (a @ W.T + b)[:, 0]}
    if order >= 1:
        out["g"] = (a1 @ W.T)[..., 0]
    if order >= 2:
        out["H"] = (a2 @ W.T)[..., 0]
    if order >= 3:
        out["T"] = (a3 @ W.T)[..., 0]
    if keep:
        out["stash"] = stash
        out["a_last"] = (a, a1, a2)
    return out


# --------------------------------------------------------------------------------------
# small symmetric eigenproblems
# --------------------------------------------------------------------------------------
def eig_top(H):
    """Eigen-decomposition of symmetric 3x3 matrices, ascending (as torch.linalg.eigh).
    Uses the lower triangle like LAPACK 'L'.  Returns (lam (P,3), V (P,3,3)); V[...,k] is
    the k-th eigenvector; the sign is the LAPACK one and must be treated as arbitrary."""
    H = np.asarray(H)
    Hs = np.tril(H) + np.swapaxes(np.tril(H, -1), -1, -2)
    lam, V = np.linalg.eigh(Hs)
    return lam, V


# --------------------------------------------------------------------------------------
# inverse scalings (src/inverses.py)
# --------------------------------------------------------------------------------------
def inv_tanh(f, alpha, min_step=0.01):
    f = np.asarray(f)
    with np.errstate(invalid="ignore"):
        return np.where(f < 1.0 / alpha, np.sqrt(f / alpha), f)


def inv_siren(f, alpha=None, min_step=0.01):
    f = np.asarray(f)
    return np.where(f > 0, f, np.ones_like(f) * min_step)


def inv_squared(f, alpha, min_step=0.01):
    f = np.asarray(f, dtype=np.float64)
    inv = np.ones_like(f) * min_step
    np.sqrt(f, out=inv, where=f > 0)
    return inv / np.sqrt(alpha)


def inverse(gt_mode, f, alpha, min_step=0.01):
    return {"tanh": inv_tanh, "siren": inv_siren, "squared": inv_squared}[gt_mode](f, alpha, min_step)


# --------------------------------------------------------------------------------------
# losses (values) — src/loss_functions.py
# --------------------------------------------------------------------------------------
def _cos_sim(a, b, eps=1e-8):
    # torch.nn.functional.cosine_similarity: x.y / (max(|x|,eps) * max(|y|,eps))
    na = np.maximum(np.linalg.norm(a, axis=-1), eps)
    nb = np.maximum(np.linalg.norm(b, axis=-1), eps)
    return np.sum(a * b, axis=-1) / (na * nb)


def loss_s1(params, x, normals, d, weights, alpha, w0=30.0, dtype=np.float64):
    x = np.asarray(x, dtype).reshape(-1, 3)
    n_gt = np.asarray(normals, dtype).reshape(-1, 3)
    d = np.asarray(d, dtype).reshape(-1)
    P = x.shape[0]
    order = 2 if weights[2] != 0 else (1 if weights[3] != 0 else 0)
    j = siren_jet(params, x, order, w0, dtype)
    f = j["f"]
    t = np.tanh(alpha * d)
    on = d == 0
    out = {
        "sdf_on_surf": np.sum(np.abs(f) * on) / P * weights[0],
        "sdf_off_surf": np.sum(np.abs(d * t - f) * (~on)) / P * weights[1],
    }
    if weights[2] != 0:
        _, V = eig_top(j["H"])
        v = V[..., 2]
        out["hessian_constraint"] = np.sum((1 - np.abs(_cos_sim(n_gt, v))) * on) / P * weights[2]
    else:
        out["hessian_constraint"] = 0.0
    if weights[3] != 0:
        tgt = np.abs(t + d * alpha * (1 - t * t))
        out["grad_constraint"] = np.sum(np.abs(np.linalg.norm(j["g"], axis=-1) - tgt)) / P * weights[3]
    else:
        out["grad_constraint"] = 0.0
    return out


def loss_s2(params, x, normals, d, weights, alpha=None, w0=30.0, dtype=np.float64):
    x = np.asarray(x, dtype).reshape(-1, 3)
    d = np.asarray(d, dtype).reshape(-1)
    f = siren_jet(params, x, 0, w0, dtype)["f"]
    s = f[d == 0]
    return {"sdf_on_surf": np.abs(np.mean(s)) * weights[0],
            "std_on_surf": np.std(s, ddof=1) * weights[1]}


def loss_siren(params, x, normals, d, weights, w0=30.0, dtype=np.float64):
    x = np.asarray(x, dtype).reshape(-1, 3)
    n_gt = np.asarray(normals, dtype).reshape(-1, 3)
    d = np.asarray(d, dtype).reshape(-1)
    P = x.shape[0]
    j = siren_jet(params, x, 1, w0, dtype)
    f, g = j["f"], j["g"]
    on = d == 0
    return {
        "sdf_on_surf": np.sum(np.abs(f) * on) / P * weights[0],
        "sdf_off_surf": np.sum(np.exp(-1e2 * np.abs(f)) * (~on)) / P * weights[1],
        "normal_constraint": np.sum((1 - _cos_sim(g, n_gt)) * on) / P * weights[2],
        "grad_constraint": np.sum((np.linalg.norm(g, axis=-1) - 1.0) ** 2) / P * weights[3],
    }


# --------------------------------------------------------------------------------------
# analytic adjoints: d(sum of loss terms)/d(params)
# --------------------------------------------------------------------------------------
def loss_seeds(mode, f, g, H, n_gt, d, weights, alpha, P_global=None, upstream=None, s2_stats=None):
    """Per-row adjoints (fbar (P,), gbar (P,3), Hbar (P,3,3)) of sum_k upstream[k]*term_k.
    Returns also the loss terms (as local sums already divided by P_global)."""
    P = f.shape[0]
    Pg = P if P_global is None else P_global
    on = d == 0
    dt = f.dtype
    fbar = np.zeros(P, dt)
    gbar = np.zeros((P, 3), dt) if g is not None else None
    Hbar = np.zeros((P, 3, 3), dt) if H is not None else None
    up = [1.0] * 4 if upstream is None else upstream
    terms = {}
    if mode == "s1":
        t = np.tanh(alpha * d)
        tdf = d * t
        terms["sdf_on_surf"] = np.sum(np.abs(f) * on) / Pg * weights[0]
        terms["sdf_off_surf"] = np.sum(np.abs(tdf - f) * (~on)) / Pg * weights[1]
        fbar += up[0] * weights[0] / Pg * np.sign(f) * on
        fbar += -up[1] * weights[1] / Pg * np.sign(tdf - f) * (~on)
        terms["hessian_constraint"] = 0.0
        terms["grad_constraint"] = 0.0
        if weights[2] != 0:
            lam, V = eig_top(H)
            v = V[..., 2]
            nn_ = np.maximum(np.linalg.norm(n_gt, axis=-1), 1e-8)
            vn = np.maximum(np.linalg.norm(v, axis=-1), 1e-8)
            cosv = np.sum(n_gt * v, -1) / (nn_ * vn)
            terms["hessian_constraint"] = np.sum((1 - np.abs(cosv)) * on) / Pg * weights[2]
            # d(-|cos|)/dv = -sign(cos) * (n/(|n||v|) - cos * v/|v|^2)
            coef = (-up[2] * weights[2] / Pg) * np.sign(cosv) * on
            nbar = coef[:, None] * (n_gt / (nn_ * vn)[:, None] - cosv[:, None] * v / (vn * vn)[:, None])
            for jdx in (0, 1):
                vj = V[..., jdx]
                gap = lam[:, 2] - lam[:, jdx]
                with np.errstate(divide="ignore", invalid="ignore"):
                    cj = np.where(on, np.sum(vj * nbar, -1) / gap, 0.0)
                Hbar += cj[:, None, None] * 0.5 * (vj[:, :, None] * v[:, None, :] + v[:, :, None] * vj[:, None, :])
        if weights[3] != 0:
            tgt = np.abs(t + d * alpha * (1 - t * t))
            gn = np.linalg.norm(g, axis=-1)
            terms["grad_constraint"] = np.sum(np.abs(gn - tgt)) / Pg * weights[3]
            with np.errstate(divide="ignore", invalid="ignore"):
                unit = np.where(gn[:, None] > 0, g / gn[:, None], 0.0)
            gbar += (up[3] * weights[3] / Pg) * np.sign(gn - tgt)[:, None] * unit
    elif mode == "siren":
        terms["sdf_on_surf"] = np.sum(np.abs(f) * on) / Pg * weights[0]
        e = np.exp(-1e2 * np.abs(f))
        terms["sdf_off_surf"] = np.sum(e * (~on)) / Pg * weights[1]
        fbar += up[0] * weights[0] / Pg * np.sign(f) * on
        fbar += up[1] * weights[1] / Pg * (-1e2) * np.sign(f) * e * (~on)
        gn_raw = np.linalg.norm(g, axis=-1)
        gn = np.maximum(gn_raw, 1e-8)
        nn_ = np.maximum(np.linalg.norm(n_gt, axis=-1), 1e-8)
        cosv = np.sum(g * n_gt, -1) / (gn * nn_)
        terms["normal_constraint"] = np.sum((1 - cosv) * on) / Pg * weights[2]
        terms["grad_constraint"] = np.sum((gn_raw - 1.0) ** 2) / Pg * weights[3]
        # d(1-cos)/dg (|g| > eps): -(n/(|g||n|) - cos * g/|g|^2)
        coef = (up[2] * weights[2] / Pg) * on
        gbar += -coef[:, None] * (n_gt / (gn * nn_)[:, None] - cosv[:, None] * g / (gn * gn)[:, None])
        with np.errstate(divide="ignore", invalid="ignore"):
            unit = np.where(gn_raw[:, None] > 0, g / gn_raw[:, None], 0.0)
        gbar += (up[3] * weights[3] / Pg) * 2.0 * (gn_raw - 1.0)[:, None] * unit
    elif mode == "s2":
        # s2_stats = (n, sum, sumsq) over ALL ranks' on-surface rows (global), default local
        s = f[on]
        if s2_stats is None:
            n, s1_, s2_ = s.size, np.sum(s), np.sum(s * s)
        else:
            n, s1_, s2_ = s2_stats
        mean = s1_ / n
        var = (s2_ - n * mean * mean) / (n - 1)
        std = np.sqrt(var)
        terms["sdf_on_surf"] = np.abs(mean) * weights[0]
        terms["std_on_surf"] = std * weights[1]
        fbar += on * (up[0] * weights[0] * np.sign(mean) / n)
        fbar += on * (up[1] * weights[1] * (f - mean) / ((n - 1) * std))
    else:
        raise ValueError(mode)
    return terms, fbar, gbar, Hbar


def reverse_sweep(params, x, jet, fbar, gbar=None, Hbar=None, w0=30.0):
    """Reverse sweep through the jet network: returns list of (Wbar, bbar).
    Follows SURVEY.md §8 a-M (verified there against autograd)."""
    dt = fbar.dtype
    w = dt.type(w0)
    L = len(params) - 1
    stash = jet["stash"]
    a, a1, a2 = jet["a_last"]
    P = fbar.shape[0]
    order = 2 if Hbar is not None else (1 if gbar is not None else 0)
    grads = [None] * (L + 1)
    W = params[L][0].astype(dt)               # (1, n)
    Wb = fbar[:, None] * a                    # rows: fbar * a
    Wbar = np.sum(Wb, 0)[None, :]
    if order >= 1:
        Wbar = Wbar + np.einsum("pi,pin->n", gbar, a1)[None, :]
    if order >= 2:
        Wbar = Wbar + np.einsum("pij,pijn->n", Hbar, a2)[None, :]
    grads[L] = (Wbar, np.array([np.sum(fbar)], dt))
    ab = fbar[:, None] * W                    # (P, n)
    ab1 = gbar[:, :, None] * W[None] if order >= 1 else None
    ab2 = Hbar[:, :, :, None] * W[None, None] if order >= 2 else None
    for l in range(L - 1, -1, -1):
        st = stash[l]
        s, c, z1, z2 = st["s"], st["c"], st["z1"], st["z2"]
        zb = w * c * ab
        if order >= 1:
            zb = zb - w * w * s * np.einsum("pin,pin->pn", ab1, z1)
            zb1 = w * c[:, None, :] * ab1
        if order >= 2:
            zb = zb - np.einsum("pijn,pijn->pn", ab2,
                                w * w * s[:, None, None, :] * z2
                                + w ** 3 * c[:, None, None, :] * z1[:, :, None, :] * z1[:, None, :, :])
            zb1 = zb1 - w * w * s[:, None, :] * np.einsum("pijn,pjn->pin", ab2 + ab2.transpose(0, 2, 1, 3), z1)
            zb2 = w * c[:, None, None, :] * ab2
        Wl = params[l][0].astype(dt)
        ain, ain1, ain2 = st["a"], st["a1"], st["a2"]
        Wbar = zb.T @ ain
        if l == 0:
            if order >= 1:
                Wbar = Wbar + np.sum(zb1, 0).T       # d_i x = e_i
        else:
            if order >= 1:
                Wbar = Wbar + np.einsum("pin,pik->nk", zb1, ain1)
            if order >= 2:
                Wbar = Wbar + np.einsum("pijn,pijk->nk", zb2, ain2)
        grads[l] = (Wbar, np.sum(zb, 0))
        if l > 0:
            ab = zb @ Wl
            if order >= 1:
                ab1 = zb1 @ Wl
            if order >= 2:
                ab2 = zb2 @ Wl
    return grads


def train_grads(params, x, normals, d, mode, weights, alpha, w0=30.0, dtype=np.float64,
                P_global=None, upstream=None, s2_stats=None, hess_all_rows=True):
    """Loss terms and d(sum terms)/d(params) for one batch (train.py:204-221)."""
    x = np.asarray(x, dtype).reshape(-1, 3)
    n_gt = np.asarray(normals, dtype).reshape(-1, 3)
    d = np.asarray(d, dtype).reshape(-1)
    if mode == "s1":
        order = 2 if weights[2] != 0 else (1 if weights[3] != 0 else 0)
    elif mode == "siren":
        order = 1
    else:
        order = 0
    jet = siren_jet(params, x, order, w0, dtype, keep=True)
    terms, fbar, gbar, Hbar = loss_seeds(mode, jet["f"], jet.get("g"), jet.get("H"), n_gt, d,
                                         weights, alpha, P_global, upstream, s2_stats)
    if order < 2:
        Hbar = None
    if order < 1:
        gbar = None
    grads = reverse_sweep(params, x, jet, fbar, gbar, Hbar, w0)
    return terms, grads


def adam_step(params, grads, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad); t is the 1-based step count.
    params/m/v are lists of (W,b) updated functionally; returns (params, m, v)."""
    np_, nm, nv = [], [], []
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    for (W, b), (gW, gb), (mW, mb), (vW, vb) in zip(params, grads, m, v):
        o = []
        for p, g, mm, vv in ((W, gW, mW, vW), (b, gb, mb, vb)):
            g = g.reshape(p.shape).astype(p.dtype)
            mm = b1 * mm + (1 - b1) * g
            vv = b2 * vv + (1 - b2) * g * g
            step = lr / bc1
            denom = np.sqrt(vv) / np.sqrt(bc2) + eps
            p = p - step * mm / denom
            o.append((p.astype(W.dtype), mm, vv))
        np_.append((o[0][0], o[1][0]))
        nm.append((o[0][1], o[1][1]))
        nv.append((o[0][2], o[1][2]))
    return np_, nm, nv


def lr_schedule(epoch, epochs, s1_epochs, warmup_epochs, warmup_lr, lr_s1, lr_s2):
    """Learning rate in effect during ``epoch`` (train.py:167-191)."""
    if epoch >= s1_epochs:
        return 0.5 * (np.cos(epoch / (epochs - s1_epochs) * np.pi) + 1) * lr_s2
    if epoch >= warmup_epochs:
        return lr_s1
    return warmup_lr


# --------------------------------------------------------------------------------------
# field queries
# --------------------------------------------------------------------------------------
def evaluate(params, samples, want_grad=False, want_hess=False, w0=30.0, dtype=np.float32):
    """src/evaluate.py:5-37: returns (f (N,1) f64, grads (N,3) f64 | None, hess (N,3,3) f64 | None)."""
    order = 2 if want_hess else (1 if want_grad else 0)
    j = siren_jet(params, samples, order, w0, dtype)
    f = j["f"].astype(np.float64)[:, None]
    g = j["g"].astype(np.float64) if want_grad else None
    H = j["H"].astype(np.float64) if want_hess else None
    return f, g, H


def grid_coords(N, dtype=np.float32):
    """src/render_mc.py:36-49: index (i0,i1,i2) with i2 fastest; coord = idx*voxel - 1 in fp32."""
    idx = np.arange(N ** 3, dtype=np.int64)
    vs = np.float32(2.0 / (N - 1))
    out = np.empty((N ** 3, 3), np.float32)
    out[:, 2] = (idx % N).astype(np.float32) * vs + np.float32(-1)
    out[:, 1] = ((idx // N) % N).astype(np.float32) * vs + np.float32(-1)
    out[:, 0] = ((idx // N // N) % N).astype(np.float32) * vs + np.float32(-1)
    return out.astype(dtype)


def _normalize_rows(v, eps=1e-12):
    # torch.nn.functional.normalize: v / max(|v|, eps)
    n = np.maximum(np.linalg.norm(v, axis=-1, keepdims=True), eps)
    return v / n


def extract_fields(params, N, gt_mode, alpha, w0=30.0, dtype=np.float32, chunk=None):
    """src/render_mc.py:20-101: returns df (N,N,N) fp32 and vecs (N,N,N,3) fp32.  chunk: evaluate that many grid points at a
    time (the reference itself evaluates in chunks of max_batch points, src/evaluate.py:21-35; points are independent)."""
    xs = grid_coords(N)
    total = xs.shape[0]
    chunk = total if chunk is None else int(chunk)
    df = np.empty(total, np.float32)
    vecs = np.empty((total, 3), np.float32)
    for a in range(0, total, chunk):
        f, g, H = evaluate(params, xs[a:a + chunk], True, True, w0, dtype)
        df[a:a + chunk] = inverse(gt_mode, np.abs(f), alpha)[:, 0]
        grads = -1.0 * _normalize_rows(g)
        lam, V = eig_top(H)
        n = V[..., 2]
        sgn = np.where(np.sum(grads * n, -1, keepdims=True) < 0, -1.0, 1.0)
        n = n * sgn
        gn = np.linalg.norm(grads, axis=-1, keepdims=True)
        vecs[a:a + chunk] = np.where(gn < 0.04, n, grads)
    return df.reshape(N, N, N), vecs.reshape(N, N, N, 3)


def propagate_rays(params, rays, t0, mask, gt_mode, alpha, surface_threshold, max_iterations,
                   w0=30.0, dtype=np.float32):
    """src/render_st.py:136-161.  t0 (R,3) f64 and mask (R,) bool are updated in place like
    the reference; returns hits (R,) bool and the number of field evaluations performed."""
    hits = np.zeros_like(mask, dtype=bool)
    it = 0
    nq = 0
    while np.sum(mask) > 0 and it < max_iterations:
        udfs = siren_jet(params, t0[mask].astype(np.float32), 0, w0, dtype)["f"].astype(np.float32)[:, None]
        nq += udfs.shape[0]
        steps = inverse(gt_mode, np.abs(udfs), alpha)
        t0[mask] += rays[mask] * steps
        if gt_mode == "siren":
            thr = udfs.flatten() < surface_threshold
        else:
            thr = np.abs(steps).flatten() < surface_threshold
        ind = np.logical_and(np.all(t0[mask] > -1, axis=1), np.all(t0[mask] < 1, axis=1))
        hits[mask] += np.logical_and(thr, ind)
        mask[mask] *= np.logical_and(np.logical_not(thr), ind)
        it += 1
    return hits, nq


def normals_and_curvature(params, x, w0=30.0, dtype=np.float64, kind="mean"):
    """src/render_st.py:42-62: n = top eigenvector of Hess f, principal directions V[:, :2],
    mean curvature = tr(dn/dx)/2, gaussian = -det [[dn/dx, n],[n^T, 0]] using
    dn/dx_k = sum_{j<2} v_j (v_j^T T[:,:,k] n)/(lam_2 - lam_j)   (SURVEY.md §8 a-M)."""
    j = siren_jet(params, x, 3, w0, dtype)
    lam, V = eig_top(j["H"])
    n = V[..., 2]
    T = j["T"]
    J = np.zeros((n.shape[0], 3, 3), dtype)          # J[p, i, k] = d n_i / d x_k
    for jdx in (0, 1):
        vj = V[..., jdx]
        c = np.einsum("pa,pabk,pb->pk", vj, T, n) / (lam[:, 2] - lam[:, jdx])[:, None]
        J += vj[:, :, None] * c[:, None, :]
    mean = 0.5 * np.trace(J, axis1=1, axis2=2)
    ext = np.zeros((n.shape[0], 4, 4), dtype)
    ext[:, :3, :3] = J
    ext[:, :3, 3] = n
    ext[:, 3, :3] = n
    gauss = -np.linalg.det(ext)
    return dict(n=n, dirs=V[..., :2], lam=lam, mean=mean, gauss=gauss, H=j["H"], J=J, f=j["f"], g=j["g"])


def project_points(params, samples, num_steps, gt_mode, alpha, w0=30.0, dtype=np.float32):
    """src/render_pc.py:43-53 (inner loop): returns final samples (f64), last steps, last
    gradients and the Hessians evaluated at the positions of the last step's first query."""
    samples = np.array(samples, dtype=np.float64)
    H = None
    for step in range(num_steps):
        if step == num_steps - 1:
            H = siren_jet(params, samples.astype(np.float32), 2, w0, dtype)["H"].astype(np.float64)
        j = siren_jet(params, samples.astype(np.float32), 1, w0, dtype)
        udfs = j["f"].astype(np.float64)[:, None]
        g = j["g"].astype(np.float64)
        steps = inverse(gt_mode, udfs, alpha, min_step=0)
        gn = np.linalg.norm(g, axis=1, keepdims=True)
        samples = samples - steps * (g / gn)
    return samples, steps, g, H


# ---------------------------------------------------------------------------------------------
# Point-cloud batch sampler: src/dataset.py:72-131 with the random draws as inputs
# ---------------------------------------------------------------------------------------------
def shortest_distance(P, X):
    """Distance from each row of P to its nearest row of X, with the reference's expansion
    sqrt(min_x(|x|^2 - 2 p.x) + |p|^2) (src/dataset.py:72-78)."""
    sqP = np.sum(P * P, axis=1)
    sqX = np.sum(X * X, axis=1)
    m = np.min(sqX[None, :] - 2.0 * (P @ X.T), axis=1)
    return np.sqrt(m + sqP)


def sample_training_data_pc(surf_pts, surf_nrm, n_on, n_off, on_idx, far, near_idx, near_off):
    """One [on | far | near] batch from an oriented point cloud (src/dataset.py:80-131).  Draws: on_idx (n_on,) indices
    into the cloud, far (n_off // 2, 3) uniform domain points, near_idx (n_near,) indices into the ON rows, near_off
    (n_near, 1) normal offsets.  Returns coords (1, P, 3), normals (1, P, 3), sdf (1, P, 1), float32."""
    X = np.asarray(surf_pts, np.float64)
    Nn = np.asarray(surf_nrm, np.float64)
    on_p, on_n = X[on_idx], Nn[on_idx]
    far = np.asarray(far, np.float64)
    far_d = shortest_distance(far, X)
    off = np.asarray(near_off, np.float64).reshape(-1, 1)
    near_p = on_p[near_idx] + on_n[near_idx] * off
    coords = np.vstack([on_p, far, near_p])
    normals = np.vstack([on_n, np.zeros((n_off, 3))])
    sdf = np.concatenate([np.zeros(n_on), far_d, np.abs(off[:, 0])])[:, None]
    return coords.astype(np.float32)[None], normals.astype(np.float32)[None], sdf.astype(np.float32)[None]
