"""CPU emulation of the tensor-core operand roundings: which products of the training step need a split (hi + lo fp16)
operand for the parameter gradients / jets to reach north_star's tolerance?  Runs the oracle's closed-form jet and reverse
sweep with the hidden-layer contractions replaced by rounded-operand products (fp64 accumulation; the kernels accumulate
in fp32, 1e-7-class).  Test infrastructure: imports oracle/, never used by the product path.

  python tools/precision_study.py            # table for loss_s1 at the SIREN init and on the trained weights
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dudf_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def h16(a):
    return np.asarray(a, np.float64).astype(np.float16).astype(np.float64)


def split16(a, scale=1.0):
    hi = h16(a * scale)
    lo = h16(a * scale - hi)
    return hi / scale, lo / scale


def mm(A, B, mode):
    """A @ B with operand rounding.  mode: 'x' exact, '1' fp16 x fp16, '3' (Ah + Al) Bh + Ah Bl, 'a' only A split, 'b' only B split"""
    if mode == "x":
        return A @ B
    # per-tensor power-of-two scaling keeps the lo parts out of the subnormal range (the kernels scale by powers of two)
    sa = 2.0 ** np.floor(np.log2(1024.0 / max(np.abs(A).max(), 1e-300)))
    sb = 2.0 ** np.floor(np.log2(1024.0 / max(np.abs(B).max(), 1e-300)))
    Ah, Al = split16(A, sa)
    Bh, Bl = split16(B, sb)
    if mode == "1":
        return Ah @ Bh
    if mode == "3":
        return Ah @ Bh + Al @ Bh + Ah @ Bl
    if mode == "a":
        return Ah @ Bh + Al @ Bh
    if mode == "b":
        return Ah @ Bh + Ah @ Bl
    raise ValueError(mode)


def jet_rounded(params, x, order, fwd, stash16, out16, w0=30.0):
    """O.siren_jet with rounded hidden-layer products (fwd mode), optional fp16 rounding of the stashed derivative channels
    (what the reverse sweep reads) and of the last activations (output layer reads the fp16 tile)."""
    x = np.asarray(x, np.float64).reshape(-1, 3)
    P = x.shape[0]
    L = len(params) - 1
    a, a1, a2 = x, None, None
    stash = []
    for l in range(L):
        W = params[l][0].astype(np.float64)
        b = params[l][1].astype(np.float64)
        if l == 0:
            z = a @ W.T + b
            z1 = np.broadcast_to(W.T[None], (P, 3, W.shape[0])).copy() if order >= 1 else None
            z2 = np.zeros((P, 3, 3, W.shape[0])) if order >= 2 else None
        else:
            Wt = (w0 * W).T                                   # the kernels fold omega into the packed weights
            z = mm(a, Wt, fwd) / w0 + b
            z1 = (mm(a1.reshape(P * 3, -1), Wt, fwd) / w0).reshape(P, 3, -1) if order >= 1 else None
            z2 = (mm(a2.reshape(P * 9, -1), Wt, fwd) / w0).reshape(P, 3, 3, -1) if order >= 2 else None
        s, c = np.sin(w0 * z), np.cos(w0 * z)
        z1s = h16(z1) if (stash16 and z1 is not None and l > 0) else z1
        z2s = h16(z2) if (stash16 and z2 is not None and l > 0) else z2
        stash.append(dict(a=a, a1=a1, a2=a2, z=z, z1=z1s, z2=z2s, s=s, c=c))
        a = s
        if order >= 1:
            a1 = w0 * c[:, None, :] * z1
        if order >= 2:
            a2 = w0 * c[:, None, None, :] * z2 - w0 * w0 * s[:, None, None, :] * z1[:, :, None, :] * z1[:, None, :, :]
    W = params[L][0].astype(np.float64)
    b = params[L][1].astype(np.float64)
    r = h16 if out16 else (lambda t: t)
    out = {"f": (r(a) @ W.T + b)[:, 0]}
    if order >= 1:
        out["g"] = (r(a1) @ W.T)[..., 0]
    if order >= 2:
        out["H"] = (r(a2) @ W.T)[..., 0]
    out["stash"] = stash
    out["a_last"] = (a, a1, a2)
    return out


def reverse_rounded(params, jet, fbar, gbar, Hbar, dgrad, wgrad, w0=30.0):
    """O.reverse_sweep with rounded dgrad / wgrad products."""
    L = len(params) - 1
    stash = jet["stash"]
    a, a1, a2 = jet["a_last"]
    P = fbar.shape[0]
    order = 2 if Hbar is not None else (1 if gbar is not None else 0)
    grads = [None] * (L + 1)
    W = params[L][0].astype(np.float64)
    Wbar = np.sum(fbar[:, None] * a, 0)[None, :]
    if order >= 1:
        Wbar = Wbar + np.einsum("pi,pin->n", gbar, a1)[None, :]
    if order >= 2:
        Wbar = Wbar + np.einsum("pij,pijn->n", Hbar, a2)[None, :]
    grads[L] = (Wbar, np.array([np.sum(fbar)]))
    ab = fbar[:, None] * W
    ab1 = gbar[:, :, None] * W[None] if order >= 1 else None
    ab2 = Hbar[:, :, :, None] * W[None, None] if order >= 2 else None
    w = w0
    for l in range(L - 1, -1, -1):
        st = stash[l]
        s, c, z1, z2 = st["s"], st["c"], st["z1"], st["z2"]
        zb = w * c * ab
        if order >= 1:
            zb = zb - w * w * s * np.einsum("pin,pin->pn", ab1, z1)
            zb1 = w * c[:, None, :] * ab1
        if order >= 2:
            zb = zb - np.einsum("pijn,pijn->pn", ab2, w * w * s[:, None, None, :] * z2 + w ** 3 * c[:, None, None, :] * z1[:, :, None, :] * z1[:, None, :, :])
            zb1 = zb1 - w * w * s[:, None, :] * np.einsum("pijn,pjn->pin", ab2 + ab2.transpose(0, 2, 1, 3), z1)
            zb2 = w * c[:, None, None, :] * ab2
        Wl = params[l][0].astype(np.float64)
        ain, ain1, ain2 = st["a"], st["a1"], st["a2"]
        if l == 0:
            Wbar = zb.T @ ain
            if order >= 1:
                Wbar = Wbar + np.sum(zb1, 0).T
        else:
            # all channels of all points are the reduction dimension of ONE GEMM
            Z = [zb]
            A = [ain]
            if order >= 1:
                Z.append(zb1.reshape(P * 3, -1)); A.append(ain1.reshape(P * 3, -1))
            if order >= 2:
                Z.append(zb2.reshape(P * 9, -1)); A.append(ain2.reshape(P * 9, -1))
            Z = np.concatenate(Z, 0)
            A = np.concatenate(A, 0)
            Wbar = mm(Z.T, A, wgrad)
        grads[l] = (Wbar, np.sum(zb, 0))
        if l > 0:
            ab = mm(zb, Wl, dgrad)
            if order >= 1:
                ab1 = mm(zb1.reshape(P * 3, -1), Wl, dgrad).reshape(P, 3, -1)
            if order >= 2:
                ab2 = mm(zb2.reshape(P * 9, -1), Wl, dgrad).reshape(P, 3, 3, -1)
    return grads


def rel_max(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def study(tag, mode="s1", w=(1e4, 1e4, 1e4, 1e3)):
    params = O.load_params(os.path.join(GOLDEN, f"weights_{tag}.npz"))
    Ld = np.load(os.path.join(GOLDEN, f"losses_{tag}.npz"))
    x, n, d = Ld["x"].reshape(-1, 3).astype(np.float64), Ld["normals"].reshape(-1, 3).astype(np.float64), Ld["d"].reshape(-1).astype(np.float64)
    order = 2 if (mode == "s1" and w[2] != 0) else (0 if mode == "s2" else 1)
    ref = jet_rounded(params, x, order, "x", False, False)
    terms_ref, grads_ref = O.train_grads(params, x, n, d, mode, list(w), 100.0)
    print(f"== {tag} {mode} w={w}")
    rows = [("all single fp16 (r1 tc16)", "1", True, True, "1", "1"),
            ("fwd split, rest single", "3", True, True, "1", "1"),
            ("fwd split + fp32 stash/out", "3", False, False, "1", "1"),
            ("fwd+dgrad split, wgrad single", "3", False, False, "3", "1"),
            ("fwd+dgrad split, fp16 stash, wgrad single", "3", True, False, "3", "1"),
            ("fwd single, dgrad+wgrad split", "1", False, False, "3", "3"),
            ("fwd+wgrad split, dgrad single", "3", False, False, "1", "3"),
            ("all split", "3", False, False, "3", "3"),
            ("all split, fp16 stash", "3", True, False, "3", "3"),
            ("all split, fp16 stash+out", "3", True, True, "3", "3")]
    for name, fwd, st16, o16, dg, wg in rows:
        jet = jet_rounded(params, x, order, fwd, st16, o16)
        ef = rel_max(jet["f"], ref["f"])
        eg = rel_max(jet["g"], ref["g"]) if order >= 1 else 0.0
        eH = rel_max(jet["H"], ref["H"]) if order >= 2 else 0.0
        terms, fbar, gbar, Hbar = O.loss_seeds(mode, jet["f"], jet.get("g"), jet.get("H"), n, d, list(w), 100.0)
        if order < 2:
            Hbar = None
        if order < 1:
            gbar = None
        grads = reverse_rounded(params, jet, fbar, gbar, Hbar, dg, wg)
        eW = max(rel_max(g[0].reshape(r[0].shape), r[0]) for g, r in zip(grads, grads_ref))
        eb = max(rel_max(g[1].reshape(r[1].shape), r[1]) for g, r in zip(grads, grads_ref))
        et = max(abs(float(terms[k]) - float(terms_ref[k])) / max(abs(float(terms_ref[k])), 1e-2) for k in terms_ref)
        print(f"  {name:44s} f {ef:.1e} g {eg:.1e} H {eH:.1e} | terms {et:.1e} | gradW {eW:.1e} gradb {eb:.1e}")


if __name__ == "__main__":
    for tag in ("init", "trained"):
        study(tag, "s1", (1e4, 1e4, 1e4, 1e3))
        study(tag, "s1", (1e4, 1e4, 0, 1e3))
    study("init", "siren", (3e3, 1e2, 1e2, 5e1))
