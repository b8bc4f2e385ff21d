"""Drop-in for /root/reference/src/diff_operators.py (gradient :208-212, hessian :187-193,
jacobian :214-227, laplace :196-198, divergence :201-205).

The reference differentiates through an autograd graph over the coordinates.  Here the derivatives
are forward-mode jets computed by the fused kernels, so `x` must be the `model_in` tensor returned
by `diffudf_b200.SIREN.forward` (it carries the field record) and `y` its `model_out`.  Anything
else raises: there is no autograd fallback.
"""
import torch


def _record(x, who):
    rec = getattr(x, "_dudf", None)
    if rec is None:
        raise RuntimeError(f"{who}: `x` must be the 'model_in' tensor of a diffudf_b200.SIREN forward "
                           "(no autograd fallback is provided)")
    return rec


def gradient(y, x, grad_outputs=None):
    """d(sum y)/dx, shaped like x.  Differentiable w.r.t. the network parameters."""
    rec = _record(x, "gradient")
    _, g, _ = rec.jets(1)
    if grad_outputs is not None:
        go = grad_outputs.reshape(*rec.lead, 1).to(g.dtype)
        g = g * go
    return g


def hessian(y, x):
    """Hessian of the scalar field, shape (1, P, 3, 3) like the reference (which always unsqueezes)."""
    rec = _record(x, "hessian")
    _, _, H = rec.jets(2)
    return H.reshape(1, -1, 3, 3)


def laplace(y, x):
    rec = _record(x, "laplace")
    _, _, H = rec.jets(2)
    return (H[..., 0, 0] + H[..., 1, 1] + H[..., 2, 2]).unsqueeze(-1)


def divergence(y, x):
    """Only the divergence of the field gradient (= laplace) can be served by the jets."""
    rec = _record(x, "divergence")
    g = rec.cache.get(1, rec.cache.get(2, (None, None, None)))[1]
    if g is None or y.data_ptr() != g.data_ptr():
        raise RuntimeError("divergence: only y = gradient(model_out, model_in) is supported")
    return laplace(y, x)


def jacobian(y, x):
    """Jacobian of the eigen-normal field w.r.t. x: (jac (1, P, 3, 3), status).

    The reference's only caller passes y = top eigenvector of hessian(model_out, model_in)
    (src/render_st.py:42-62); the jacobian then needs third derivatives of the field.  They are
    computed by the order-3 kernel and contracted in closed form (SURVEY.md §8 a-M).  Rows of y that
    are not (+/-) that eigenvector make the call fail."""
    rec = _record(x, "jacobian")
    if y.shape[-1] != 3:
        raise RuntimeError("jacobian: only the eigen-normal field (last dim 3) is supported")
    H, T = rec.third()
    n, _, _, J = rec.model._engine.curvature(H, T)
    yv = y.detach().reshape(-1, 3).to(torch.float32)
    dots = (yv * n).sum(-1)
    if not bool((dots.abs() > 0.98).all()):
        raise RuntimeError("jacobian: y is not the top eigenvector of the field Hessian at x (unsupported graph)")
    J = J * torch.sign(dots)[:, None, None]
    jac = J.reshape(1, -1, 3, 3)
    status = -1 if bool(torch.isnan(jac).any()) else 0
    return jac, status
