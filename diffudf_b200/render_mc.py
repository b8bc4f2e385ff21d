"""Field extraction for marching cubes: drop-in for extract_fields / get_mesh_sdf's grid evaluation
(/root/reference/src/render_mc.py:20-101, :314-358).  The grid coordinates are generated inside the
query kernel (no N^3 x 7 sample tensor, no host float64 temporaries); the meshing itself (MeshUDF /
CAP marching cubes) is a downstream consumer and out of scope."""
import torch

from .engine import Q_ABS_INV_TANH, Q_NEG_NORMALIZE
from .inverses import inverse_torch


def extract_fields(decoder, latent_vec, N, gt_mode, device, alpha, first=0, count=None):
    """Returns df (N,N,N) fp32 = inverse(gt_mode, |f|) and vecs (N,N,N,3) fp32 = -normalize(grad f) on `device`.
    Where grad f is exactly zero the sign-aligned top Hessian eigenvector is used (render_mc.py:77-93).
    `first`/`count` restrict the evaluation to a range of the flat index (slab sharding); the
    returned tensors are then flat: df (count,), vecs (count,3)."""
    if latent_vec is not None and torch.is_tensor(latent_vec) and latent_vec.numel() != 0:
        raise ValueError("extract_fields: latent conditioning is not supported")
    eng = decoder._engine_synced()
    total = N ** 3
    full = count is None
    count = total - first if count is None else count
    flags = Q_NEG_NORMALIZE | (Q_ABS_INV_TANH if gt_mode == "tanh" else 0)
    df, vecs, _ = eng.query_grid(N, first, count, decoder.precision, flags, alpha, want_vecs=True)
    if gt_mode != "tanh":
        df = inverse_torch(gt_mode, df.abs(), alpha)
    # zero-gradient fallback (rare): Hessian eigenvector, aligned with the (zero) gradient like the reference
    small = torch.linalg.norm(vecs, dim=-1) < 0.04
    if bool(small.any()):
        idx = torch.nonzero(small).reshape(-1)
        from .render_st import grid_points
        pts = grid_points(N, idx + first, vecs.device)
        _, _, H, _ = eng.query(pts, 2, "fp32")
        vecs[idx] = eng.field_vectors(vecs[idx].contiguous(), H)
    if full and first == 0:
        return df.reshape(N, N, N), vecs.reshape(N, N, N, 3)
    return df, vecs


def grid_values(decoder, N, device=None, first=0, count=None):
    """Raw field values on the grid (the evaluation loop of get_mesh_sdf, render_mc.py:331-345)."""
    eng = decoder._engine_synced()
    count = N ** 3 - first if count is None else count
    f, _, _ = eng.query_grid(N, first, count, decoder.precision, 0, 0.0, want_vecs=False)
    return f.reshape(N, N, N) if (first == 0 and count == N ** 3) else f
