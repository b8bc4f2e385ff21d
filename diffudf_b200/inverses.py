"""Drop-in for /root/reference/src/inverses.py:3-22 (host-side numpy; the device-side versions are fused
into the query kernels as the DUDF_Q_ABS_INV_TANH epilogue)."""
import numpy as np


def inv_squared(pred_df, alpha, min_step):
    out = np.ones_like(pred_df) * min_step
    np.sqrt(pred_df, out=out, where=pred_df > 0)
    out /= np.sqrt(alpha)
    return out


def inv_tanh(pred_df, alpha, min_step):
    with np.errstate(invalid="ignore"):
        return np.where(pred_df < 1 / alpha, np.sqrt(pred_df / alpha), pred_df)


def inv_siren(pred_df, alpha, min_step):
    return np.where(pred_df > 0, pred_df, np.ones_like(pred_df) * min_step)


def inverse(gt_mode, pred_df, alpha, min_step=0.01):
    table = {"siren": inv_siren, "squared": inv_squared, "tanh": inv_tanh}
    return table[gt_mode](pred_df, alpha, min_step)


def inverse_torch(gt_mode, f, alpha, min_step=0.01):
    """Same functions on torch tensors (device resident drivers), with the arithmetic numpy performs on an array of f's dtype:
    a true division by alpha (torch multiplies a tensor by the reciprocal of a Python scalar divisor, which differs in the last
    bit), the comparison against 1/alpha rounded to f's dtype."""
    import torch
    a = torch.tensor(alpha, dtype=f.dtype, device=f.device)
    if gt_mode == "tanh":
        return torch.where(f < 1 / alpha, torch.sqrt(f / a), f)
    if gt_mode == "siren":
        return torch.where(f > 0, f, torch.full_like(f, min_step))
    if gt_mode == "squared":
        r = torch.where(f > 0, torch.sqrt(f.clamp_min(0)), torch.full_like(f, min_step))
        return (r.to(torch.float64) / (alpha ** 0.5)).to(f.dtype)
    raise KeyError(gt_mode)
