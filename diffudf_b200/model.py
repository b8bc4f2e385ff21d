"""Drop-in for the reference network module (/root/reference/src/model.py:85-135).

Same constructor, same state_dict keys (`net.{i}.0.weight|bias`, fp32), same forward contract
(`{"model_in": coords, "model_out": f}` with `model_in` a fresh leaf) — but the field and its
input-space derivatives come from the hand-written sm_100a kernels through the C ABI
(include/dudf_b200.h) instead of nn.Linear + torch.sin + autograd double-backward.

The derivatives are forward-mode jets, so `diff_operators.gradient / hessian / jacobian` do not walk
an autograd graph over the coordinates: `forward` attaches a `FieldRecord` to `model_in`, and the
operators read (and lazily extend) it.  Everything stays differentiable w.r.t. the parameters
through one custom autograd node per jet order (reverse sweep kernels).  Configurations the kernels
do not cover (width != 256, non-sine activation, CPU tensors, non-sm_100 devices) raise — there is
no fallback.
"""
import numpy as np
import torch
from torch import nn

from .engine import Engine, NCH, SYM6


class SineLayer(nn.Module):
    """Parameter-free marker of the sine non-linearity (kept so that module trees / reprs match
    the reference, src/model.py:22-33); the math runs inside the fused kernels."""

    def __init__(self, w0=30):
        super().__init__()
        self.w0 = w0

    def forward(self, x):
        raise RuntimeError("SineLayer is evaluated inside the fused SIREN kernels; call the SIREN module")

    def __repr__(self):
        return f"SineLayer(w0={self.w0})"


@torch.no_grad()
def sine_init(m, w0):
    # src/model.py:7-12
    if hasattr(m, "weight"):
        k = m.weight.size(-1)
        m.weight.uniform_(-np.sqrt(6 / k) / w0, np.sqrt(6 / k) / w0)


@torch.no_grad()
def first_layer_sine_init(m):
    # src/model.py:15-19
    if hasattr(m, "weight"):
        k = m.weight.size(-1)
        m.weight.uniform_(-1 / k, 1 / k)


def _unpack(packed, order, lead):
    """packed (P, NCH) -> f lead+(1,), g lead+(3,), H lead+(3,3)."""
    f = packed[:, 0].reshape(*lead, 1)
    g = packed[:, 1:4].reshape(*lead, 3) if order >= 1 else None
    H = None
    if order >= 2:
        idx = torch.tensor(SYM6, device=packed.device).reshape(-1)
        H = packed[:, idx].reshape(*lead, 3, 3)
    return f, g, H


class _JetFn(torch.autograd.Function):
    """(f, g, H) = jets(x; params), differentiable w.r.t. the parameters (not the coordinates)."""

    @staticmethod
    def forward(ctx, model, x, order, *params):
        eng = model._engine_synced()
        P = x.shape[0]
        nch = NCH[order]
        ld = eng.stash_columns(order, P)
        ld = (ld + 3) // 4 * 4
        L = model.n_hidden
        Z = torch.empty(L, 256, ld, device=x.device, dtype=torch.float32)
        A = torch.empty(L, 256, ld, device=x.device, dtype=torch.float32)
        packed = torch.empty(P, nch, device=x.device, dtype=torch.float32)
        eng.jet_forward(x, order, packed, Z, A, ld, 0, "fp32")
        ctx.model, ctx.order, ctx.ld = model, order, ld
        ctx.save_for_backward(x, Z, A)
        ctx.sig = eng._sig
        outs = [packed[:, 0].clone()]
        outs.append(packed[:, 1:4].clone() if order >= 1 else x.new_zeros(0))
        if order >= 2:
            idx = torch.tensor(SYM6, device=x.device).reshape(-1)
            outs.append(packed[:, idx].reshape(P, 3, 3))
        else:
            outs.append(x.new_zeros(0))
        return tuple(outs)

    @staticmethod
    def backward(ctx, fb, gb, Hb):
        model, order, ld = ctx.model, ctx.order, ctx.ld
        x, Z, A = ctx.saved_tensors
        eng = model._engine_synced(1, reuse=True)
        if eng._sig != ctx.sig:
            raise RuntimeError("SIREN parameters changed between forward and backward")
        P = x.shape[0]
        nch = NCH[order]
        seeds = torch.zeros(P, nch, device=x.device, dtype=torch.float32)
        if fb is not None:
            seeds[:, 0] = fb
        if order >= 1 and gb is not None and gb.numel():
            seeds[:, 1:4] = gb
        if order >= 2 and Hb is not None and Hb.numel():
            Hs = Hb + Hb.transpose(1, 2)
            seeds[:, 4] = Hb[:, 0, 0]
            seeds[:, 5] = Hs[:, 0, 1]
            seeds[:, 6] = Hs[:, 0, 2]
            seeds[:, 7] = Hb[:, 1, 1]
            seeds[:, 8] = Hs[:, 1, 2]
            seeds[:, 9] = Hb[:, 2, 2]
        Zb = torch.empty_like(Z)
        ws, bs = model._weights_biases()
        gW = [torch.zeros_like(w) for w in ws]
        gB = [torch.zeros_like(b) for b in bs]
        eng.jet_backward(x, order, seeds, Z, Zb, ld, 0, gW, gB, "fp32")
        eng.jet_wgrad(Zb, A, ld, eng.stash_columns(order, P), gW, "fp32")
        grads = []
        for w, b in zip(gW, gB):
            grads += [w, b]
        return (None, None, None, *grads)


class FieldRecord:
    """What one SIREN.forward call knows about its field: attached to `model_in` as `_dudf`."""

    def __init__(self, model, coords):
        self.model = model
        self.lead = tuple(coords.shape[:-1])          # (the record does not keep `coords`: it is attached TO coords, and a
                                                      #  back-reference would leave the stashes to the cyclic collector)
        self.x = coords.detach().reshape(-1, 3).to(torch.float32).contiguous()
        self.cache = {}
        self.train = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())

    def jets(self, order):
        """(f lead+(1,), g lead+(3,), H lead+(3,3)) of at least `order`."""
        for o in sorted(self.cache):
            if o >= order:
                return self.cache[o]
        m = self.model
        if self.train:
            params = m._flat_params()
            f, g, H = _JetFn.apply(m, self.x, order, *params)
            res = (f.reshape(*self.lead, 1), g.reshape(*self.lead, 3) if order >= 1 else None,
                   H.reshape(*self.lead, 3, 3) if order >= 2 else None)
        else:
            eng = m._engine_synced()
            f, g, H, _ = eng.query(self.x, order, m.precision)
            res = (f.reshape(*self.lead, 1), g.reshape(*self.lead, 3) if g is not None else None,
                   H.reshape(*self.lead, 3, 3) if H is not None else None)
        self.cache[order] = res
        return res

    def third(self):
        """(H (P,3,3), T (P,10)) from the order-3 query (fp32 path)."""
        eng = self.model._engine_synced()
        _, _, H, T = eng.query(self.x, 3, "fp32")
        return H, T


class SIREN(nn.Module):
    """SIREN(n_in_features, n_out_features, hidden_layer_config, w0=30, ww=None, delay_init=False,
    activation='sine') — see /root/reference/src/model.py:48-113 for the parameter docs.

    Extras (not in the reference): `precision` ('fp32' CUDA-core path, 'tcx3' split-precision tcgen05 path with
    fp32-grade results, 'tc16' single-pass tcgen05 path with 1e-3-class results; used for queries) and `jet_order` (a hint: evaluate this derivative order already in forward so that a
    following gradient()/hessian() costs nothing more)."""

    def __init__(self, n_in_features, n_out_features, hidden_layer_config=[], w0=30, ww=None, delay_init=False,
                 activation="sine"):
        super().__init__()
        if activation != "sine":
            raise ValueError("diffudf_b200.SIREN: only activation='sine' is implemented (no fallback)")
        if n_in_features != 3 or n_out_features != 1:
            raise ValueError("diffudf_b200.SIREN: kernels are specialised for 3 -> 1 fields")
        if len(hidden_layer_config) < 1 or any(int(h) != 256 for h in hidden_layer_config):
            raise ValueError("diffudf_b200.SIREN: hidden layers must all have width 256 "
                             f"(got {list(hidden_layer_config)}); no fallback path exists")
        if len(hidden_layer_config) > 15:
            raise ValueError("diffudf_b200.SIREN: at most 15 hidden layers")
        self.w0 = w0
        self.ww = w0 if ww is None else ww
        net = [nn.Sequential(nn.Linear(n_in_features, hidden_layer_config[0]), SineLayer(self.w0))]
        for i in range(1, len(hidden_layer_config)):
            net.append(nn.Sequential(nn.Linear(hidden_layer_config[i - 1], hidden_layer_config[i]), SineLayer(self.ww)))
        net.append(nn.Sequential(nn.Linear(hidden_layer_config[-1], n_out_features)))
        self.net = nn.Sequential(*net)
        if not delay_init:
            self.net[0].apply(first_layer_sine_init)
            self.net[1:].apply(lambda module: sine_init(module, self.ww))
        self.n_hidden = len(hidden_layer_config)
        self.precision = "fp32"           # arithmetic of field queries: 'fp32' | 'tcx3' (split tensor-core, fp32-grade) | 'tc16'
        self.train_precision = "fp32"     # arithmetic of the fused losses / trainer (same choices)
        self.jet_order = 0
        self._engine = None

    # ---- plumbing ----
    def _weights_biases(self):
        ws = [self.net[i][0].weight for i in range(self.n_hidden + 1)]
        bs = [self.net[i][0].bias for i in range(self.n_hidden + 1)]
        return ws, bs

    def _flat_params(self):
        out = []
        for i in range(self.n_hidden + 1):
            out += [self.net[i][0].weight, self.net[i][0].bias]
        return out

    def _engine_synced(self, need=7, reuse=False):
        """The native engine with the weight images `need`ed up to date (bit 1: fp32 path, 2: single-pass tensor-core
        images, 4: split tensor-core images; engine.NEED maps a precision to its bits).  reuse=True keeps images that were
        built for the same parameter signature (see Engine.sync_weights)."""
        ws, bs = self._weights_biases()
        dev = ws[0].device
        if dev.type != "cuda":
            raise RuntimeError("diffudf_b200.SIREN has no CPU path: move the module to a CUDA (sm_100) device")
        if self._engine is None or self._engine.device != dev:
            self._engine = Engine(self.n_hidden, self.w0, self.ww, dev)
        self._engine.sync_weights(ws, bs, need, reuse)
        return self._engine

    def invalidate_weights(self):
        """Call after writing parameters behind torch's back (`p.data.copy_()` and friends) while a loss forward is
        pending; queries and new forwards re-read the parameters on their own."""
        if self._engine is not None:
            self._engine.invalidate()

    def forward(self, x):
        """x: (..., 3) coordinates.  Returns {'model_in': leaf copy of x, 'model_out': f(x) (..., 1)}
        in this key order (callers unpack `.values()`, src/evaluate.py:26)."""
        coords_org = x.clone().detach().requires_grad_(True)
        if coords_org.shape[-1] != 3:
            raise ValueError("SIREN.forward expects coordinates of shape (..., 3)")
        if not coords_org.is_cuda:
            raise RuntimeError("diffudf_b200.SIREN has no CPU path: pass CUDA tensors")
        rec = FieldRecord(self, coords_org)
        y = rec.jets(self.jet_order)[0]
        coords_org._dudf = rec
        return {"model_in": coords_org, "model_out": y}
