"""Thin Python layer over the C ABI (include/dudf_b200.h): one `Engine` per SIREN module holds the
native context, keeps its weight images in sync with the torch parameters and exposes the query /
training primitives on torch CUDA tensors.  torch is used for device memory and streams only.
"""
import ctypes

import torch

from . import _lib

PRECISIONS = {"fp32": 0, "tc16": 1, "tcx3": 2}
TC_PRECISIONS = ("tc16", "tcx3")          # tensor-core arithmetics (share the stash / operand-image layouts)
# weight images a precision needs (dudf_refresh_weights bits): fp32 transposes 1, single-pass fp16 images 2 (the reverse
# sweep of both tensor-core modes), hi + lo fp16 images 4
NEED = {"fp32": 1, "tc16": 2, "tcx3": 6}
LOSS_MODES = {"s1": 0, "s2": 1, "siren": 2}
GT_MODES = {"tanh": 0, "siren": 1, "squared": 2}
Q_ABS_INV_TANH = 1
Q_NEG_NORMALIZE = 2
NCH = {0: 1, 1: 4, 2: 10, 3: 20}
# packed channel index of H[i][j]
SYM6 = [[4, 5, 6], [5, 7, 8], [6, 8, 9]]


def _f32c(t, device):
    """float32 contiguous CUDA view/copy of a tensor or numpy array."""
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.float32).contiguous()


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class Engine:
    def __init__(self, n_hidden, w0, ww, device):
        self.L = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("diffudf_b200 runs on CUDA devices only (no CPU fallback); move the model with .to('cuda')")
        self.n_hidden = n_hidden
        self.h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.L.dudf_create(n_hidden, float(w0), float(ww), ctypes.byref(self.h)), "dudf_create")
        self._sig = None
        self._fresh = 0
        self._bound = None

    def __del__(self):
        try:
            if self.h:
                self.L.dudf_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- weights ----
    def sync_weights(self, weights, biases, need=7, reuse=False):
        """Bind the parameter tensors (zero copy) and rebuild the derived operand images that `need` asks for
        (bit 1: fp32 transposes, bit 2: single-pass fp16 images, bit 4: hi + lo fp16 images).

        torch's version counters do not see writes through `.data` (EMA, clipping, manual re-initialisation), so the
        images are rebuilt on EVERY call (three small kernels, ~15 us of device time, no host synchronisation) unless the
        caller passes reuse=True — which is what the loss backward does after checking that the signature it recorded in
        its forward is unchanged.  `invalidate()` forces a rebuild explicitly."""
        sig = tuple((t.data_ptr(), t._version) for t in list(weights) + list(biases))
        if sig != self._sig or not reuse:
            self._fresh = 0
        if sig != self._sig:
            ws = [w.detach() for w in weights]
            bs = [b.detach() for b in biases]
            for t in ws + bs:
                if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                    raise RuntimeError("SIREN parameters must be contiguous float32 CUDA tensors")
            _lib.check(self.L.dudf_bind_weights(self.h, _ptr_array(ws), _ptr_array(bs)), "dudf_bind_weights")
            self._bound = (ws, bs)               # keep the storages alive while the library holds their pointers
            self._sig = sig
        todo = need & ~self._fresh
        if todo:
            with torch.cuda.device(self.device):
                _lib.check(self.L.dudf_refresh_weights(self.h, todo, _lib.current_stream()), "dudf_refresh_weights")
            self._fresh |= todo

    def invalidate(self):
        """Forget the derived weight images: the next use rebuilds them from the bound parameters."""
        self._sig = None
        self._fresh = 0

    # ---- queries ----
    def query(self, x, order, precision="fp32", flags=0, alpha=0.0):
        """x: (P,3) float32 CUDA.  Returns (f (P,), g (P,3)|None, H (P,3,3)|None, T (P,10)|None)."""
        P = x.shape[0]
        dev = x.device
        f = torch.empty(P, device=dev, dtype=torch.float32)
        g = torch.empty(P, 3, device=dev, dtype=torch.float32) if order >= 1 else None
        H = torch.empty(P, 3, 3, device=dev, dtype=torch.float32) if order >= 2 else None
        T = torch.empty(P, 10, device=dev, dtype=torch.float32) if order >= 3 else None
        with torch.cuda.device(dev):
            _lib.check(self.L.dudf_query_points(self.h, x.data_ptr(), P, order, flags, float(alpha), _lib.ptr(f), _lib.ptr(g),
                                                _lib.ptr(H), _lib.ptr(T), PRECISIONS[precision], _lib.current_stream()),
                       "dudf_query_points")
        return f, g, H, T

    def march_rays(self, pos, dirs, active, hit, gt_mode, alpha, thr, max_it, precision="fp32"):
        """Sphere tracing on the device (dudf_march_rays; src/render_st.py:136-172).  pos (R,3) float64 is advanced in place,
        active / hit (R,) uint8 are updated in place.  Returns the number of value queries."""
        if gt_mode not in GT_MODES:
            raise KeyError(gt_mode)
        for t, dt in ((pos, torch.float64), (dirs, torch.float64), (active, torch.uint8), (hit, torch.uint8)):
            if t.dtype != dt or not t.is_cuda or not t.is_contiguous():
                raise RuntimeError("march_rays: contiguous CUDA tensors (float64 positions / directions, uint8 masks) required")
        nq = ctypes.c_int64(0)
        with torch.cuda.device(pos.device):
            _lib.check(self.L.dudf_march_rays(self.h, pos.data_ptr(), dirs.data_ptr(), active.data_ptr(), hit.data_ptr(), pos.shape[0],
                                              GT_MODES[gt_mode], float(alpha), float(thr), int(max_it), PRECISIONS[precision],
                                              ctypes.byref(nq), _lib.current_stream()), "dudf_march_rays")
        return int(nq.value)

    def project_points(self, x, num_steps, gt_mode, alpha, precision="fp32", want_hess=True):
        """num_steps projection steps on the device (dudf_project_points; src/render_pc.py:43-53).  x (P,3) float64 is updated
        in place.  Returns (last steps (P,), last gradients (P,3), Hessians of the last step (P,3,3) | None)."""
        if gt_mode not in GT_MODES:
            raise KeyError(gt_mode)
        if x.dtype != torch.float64 or not x.is_cuda or not x.is_contiguous():
            raise RuntimeError("project_points: contiguous float64 CUDA tensor required")
        P = x.shape[0]
        steps = torch.empty(P, device=x.device, dtype=torch.float64)
        g = torch.empty(P, 3, device=x.device, dtype=torch.float32)
        H = torch.empty(P, 3, 3, device=x.device, dtype=torch.float32) if want_hess else None
        with torch.cuda.device(x.device):
            _lib.check(self.L.dudf_project_points(self.h, x.data_ptr(), P, int(num_steps), GT_MODES[gt_mode], float(alpha), steps.data_ptr(),
                                                  g.data_ptr(), _lib.ptr(H), PRECISIONS[precision], _lib.current_stream()),
                       "dudf_project_points")
        return steps, g, H

    def query_grid(self, N, first, count, precision="fp32", flags=0, alpha=0.0, want_vecs=True, want_hess=False, out=None):
        dev = self.device
        if out is None:
            df = torch.empty(count, device=dev, dtype=torch.float32)
            vecs = torch.empty(count, 3, device=dev, dtype=torch.float32) if want_vecs else None
        else:
            df, vecs = out
        H = torch.empty(count, 3, 3, device=dev, dtype=torch.float32) if want_hess else None
        with torch.cuda.device(dev):
            _lib.check(self.L.dudf_query_grid(self.h, N, first, count, flags, float(alpha), _lib.ptr(df), _lib.ptr(vecs), _lib.ptr(H),
                                              PRECISIONS[precision], _lib.current_stream()), "dudf_query_grid")
        return df, vecs, H

    def eig_normals(self, H, ref_dir=None, ref_mode=0, want_dirs=False, want_lam=False):
        P = H.shape[0]
        n = torch.empty(P, 3, device=H.device, dtype=torch.float32)
        dirs = torch.empty(P, 3, 2, device=H.device, dtype=torch.float32) if want_dirs else None
        lam = torch.empty(P, 3, device=H.device, dtype=torch.float32) if want_lam else None
        with torch.cuda.device(H.device):
            _lib.check(self.L.dudf_eig_normals(H.data_ptr(), _lib.ptr(ref_dir), ref_mode, P, n.data_ptr(), _lib.ptr(dirs), _lib.ptr(lam),
                                               _lib.current_stream()), "dudf_eig_normals")
        return n, dirs, lam

    def curvature(self, H, T):
        P = H.shape[0]
        n = torch.empty(P, 3, device=H.device, dtype=torch.float32)
        mean = torch.empty(P, device=H.device, dtype=torch.float32)
        gauss = torch.empty(P, device=H.device, dtype=torch.float32)
        J = torch.empty(P, 3, 3, device=H.device, dtype=torch.float32)
        with torch.cuda.device(H.device):
            _lib.check(self.L.dudf_curvature(H.data_ptr(), T.data_ptr(), P, n.data_ptr(), mean.data_ptr(), gauss.data_ptr(),
                                             J.data_ptr(), _lib.current_stream()), "dudf_curvature")
        return n, mean, gauss, J

    def mean_curvature(self, x):
        """Eigen-normals, principal directions and mean curvature at x (P,3) on tensor cores (dudf_mean_curvature): Hessian jet,
        eigen-solve and the 10-channel directional third-order jet.  Returns (normals (P,3), dirs (P,3,2), mean (P,))."""
        P = x.shape[0]
        n = torch.empty(P, 3, device=x.device, dtype=torch.float32)
        dirs = torch.empty(P, 3, 2, device=x.device, dtype=torch.float32)
        mean = torch.empty(P, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(self.L.dudf_mean_curvature(self.h, x.data_ptr(), P, n.data_ptr(), dirs.data_ptr(), mean.data_ptr(),
                                                  _lib.current_stream()), "dudf_mean_curvature")
        return n, dirs, mean

    def field_vectors(self, g, H):
        vecs = torch.empty_like(g)
        with torch.cuda.device(g.device):
            _lib.check(self.L.dudf_field_vectors(g.data_ptr(), H.data_ptr(), g.shape[0], vecs.data_ptr(), _lib.current_stream()),
                       "dudf_field_vectors")
        return vecs

    def evaluate_host(self, x_host, order, f_host, g_host, H_host, max_batch, precision="fp32"):
        with torch.cuda.device(self.device):
            _lib.check(self.L.dudf_evaluate_host(self.h, x_host.ctypes.data, x_host.shape[0], order, _lib.ptr(f_host), _lib.ptr(g_host),
                                                 _lib.ptr(H_host), max_batch, PRECISIONS[precision], _lib.current_stream()),
                       "dudf_evaluate_host")

    # ---- training primitives ----
    def stash_columns(self, order, P, precision="fp32"):
        return int(self.L.dudf_stash_columns(order, P, PRECISIONS[precision]))

    def jet_forward(self, x, order, packed, Z, A, ld, col0, precision="fp32"):
        with torch.cuda.device(x.device):
            _lib.check(self.L.dudf_jet_forward(self.h, x.data_ptr(), x.shape[0], order, packed.data_ptr(), Z.data_ptr(), A.data_ptr(), ld,
                                               col0, PRECISIONS[precision], _lib.current_stream()), "dudf_jet_forward")

    def jet_backward(self, x, order, seeds, Z, Zb, ld, col0, gW, gb, precision="fp32", seed_absmax=None):
        with torch.cuda.device(x.device):
            _lib.check(self.L.dudf_jet_backward(self.h, x.data_ptr(), x.shape[0], order, seeds.data_ptr(), _lib.ptr(seed_absmax),
                                                Z.data_ptr(), Zb.data_ptr(), ld, col0, _ptr_array(gW), _ptr_array(gb),
                                                PRECISIONS[precision], _lib.current_stream()), "dudf_jet_backward")

    @staticmethod
    def _segments(segs, packed_key):
        arr = (_lib.Segment * len(segs))()
        for i, s in enumerate(segs):
            arr[i].x = s["x"].data_ptr()
            arr[i].rows = s["x"].shape[0]
            arr[i].order = s["order"]
            arr[i].packed = s["packed"].data_ptr() if packed_key == "packed" else None
            arr[i].seeds = s["seeds"].data_ptr() if packed_key == "seeds" else None
            arr[i].col0 = s["col0"]
        return arr

    def jet_forward_multi(self, segs, Z, A, ld, precision="fp32"):
        """segs: list of dict(x (rows,3), order, packed (rows*NCH,), col0) — one launch on the tensor-core path."""
        with torch.cuda.device(Z.device):
            _lib.check(self.L.dudf_jet_forward_multi(self.h, self._segments(segs, "packed"), len(segs), Z.data_ptr(), A.data_ptr(), ld,
                                                     PRECISIONS[precision], _lib.current_stream()), "dudf_jet_forward_multi")

    def jet_backward_multi(self, segs, Z, Zb, ld, gW, gb, precision="fp32", seed_absmax=None):
        with torch.cuda.device(Z.device):
            _lib.check(self.L.dudf_jet_backward_multi(self.h, self._segments(segs, "seeds"), len(segs), _lib.ptr(seed_absmax), Z.data_ptr(),
                                                      Zb.data_ptr(), ld, _ptr_array(gW), _ptr_array(gb), PRECISIONS[precision],
                                                      _lib.current_stream()), "dudf_jet_backward_multi")

    def jet_wgrad(self, Zb, A, ld, ncols, gW, precision="fp32", seed_absmax=None):
        with torch.cuda.device(Zb.device):
            _lib.check(self.L.dudf_jet_wgrad(self.h, Zb.data_ptr(), A.data_ptr(), ld, ncols, _lib.ptr(seed_absmax), _ptr_array(gW),
                                             PRECISIONS[precision], _lib.current_stream()), "dudf_jet_wgrad")

    def jet_wgrad_layers(self, Zb, A, ld, gW, layer_lo, layer_hi, precision="tc16", seed_absmax=None):
        """weight gradients of the hidden layers [layer_lo, layer_hi) only (tensor-core precisions)"""
        with torch.cuda.device(Zb.device):
            _lib.check(self.L.dudf_jet_wgrad_layers(self.h, Zb.data_ptr(), A.data_ptr(), ld, _lib.ptr(seed_absmax), _ptr_array(gW),
                                                    int(layer_lo), int(layer_hi), PRECISIONS[precision], _lib.current_stream()),
                       "dudf_jet_wgrad_layers")

    def fused_scratch_bytes(self):
        return int(self.L.dudf_fused_scratch_bytes(self.h))

    def train_step_fused(self, mode, segs, P_global, w, alpha, terms, amax_prev, amax_next, scratch, A, Zb, ld, gW, gb, flags=3):
        """loss_s1 / loss_siren forward + loss + reverse sweep + weight gradients, tensor-core path, two launches.
        segs: list of dict(x, normals, d, order[, packed])."""
        arr = (_lib.TrainSegment * len(segs))()
        for i, s in enumerate(segs):
            arr[i].x, arr[i].normals, arr[i].dist = s["x"].data_ptr(), s["normals"].data_ptr(), s["d"].data_ptr()
            arr[i].rows, arr[i].order = s["x"].shape[0], s["order"]
            arr[i].packed = s["packed"].data_ptr() if s.get("packed") is not None else None
        w4 = (ctypes.c_float * 4)(*([float(v) for v in w] + [0.0] * (4 - len(w))))
        with torch.cuda.device(A.device):
            _lib.check(self.L.dudf_train_step_fused(self.h, LOSS_MODES[mode], arr, len(segs), P_global, w4, float(alpha), terms.data_ptr(),
                                                    amax_prev.data_ptr(), amax_next.data_ptr(), scratch.data_ptr(), A.data_ptr(),
                                                    Zb.data_ptr(), ld, _ptr_array(gW), _ptr_array(gb), flags, _lib.current_stream()),
                       "dudf_train_step_fused")

    def loss(self, mode, packed, nch, normals, dist, P, P_global, w, alpha, upstream=None, seeds=None, terms=None, s2_stats=None,
             seed_absmax=None):
        w4 = (ctypes.c_float * 4)(*([float(v) for v in w] + [0.0] * (4 - len(w))))
        with torch.cuda.device(packed.device):
            _lib.check(self.L.dudf_loss(LOSS_MODES[mode], packed.data_ptr(), nch, _lib.ptr(normals), dist.data_ptr(), P, P_global, w4,
                                        float(alpha), _lib.ptr(upstream), _lib.ptr(seeds), _lib.ptr(seed_absmax), _lib.ptr(terms),
                                        _lib.ptr(s2_stats), _lib.current_stream()), "dudf_loss")

    def loss_s2_stats(self, packed, dist, P, stats):
        with torch.cuda.device(packed.device):
            _lib.check(self.L.dudf_loss_s2_stats(packed.data_ptr(), dist.data_ptr(), P, stats.data_ptr(), _lib.current_stream()),
                       "dudf_loss_s2_stats")

    def loss_s2_finish(self, stats, w0, w1, terms):
        with torch.cuda.device(stats.device):
            _lib.check(self.L.dudf_loss_s2_finish(stats.data_ptr(), float(w0), float(w1), terms.data_ptr(), _lib.current_stream()),
                       "dudf_loss_s2_finish")


def adam_step(p, g, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-8, unsafe_flag=None, skipped=None):
    """In-place Adam on flat fp32 CUDA tensors (train.py:334-337 semantics).  unsafe_flag (1-element fp32 device tensor): the
    update is skipped, and `skipped` (1-element int64 device tensor) incremented, when it is non-zero (dudf_adam_step_guarded)."""
    L = _lib.lib()
    with torch.cuda.device(p.device):
        if unsafe_flag is None:
            _lib.check(L.dudf_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr), float(beta1),
                                        float(beta2), float(eps), int(t), _lib.current_stream()), "dudf_adam_step")
        else:
            _lib.check(L.dudf_adam_step_guarded(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr), float(beta1),
                                                float(beta2), float(eps), int(t), unsafe_flag.data_ptr(), _lib.ptr(skipped),
                                                _lib.current_stream()), "dudf_adam_step_guarded")


def adam_step_dev(p, g, m, v, state, beta1=0.9, beta2=0.999, eps=1e-8):
    """Adam with the learning rate and the step count read from (and the count advanced in) the 6-float device array `state`
    (dudf_adam_step_dev): the form a CUDA graph of the training step can replay."""
    L = _lib.lib()
    with torch.cuda.device(p.device):
        _lib.check(L.dudf_adam_step_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), state.data_ptr(), float(beta1),
                                        float(beta2), float(eps), _lib.current_stream()), "dudf_adam_step_dev")


def adam_step_peers(p, peer_ptrs, world, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-8, guarded=False, skipped=None, g_sum_out=None):
    """Adam with the gradient all-reduce fused in (dudf_adam_step_peers): peer_ptrs = ctypes array of `world` device pointers to
    the ranks' (n + 1)-float gradient buffers in peer-mapped memory; the caller has ordered a cross-rank barrier before."""
    L = _lib.lib()
    with torch.cuda.device(p.device):
        _lib.check(L.dudf_adam_step_peers(p.data_ptr(), peer_ptrs, int(world), m.data_ptr(), v.data_ptr(), p.numel(), float(lr), float(beta1),
                                          float(beta2), float(eps), int(t), 1 if guarded else 0, _lib.ptr(skipped), _lib.ptr(g_sum_out),
                                          _lib.current_stream()), "dudf_adam_step_peers")


def scale_guard(amax_prev, amax_next, flag, limit=16384.0):
    """flag += 1 when this step's seeds outgrew the loss scale derived from the previous step (dudf_scale_guard)."""
    L = _lib.lib()
    with torch.cuda.device(flag.device):
        _lib.check(L.dudf_scale_guard(amax_prev.data_ptr(), amax_next.data_ptr(), float(limit), flag.data_ptr(), _lib.current_stream()),
                   "dudf_scale_guard")
