"""Training loop: drop-in for the step / schedule semantics of /root/reference/train.py
(train_model_tanh :146-283, train_model_siren :23-143, Adam as built at :334-337,365-368).

`FusedTrainer.step` is the fast path: no autograd graph, no host synchronisation — jet forward,
loss epilogue, reverse sweep, weight gradients, (all-reduce), Adam, all enqueued on the current
stream.  The reference's per-step `.item()` calls become one read-back per epoch.
File IO (TensorBoard, checkpoints every epoch, plots, meshing) is host glue and is left to the caller.
"""
import copy
import os
import time

import numpy as np
import torch

from .engine import adam_step, adam_step_dev, adam_step_peers, scale_guard
from .loss_functions import S1_KEYS, S2_KEYS, SIREN_KEYS, TrainCore


def lr_for_epoch(epoch, epochs, s1_epochs, warmup_epochs, warmup_lr, lr_s1, lr_s2):
    """Learning rate in effect during `epoch` (train.py:167-191)."""
    if epoch >= s1_epochs:
        return 0.5 * (np.cos(epoch / (epochs - s1_epochs) * np.pi) + 1) * lr_s2
    if epoch >= warmup_epochs:
        return lr_s1
    return warmup_lr


class FusedTrainer:
    """Owns flat fp32 parameter / gradient / Adam-moment buffers; the module's parameters become views
    of the flat buffer so that state_dict()/load_state_dict() keep working."""

    def __init__(self, model, betas=(0.9, 0.999), eps=1e-8, dp=None, precision=None, fused=True, graph=False):
        """graph=True (single GPU): a step's launches — weight re-pack, forward, loss, reverse sweep, weight gradient, Adam — are
        captured once per (mode, batch shape, loss weights) into a CUDA graph and replayed; the scalars that change from step to step
        (learning rate, Adam's step count) live on the device (dudf_adam_step_dev).  A step that is dominated by launch overhead
        (loss_s2: a dozen small kernels in ~0.1 ms of device time) no longer waits for the host."""
        self.model = model
        self.fused = fused                  # tensor-core steps of loss_s1 / loss_siren go through the single fused launch
        self.graph = bool(graph)
        self._graphs = {}                   # key -> dict(graph, x, n, d, terms) | "warm" (one eager step taken with this key)
        self._adam_state = None             # device [lr, -, count lo, count hi, ticket, -] of dudf_adam_step_dev
        self._dev_t = -1                    # host's idea of the device step count (-1: stale)
        self._dev_lr = None
        self.betas, self.eps, self.dp = betas, eps, dp
        ws, bs = model._weights_biases()
        dev = ws[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedTrainer needs the model on a CUDA (sm_100) device")
        tensors = []
        for w, b in zip(ws, bs):
            tensors += [w, b]
        n = sum(t.numel() for t in tensors)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.n = n
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.skipped = torch.zeros(1, device=dev, dtype=torch.int64)
        off = 0
        self.slices = []
        for i, t in enumerate(tensors):
            k = t.numel()
            self.flat[off:off + k].copy_(t.detach().reshape(-1))
            t.data = self.flat[off:off + k].view_as(t)
            self.slices.append((off, k, tuple(t.shape)))
            off += k
        # Gradient buffers.  One slot behind the flat gradient carries the "loss scale outgrown" flag of the fused step: zeroed
        # with the gradient, summed over the ranks with it, read by the guarded Adam kernel.
        # Data parallel, peer mode (default when torch's symmetric memory can map the ranks' buffers; DUDF_DP_PEER=0 disables):
        # two buffer sets in peer-mapped memory, alternated per step, reduced INSIDE the Adam kernel (dudf_adam_step_peers).
        self.peer = None
        if dp is not None and dp.world > 1 and os.environ.get("DUDF_DP_PEER", "1") != "0":
            try:
                from .parallel import PeerGradients
                self.peer = PeerGradients(dp, n, dev)
            except Exception as exc:            # no peer access / symmetric memory unavailable: NCCL all-reduce path
                import warnings
                warnings.warn(f"diffudf_b200: peer-memory gradient exchange unavailable ({exc!r}); using the NCCL all-reduce")
                self.peer = None
        self.sets = []
        for k in range(2 if self.peer is not None else 1):
            ga = self.peer.local(k) if self.peer is not None else torch.zeros(n + 1, device=dev, dtype=torch.float32)
            gW, gB = [], []
            for i, (o, cnt, shape) in enumerate(self.slices):
                (gW if i % 2 == 0 else gB).append(ga[o:o + cnt].view(shape))
            self.sets.append((ga, gW, gB))
        self.grad_sum = None                     # peer mode: set to a (n,) tensor to receive the summed gradient of every step
        self._use_set(0)
        self.core = TrainCore(model, precision)
        self.t = 0
        # NCCL route (no peer memory): one all-reduce of the flat gradient.  DUDF_DP_GROUPS=g > 1 launches the weight-gradient GEMMs
        # in g layer groups and all-reduces each finished group on a side stream under the next group's GEMM — measured SLOWER at
        # N = 2 (1.250 vs 1.226 ms / step: the collective is latency-bound, two of them cost more than one hides), so it is opt-in
        self.groups, self.side = None, None
        if dp is not None:
            from .parallel import grad_groups
            ng = int(os.environ.get("DUDF_DP_GROUPS", "1"))
            if ng > 1:
                self.groups = grad_groups([(w.numel(), b.numel()) for w, b in zip(ws, bs)], ng)
                self.side = torch.cuda.Stream(dev)

    def _use_set(self, k):
        self.grad_all, self.gW, self.gB = self.sets[k]
        self.grad = self.grad_all[:self.n]

    GRAPH_LIMIT = 8            # distinct step signatures captured before falling back to eager launches

    def _graph_step(self, mode, x, normals, d, n_on, weights, alpha, lr):
        """Replay (or, the second time a signature is seen, capture) the step as a CUDA graph.  Returns None when this step must be
        launched eagerly (first occurrence of a signature: it also creates every cached workspace the capture will refer to)."""
        key = (mode, tuple(x.shape), int(n_on), tuple(float(w) for w in weights), float(alpha), self.core._prec())
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= self.GRAPH_LIMIT:
                return None
            self._graphs[key] = "warm"
            return None
        if g == "eager":                                       # a capture of this signature failed before
            return None
        dev = x.device
        if self._adam_state is None:
            self._adam_state = torch.zeros(6, device=dev, dtype=torch.float32)
        if self._dev_lr != float(lr):
            self._adam_state[0:1].fill_(float(lr))
            self._dev_lr = float(lr)
        if self._dev_t != self.t:                              # eager steps were taken in between
            self._adam_state[2:4].view(torch.int64).fill_(self.t)
            self._dev_t = self.t
        if g == "warm":
            sx, sn, sd = torch.empty_like(x), torch.empty_like(normals), torch.empty_like(d)
            # the core's cached workspaces of THIS signature exist before the capture starts (a step of another shape may have evicted
            # them since the warm-up step): an allocation inside the capture would put its zero-fill — 1.3 GB of operand images — into
            # the graph and repeat it on every replay
            prec = self.core._prec()
            _, ld, nout = self.core.plan(mode, x.shape[0], n_on, weights, prec)
            self.core._stashes(self.model.n_hidden, ld, prec)
            self.core._buf("packed", (nout,))
            self.core._buf("seeds", (nout,))
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(dev)
            try:
                with torch.cuda.graph(graph):
                    self.grad_all.zero_()
                    terms = self.core.forward(mode, sx, sn, sd, n_on, weights, alpha, None, None, eager_seeds=True)
                    self.core.backward(None, self.gW, self.gB)
                    adam_step_dev(self.flat, self.grad, self.m, self.v, self._adam_state, self.betas[0], self.betas[1], self.eps)
            except Exception as exc:                           # nothing was executed: launch this signature eagerly from now on
                import warnings
                warnings.warn(f"diffudf_b200: CUDA-graph capture of the {mode} step failed ({exc!r}); launching it eagerly")
                self._graphs[key] = "eager"
                self.core.pending = None
                self._dev_t = -1
                return None
            # the graph refers to the core's cached workspaces (stash, operand images, packed jets, seeds): keep them alive when a
            # step of another shape evicts them from the cache
            g = dict(graph=graph, x=sx, n=sn, d=sd, terms=terms, keep=list(self.core.ws.values()))
            self._graphs[key] = g
        g["x"].copy_(x)
        g["n"].copy_(normals)
        g["d"].copy_(d)
        g["graph"].replay()
        self.t += 1
        self._dev_t = self.t
        if self.model._engine is not None:                     # parameters changed in place: eager users re-pack on next use
            self.model._engine._sig = None
        return g["terms"].clone()

    def step(self, mode, x, normals, d, n_on, weights, alpha, lr):
        """One optimisation step on a device-resident batch (x (P,3), normals (P,3), d (P,), fp32).
        Returns the (4,) float64 device tensor of this rank's loss-term shares (no sync)."""
        dp = self.dp
        if self.graph and dp is None and (mode == "s2" or self.core._prec() != "tc16" or not self.fused):
            terms = self._graph_step(mode, x, normals, d, n_on, weights, alpha, lr)
            if terms is not None:
                return terms
        P_global = dp.global_rows(x.shape[0]) if dp is not None else None
        if self.peer is not None:
            self._use_set(self.t % 2)
        self.grad_all.zero_()
        guarded = False
        if mode != "s2" and self.core._prec() == "tc16" and self.fused:
            terms = self.core.fused_step(mode, x, normals, d, n_on, weights, alpha, P_global, self.gW, self.gB)
            if self.core.last_fused is not None:           # the single-launch kernel ran with the previous step's loss scale
                scale_guard(*self.core.last_fused, self.grad_all[-1:])
                guarded = True
        else:
            terms = self.core.forward(mode, x, normals, d, n_on, weights, alpha, P_global, dp.reduce_stats if dp is not None else None,
                                      eager_seeds=True)
            if self.peer is not None:
                self.core.backward(None, self.gW, self.gB)
            elif self.groups is not None and self.core._prec() in ("tc16", "tcx3"):
                self.core.backward(None, self.gW, self.gB, wgrad_groups=self.groups,
                                   after_group=lambda a, b: dp.reduce_grads_behind(self.grad[a:b], self.side))
                torch.cuda.current_stream(self.grad.device).wait_stream(self.side)
                dp = None                                  # reduced
            else:
                self.core.backward(None, self.gW, self.gB)
        self.t += 1
        if self.peer is not None:
            self.peer.barrier()                    # every rank's gradient is complete
            adam_step_peers(self.flat, self.peer.ptrs[(self.t - 1) % 2], self.peer.world, self.m, self.v, lr, self.t, self.betas[0],
                            self.betas[1], self.eps, guarded=guarded, skipped=self.skipped, g_sum_out=self.grad_sum)
        else:
            if dp is not None:
                dp.reduce_grads(self.grad_all if guarded else self.grad)
            adam_step(self.flat, self.grad, self.m, self.v, lr, self.t, self.betas[0], self.betas[1], self.eps,
                      unsafe_flag=self.grad_all[-1:] if guarded else None, skipped=self.skipped if guarded else None)
        # parameters changed in place behind torch's back: make the engine re-pack on next use
        if self.model._engine is not None:
            self.model._engine._sig = None
        return terms

    def skipped_steps(self):
        """Number of fused steps whose update was skipped because the seeds outgrew the (one step stale) loss scale (one sync)."""
        return int(self.skipped.item())


class BatchFeeder:
    """Double-buffered host -> device feed of (coords, normals, dist) batches: the copy of batch i+1 runs on a side
    stream while step i computes (the reference copies synchronously inside the loop, train.py:200-202).
    Host tensors should be pinned for the copies to be asynchronous."""

    def __init__(self, device, n_buffers=2):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.n = n_buffers
        self.bufs = [None] * n_buffers
        self.ready = [torch.cuda.Event() for _ in range(n_buffers)]
        self.free = [torch.cuda.Event() for _ in range(n_buffers)]
        self.i = 0

    def _stage(self, slot, batch):
        x, n, d = batch
        shapes = ((x.numel() // 3, 3), (n.numel() // 3, 3), (d.numel(),))
        if all(torch.is_tensor(t) and t.device == self.device for t in (x, n, d)):
            # produced on the device (diffudf_b200.dataset.PointCloud) in the consumer's stream order: nothing to copy
            self.bufs[slot] = tuple(t.reshape(s) for t, s in zip((x, n, d), shapes))
            self.ready[slot].record(torch.cuda.current_stream(self.device))
            return
        if self.bufs[slot] is None or any(b.shape != s for b, s in zip(self.bufs[slot], shapes)):
            self.bufs[slot] = tuple(torch.empty(s, device=self.device, dtype=torch.float32) for s in shapes)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])        # the step that used this buffer has been enqueued and finished
            for dst, src in zip(self.bufs[slot], (x, n, d)):
                dst.copy_(src.reshape(dst.shape), non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def feed(self, batches):
        """Yields device batches (x (P,3), normals (P,3), dist (P,)); call `release()` semantics are automatic:
        a buffer is recycled after the NEXT yielded batch's step has been enqueued."""
        it = iter(batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        slot = self.i % self.n
        self._stage(slot, nxt)
        while True:
            cur_slot = slot
            try:
                nxt = next(it)
                slot = (cur_slot + 1) % self.n
                self._stage(slot, nxt)
                more = True
            except StopIteration:
                more = False
            torch.cuda.current_stream(self.device).wait_event(self.ready[cur_slot])
            yield self.bufs[cur_slot]
            self.free[cur_slot].record(torch.cuda.current_stream(self.device))
            self.i += 1
            if not more:
                return


def _epoch_loop(dataset, model, device, config, stage_fn):
    """config['dp'] (parallel.DataParallel): every rank iterates the SAME dataset (same seed) and trains on its share of each row
    group of every batch (parallel.shard_batch), so the job as a whole sees the reference's batches; the loss means are taken
    over the global row count.  Ignored keys of the reference's config (`optimizer.type`, `log_path`, checkpoint / plot settings:
    file IO and TensorBoard are left to the caller) raise nothing; an optimizer other than Adam raises."""
    epochs = config["epochs"]
    opt = config.get("optimizer")
    if isinstance(opt, dict) and str(opt.get("type", "adam")).lower() != "adam":
        raise NotImplementedError(f"optimizer {opt.get('type')!r}: the fused loop implements torch.optim.Adam (train.py:334-337)")
    dp = config.get("dp")
    # single GPU: every step signature of the schedule (loss_s1, loss_s2) is captured once as a CUDA graph and replayed
    # (config['cuda_graph'] = False launches eagerly)
    trainer = FusedTrainer(model.to(device), dp=dp, precision=config.get("precision"), graph=bool(config.get("cuda_graph", dp is None)))
    feeder = BatchFeeder(device)
    losses, best_loss, best_weights = {}, np.inf, None
    torch.cuda.synchronize(device)
    start = time.time()
    for epoch in range(epochs):
        mode, keys, weights, lr = stage_fn(epoch)
        running = torch.zeros(4, device=device, dtype=torch.float64)
        misplaced = torch.zeros(1, device=device, dtype=torch.int64)      # rows that contradict the [on | off] layout (read per epoch)
        for x, n, d in feeder.feed((tuple(torch.as_tensor(t, dtype=torch.float32) for t in b) for b in iter(dataset))):
            n_on = getattr(dataset, "samplesOnSurface", None)
            need_on = mode == "s1" and weights[2] != 0
            if n_on is None and need_on:
                from .loss_functions import on_surface_prefix
                n_on = on_surface_prefix(d)
                if n_on is None:
                    raise RuntimeError("batches must be ordered [on-surface | off-surface] for the fused step")
            elif n_on is not None and (need_on or mode == "s2"):      # (s2 evaluates the leading on-surface rows only when it knows them)
                # the reference masks the alignment term on d == 0 (loss_functions.py:45-53); the fused step takes the first n_on
                # rows: count disagreements on the device, raise at the epoch's read-back
                misplaced += (d[:n_on] != 0).sum() + (d[n_on:] == 0).sum()
            n_on = n_on or 0
            if dp is not None and dp.world > 1:
                from .parallel import shard_batch
                P = x.shape[0]
                n_far = (P - n_on) // 2                 # [on | far | near] with far = half of the off-surface rows (src/dataset.py:92-94)
                dp.rows_global = P
                x, n, d, n_on = shard_batch(x, n, d, n_on, n_far, dp.rank, dp.world)
                x, n, d = x.contiguous(), n.contiguous(), d.contiguous()
            running += trainer.step(mode, x, n, d, n_on, weights, config.get("alpha", 0.0), lr)
        if dp is not None and dp.world > 1 and mode != "s2":
            dp.reduce_terms(running)            # shares of the means; loss_s2's terms come from the all-reduced statistics and are already global
        vals = running.cpu().numpy()                        # one read-back per epoch
        if int(misplaced.item()) != 0:
            raise RuntimeError(f"{int(misplaced.item())} rows of this epoch's batches contradict the [on-surface (d == 0) | off-surface] layout "
                               "the fused loss_s1 / loss_s2 steps rely on (dataset.samplesOnSurface leading rows)")
        epoch_loss = 0.0
        for i, k in enumerate(keys):
            losses.setdefault(k, [0.0] * epochs)[epoch] = float(vals[i])
            epoch_loss += float(vals[i])
        epoch_loss /= getattr(dataset, "batchesPerEpoch", 1)
        if epoch_loss < best_loss:
            best_loss = epoch_loss
            best_weights = copy.deepcopy(model.state_dict())
    torch.cuda.synchronize(device)
    return losses, best_weights, time.time() - start


def train_model_tanh(dataset, model, device, config):
    """Two-stage hyperbolic training (train.py:146-283): loss_s1 with warm-up lr, then loss_s2 with cosine lr."""
    def stage(epoch):
        lr = lr_for_epoch(epoch, config["epochs"], config["s1_epochs"], config["warmup_epochs"], config["warmup_lr"],
                          config["lr_s1"], config["lr_s2"])
        if epoch >= config["s1_epochs"]:
            return "s2", S2_KEYS, list(config["loss_s2_weights"]), lr
        return "s1", S1_KEYS, list(config["loss_s1_weights"]), lr
    return _epoch_loop(dataset, model, device, config, stage)


def train_model_siren(dataset, model, device, config):
    """SDF baseline (train.py:23-143): loss_siren, constant lr."""
    def stage(epoch):
        return "siren", SIREN_KEYS, list(config["loss_weights"]), config["lr"]
    return _epoch_loop(dataset, model, device, config, stage)
