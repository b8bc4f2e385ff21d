"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch on
B200 boxes, gloo in the CPU tests).  The path shards by rows:

* training: each of the three row groups [on | far | near] is split evenly over the ranks so the
  on-surface mask stays balanced; every rank computes sum/P_global and its weight gradient; ONE
  all-reduce(sum) of the flat 461 825-float gradient per step (plus 3 doubles for loss_s2's
  statistics before its backward);
* grid / ray / point queries: contiguous ranges of the flat index, all-gather of the outputs.
No other collective exists on this path.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) share of n items for `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, normals, d, n_on, n_far, rank, world):
    """Split a [on | far | near] batch (P rows along dim -2 or 0 of x) into this rank's
    [on_r | far_r | near_r].  Works on numpy arrays or torch tensors of shape (P,3)/(P,3)/(P,)
    Returns (x_r, normals_r, d_r, n_on_r)."""
    P = x.shape[0]
    groups = [(0, n_on), (n_on, n_on + n_far), (n_on + n_far, P)]
    xs, ns, ds = [], [], []
    n_on_r = 0
    for gi, (a, b) in enumerate(groups):
        lo, hi = shard_range(b - a, rank, world)
        xs.append(x[a + lo:a + hi])
        ns.append(normals[a + lo:a + hi])
        ds.append(d[a + lo:a + hi])
        if gi == 0:
            n_on_r = hi - lo
    cat = torch.cat if torch.is_tensor(x) else np.concatenate
    return cat(xs), cat(ns), cat(ds), n_on_r


def grad_groups(sizes, n_groups):
    """Layer groups for the overlapped gradient all-reduce.  sizes: [(W.numel(), b.numel())] per linear layer 0..L in the order of
    the flat gradient [W0 b0 W1 b1 ... WL bL]; the 256 x 256 weight-gradient GEMMs are layers 1..L-1.  Returns
    [(layer_lo, layer_hi, flat_lo, flat_hi)]: after the GEMMs of layers [layer_lo, layer_hi) the slice [flat_lo, flat_hi) is final
    (the first group carries layer 0, the last one the output layer: those gradients come out of the reverse sweep)."""
    L = len(sizes) - 1
    n_groups = max(1, min(n_groups, L - 1))
    offs = [0]
    for a, b in sizes:
        offs.append(offs[-1] + a + b)
    bounds = [1 + (L - 1) * g // n_groups for g in range(n_groups + 1)]
    out = []
    for g in range(n_groups):
        lo, hi = bounds[g], bounds[g + 1]
        out.append((lo, hi, 0 if g == 0 else offs[lo], offs[-1] if g == n_groups - 1 else offs[hi]))
    return out


class DataParallel:
    """Attach to a model (`model._dp = DataParallel(...)`) to make the fused losses data-parallel."""

    def __init__(self, group=None, rows_global=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows_global = rows_global

    def global_rows(self, local_rows):
        """Divisor of the loss means.  Given explicitly (uneven shards) or local_rows * world."""
        return self.rows_global if self.rows_global is not None else local_rows * self.world

    def reduce_stats(self, stats):
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)

    def reduce_grads(self, flat_grad):
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)

    def reduce_grads_behind(self, flat_slice, side):
        """all-reduce of a finished slice of the flat gradient on the stream `side`, ordered behind the work enqueued so far on
        the current stream; the current stream keeps going (the next group's weight-gradient GEMM runs under the collective)."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(flat_slice.device))
        side.wait_event(ev)
        with torch.cuda.stream(side):
            dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, group=self.group)

    def reduce_terms(self, terms):
        dist.all_reduce(terms, op=dist.ReduceOp.SUM, group=self.group)

    def gather_rows(self, local, counts):
        """all-gather of row-sharded outputs with per-rank row counts (list of ints)."""
        mx = max(counts)
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        pad[:local.shape[0]] = local
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def round_robin_blocks(total, world, rounds):
    """Block decomposition for a gathered query: the flat index range is cut into world * rounds blocks of `s` rows; block b is
    computed by rank b % world in round b // world, so the blocks of one round are CONTIGUOUS in the output and one
    all_gather_into_tensor per round lands them in place (no padding pass, no list of buffers, no concatenation).
    Returns the block size s (the last blocks may be short or empty)."""
    return -(-total // (world * rounds))


def gather_round_robin(fill, total, dp, tails, dtypes, device, rounds=4):
    """Sharded evaluation + gather with the collective of round k overlapped with the compute of round k + 1.

    fill(first, count, views): writes rows [first, first + count) of the result into `views` (one (s,) + tail view of each output,
    already positioned at this rank's block).  Outputs are allocated once at world * rounds * s rows and returned as [:total]
    views.  The gathers run on a side stream (CUDA) behind an event, in place: every rank's input is its own block of the
    output tensor."""
    world, rank = dp.world, dp.rank
    s = round_robin_blocks(total, world, rounds)
    outs = [torch.empty((world * rounds * s,) + tuple(t), dtype=dt, device=device) for t, dt in zip(tails, dtypes)]
    cuda = torch.device(device).type == "cuda"
    side = torch.cuda.Stream(device) if cuda else None
    for k in range(rounds):
        first = (k * world + rank) * s
        count = max(0, min(s, total - first))
        views = [o[first:first + s] for o in outs]
        if count > 0:
            fill(first, count, views)
        if cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            side.wait_event(ev)
            with torch.cuda.stream(side):
                for o, v in zip(outs, views):
                    dist.all_gather_into_tensor(o[k * world * s:(k + 1) * world * s], v, group=dp.group)
        else:
            for o, v in zip(outs, views):
                dist.all_gather_into_tensor(o[k * world * s:(k + 1) * world * s], v.clone(), group=dp.group)
    if cuda:
        torch.cuda.current_stream(device).wait_stream(side)
    return [o[:total] for o in outs]


class PeerGradients:
    """Two flat gradient buffers of n + 1 floats per rank in symmetric (peer-mapped) memory — the operand of the fused
    all-reduce + Adam kernel (dudf_adam_step_peers).  torch.distributed._symmetric_memory provides the allocation, the exchange
    of the mappings over NVLink and the cross-rank barrier (plumbing); the reduction itself is our kernel reading the peers.
    Buffers alternate per step: a rank may start accumulating step t + 1 while a slower peer still reads step t."""

    def __init__(self, dp, n, device):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        self.n = int(n)
        self.stride = (self.n + 1 + 3) // 4 * 4
        group = dp.group if dp.group is not None else dist.group.WORLD
        self.buf = symm.empty(2 * self.stride, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        self.world = dp.world
        base = [int(self.hdl.buffer_ptrs[r]) for r in range(self.world)]
        self.ptrs = [(ctypes.c_void_p * self.world)(*[b + k * self.stride * 4 for b in base]) for k in (0, 1)]
        torch.cuda.synchronize(device)
        self.hdl.barrier(channel=0)

    def local(self, k):
        return self.buf[k * self.stride:k * self.stride + self.n + 1]

    def barrier(self):
        """stream-ordered: returns (on the device) when every rank's work enqueued before its own barrier has finished"""
        self.hdl.barrier(channel=0)


def extract_fields_sharded(model, N, gt_mode, alpha, dp, rounds=4):
    """Grid query sharded over the ranks in round-robin blocks of the flat index and gathered on every rank straight into the
    (N^3,) / (N^3, 3) outputs; the gather of one round runs under the next round's compute.  Returns df (N,N,N), vecs (N,N,N,3)."""
    from .render_mc import extract_fields
    dev = model._weights_biases()[0][0].device

    def fill(first, count, views):
        extract_fields(model, None, N, gt_mode, None, alpha, first=first, count=count, out=(views[0][:count], views[1][:count]))

    df, vecs = gather_round_robin(fill, N ** 3, dp, [(), (3,)], [torch.float32, torch.float32], dev, rounds)
    return df.reshape(N, N, N), vecs.reshape(N, N, N, 3)


def propagate_rays_sharded(model, rays, t0, mask_rays, network_config, rendering_config, dp):
    """Sphere tracing sharded by contiguous ray ranges: every rank marches its range with the device loop of
    render_st and the hit mask / positions / still-active mask are all-gathered.  Same in-place contract as
    render_st.propagate_rays on every rank; rays are independent, so the result equals the single-GPU one bit for bit."""
    import numpy as np
    from .render_st import _march
    dev = model._weights_biases()[0][0].device
    R = t0.shape[0]
    lo, hi = shard_range(R, dp.rank, dp.world)
    counts = [shard_range(R, r, dp.world)[1] - shard_range(R, r, dp.world)[0] for r in range(dp.world)]
    rays_d = torch.from_numpy(np.ascontiguousarray(rays[lo:hi], dtype=np.float64)).to(dev)
    t0_d = torch.from_numpy(np.ascontiguousarray(t0[lo:hi], dtype=np.float64)).to(dev)
    idx = torch.nonzero(torch.from_numpy(np.asarray(mask_rays[lo:hi], dtype=bool)).to(dev)).reshape(-1)
    hits, idx, _ = _march(model, rays_d, t0_d, idx, network_config["gt_mode"], network_config["alpha"],
                          rendering_config["surface_threshold"], rendering_config["max_iterations"])
    still = torch.zeros(hi - lo, dtype=torch.bool, device=dev)
    still[idx] = True
    packed = torch.cat([t0_d, hits.to(torch.float64)[:, None], still.to(torch.float64)[:, None]], 1)     # one gather
    full = dp.gather_rows(packed, counts).cpu().numpy()
    t0[...] = full[:, :3]
    mask_rays[...] = full[:, 4] != 0
    hits_np = full[:, 3] != 0
    if hits_np.sum() == 0:
        raise ValueError(f"Ray tracing did not converge in {rendering_config['max_iterations']} iterations to any point at "
                         f"distance {rendering_config['surface_threshold']} or lower from surface.")
    return hits_np


def project_points_sharded(sampler, samples, gt_mode, alpha, num_steps, dp):
    """NDF projection (render_pc.Sampler.project) sharded by contiguous point ranges, gathered on every rank.
    samples: (P, 3) float64 CUDA tensor (the same on every rank).  Returns (samples, steps, grads, hessians)."""
    P = samples.shape[0]
    lo, hi = shard_range(P, dp.rank, dp.world)
    counts = [shard_range(P, r, dp.world)[1] - shard_range(P, r, dp.world)[0] for r in range(dp.world)]
    s, st, g, H = sampler.project(samples[lo:hi].contiguous(), gt_mode, alpha, num_steps)
    packed = torch.cat([s, st.to(torch.float64)[:, None], g.to(torch.float64), H.reshape(-1, 9).to(torch.float64)], 1)
    full = dp.gather_rows(packed, counts)
    return full[:, :3], full[:, 3].to(torch.float32), full[:, 4:7].to(torch.float32), full[:, 7:16].to(torch.float32).reshape(-1, 3, 3)
