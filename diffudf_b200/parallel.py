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


def extract_fields_sharded(model, N, gt_mode, alpha, dp):
    """Grid query sharded by contiguous ranges of the flat index (slabs of the slowest axis when N
    divides evenly), gathered on every rank.  Returns df (N,N,N), vecs (N,N,N,3)."""
    from .render_mc import extract_fields
    total = N ** 3
    lo, hi = shard_range(total, dp.rank, dp.world)
    df, vecs = extract_fields(model, None, N, gt_mode, None, alpha, first=lo, count=hi - lo)
    counts = [shard_range(total, r, dp.world)[1] - shard_range(total, r, dp.world)[0] for r in range(dp.world)]
    df = dp.gather_rows(df, counts)
    vecs = dp.gather_rows(vecs, counts)
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
