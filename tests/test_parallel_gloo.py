"""World-size-2 gloo test of the data-parallel decomposition: sharding each row group, scaling
every local term by 1/P_global and summing the per-rank gradients reproduces the single-process step; loss_s2
needs its statistics reduced BEFORE the backward.  The per-rank compute is the CPU oracle (test infrastructure)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffudf_b200.parallel import DataParallel, shard_batch
    from oracle import dudf_oracle as O
    params = O.init_params(n_hidden=2, seed=5)
    Ld = np.load(os.path.join(GOLDEN, "losses_init.npz"))
    x, n, d = Ld["x"][0], Ld["normals"][0], Ld["d"][0, :, 0]
    P = x.shape[0]
    n_on = int((d == 0).sum())
    n_far = (P - n_on) // 2
    dp = DataParallel(rows_global=P)
    xr, nr, dr, on_r = shard_batch(x, n, d, n_on, n_far, rank, world)
    res = {}
    for mode, w in (("s1", [1e4, 1e4, 1e4, 1e3]), ("s2", [1e5, 1e5])):
        stats = None
        if mode == "s2":
            f = O.siren_jet(params, xr, 0)["f"][dr == 0]
            st = torch.tensor([f.size, f.sum(), (f * f).sum()], dtype=torch.float64)
            dp.reduce_stats(st)
            stats = tuple(st.tolist())
        terms, grads = O.train_grads(params, xr, nr, dr, mode, w, 100.0, P_global=dp.global_rows(xr.shape[0]), s2_stats=stats)
        flat = torch.from_numpy(np.concatenate([np.concatenate([gW.reshape(-1), gb.reshape(-1)]) for gW, gb in grads]))
        dp.reduce_grads(flat)
        t = torch.tensor([float(v) for v in terms.values()], dtype=torch.float64)
        if mode == "s1":
            dp.reduce_terms(t)
        res[mode] = (flat.numpy(), t.numpy())
    if rank == 0:
        np.savez(os.path.join(out_dir, "dp.npz"), s1_g=res["s1"][0], s1_t=res["s1"][1], s2_g=res["s2"][0], s2_t=res["s2"][1])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_data_parallel_equals_single_process(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "dp.npz"))
    from oracle import dudf_oracle as O
    params = O.init_params(n_hidden=2, seed=5)
    Ld = np.load(os.path.join(GOLDEN, "losses_init.npz"))
    x, n, d = Ld["x"][0], Ld["normals"][0], Ld["d"][0, :, 0]
    for mode, w in (("s1", [1e4, 1e4, 1e4, 1e3]), ("s2", [1e5, 1e5])):
        terms, grads = O.train_grads(params, x, n, d, mode, w, 100.0)
        flat = np.concatenate([np.concatenate([gW.reshape(-1), gb.reshape(-1)]) for gW, gb in grads])
        assert np.max(np.abs(got[f"{mode}_g"] - flat)) <= 1e-9 * np.max(np.abs(flat))
        assert np.allclose(got[f"{mode}_t"], [float(v) for v in terms.values()], rtol=1e-10)


def _gather_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffudf_b200.parallel import DataParallel, gather_round_robin, grad_groups
    dp = DataParallel()
    ok = True
    for total, rounds in ((1003, 3), (1024, 4), (5, 4), (64, 1)):
        def fill(first, count, views):
            idx = torch.arange(first, first + count, dtype=torch.float32)
            views[0][:count] = idx
            views[1][:count] = torch.stack([idx, 2 * idx, -idx], 1)
        a, b = gather_round_robin(fill, total, dp, [(), (3,)], [torch.float32, torch.float32], "cpu", rounds)
        ref = torch.arange(total, dtype=torch.float32)
        ok = ok and a.shape == (total,) and b.shape == (total, 3) and bool(torch.equal(a, ref)) and bool(torch.equal(b[:, 1], 2 * ref))
    if rank == 0:
        np.savez(os.path.join(out_dir, "gather.npz"), ok=np.array(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_round_robin_gather(tmp_path):
    """grid queries at N > 1: round-robin blocks, one in-place all_gather_into_tensor per round (ragged totals included)"""
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert bool(np.load(os.path.join(str(tmp_path), "gather.npz"))["ok"])


def test_grad_groups_cover_the_flat_gradient():
    """layer groups of the overlapped gradient all-reduce: contiguous, disjoint, complete"""
    from diffudf_b200.parallel import grad_groups
    sizes = [(256 * 3, 256)] + [(256 * 256, 256)] * 7 + [(256, 1)]
    total = sum(a + b for a, b in sizes)
    for ng in (1, 2, 3, 7):
        groups = grad_groups(sizes, ng)
        assert groups[0][2] == 0 and groups[-1][3] == total and groups[0][0] == 1 and groups[-1][1] == 8
        for (lo0, hi0, a0, b0), (lo1, hi1, a1, b1) in zip(groups, groups[1:]):
            assert hi0 == lo1 and b0 == a1 and lo0 < hi0
