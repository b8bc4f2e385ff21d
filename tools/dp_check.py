"""Data-parallel correctness on real GPUs (run under torchrun, world_size >= 2):
the gradient all-reduced over ranks that each hold a [on|far|near] shard equals the single-GPU gradient of the
whole batch, and the loss-term shares sum to the single-GPU terms.  Prints DP_CHECK_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffudf_b200 import SIREN, synthetic  # noqa: E402
from diffudf_b200.parallel import DataParallel, shard_batch  # noqa: E402
from diffudf_b200.train import FusedTrainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = synthetic.make_shape(0)
sp, sn = shape.sample_surface(50000, np.random.default_rng(0))
x, n, d = synthetic.make_batch(shape, sp, sn, 6000, (0.333, 0.666), np.random.default_rng(7))
x, n, d = x[0], n[0], d[0, :, 0]
P, n_on = x.shape[0], 1998
n_far = (P - n_on) // 2
ok = True
# three gradient exchanges: "peer" = reduction fused into the Adam kernel over peer-mapped memory (default), "nccl2" = NCCL
# all-reduce of two layer groups under the weight-gradient GEMMs, "nccl1" = one blocking all-reduce
for exchange, env in (("peer", {"DUDF_DP_PEER": "1"}), ("nccl2", {"DUDF_DP_PEER": "0", "DUDF_DP_GROUPS": "2"}), ("nccl1", {"DUDF_DP_PEER": "0", "DUDF_DP_GROUPS": "1"})):
    os.environ.update(env)
    for prec, tol in (("fp32", 2e-4), ("tcx3", 5e-3), ("tc16", 2e-2)):
        for mode, w in (("s1", [1e4, 1e4, 1e4, 1e3]), ("s2", [1e5, 1e5])):
            torch.manual_seed(123)
            m = SIREN(3, 1, [256] * 8, w0=30).cuda()
            dp = DataParallel(rows_global=P)
            tr = FusedTrainer(m, dp=dp, precision=prec)
            if exchange == "peer" and tr.peer is None:
                if rank == 0:
                    print("peer-memory exchange unavailable on this box: skipped", flush=True)
                break
            if tr.peer is not None:
                tr.grad_sum = torch.zeros(tr.n, device="cuda")
            xr, nr, dr, on_r = shard_batch(x, n, d, n_on, n_far, rank, world)
            for _ in range(2):      # lr = 0: the second step sees the same weights and, on the tc16 path, takes the fused launch
                t = tr.step(mode, torch.from_numpy(xr).cuda(), torch.from_numpy(nr).cuda(), torch.from_numpy(dr).cuda(), on_r, w, 100.0, 0.0)
            if mode == "s1":
                dp.reduce_terms(t)
            g_dp = (tr.grad_sum if tr.peer is not None else tr.grad).clone()
            # every rank must hold the bit-identical reduced gradient (replicated Adam stays in lock-step)
            gl = [torch.empty_like(g_dp) for _ in range(world)]
            dist.all_gather(gl, g_dp)
            same_everywhere = all(bool(torch.equal(gl[0], g)) for g in gl)
            if rank == 0:
                torch.manual_seed(123)
                m1 = SIREN(3, 1, [256] * 8, w0=30).cuda()
                tr1 = FusedTrainer(m1, precision=prec)
                for _ in range(2):
                    t1 = tr1.step(mode, torch.from_numpy(x).cuda(), torch.from_numpy(n).cuda(), torch.from_numpy(d).cuda(), n_on, w, 100.0, 0.0)
                eg = float((g_dp - tr1.grad).abs().max() / tr1.grad.abs().max())
                et = float(((t - t1).abs() / t1.abs().clamp_min(1e-6)).max())
                print(f"{exchange} {prec} {mode}: grad err {eg:.2e} terms err {et:.2e} identical on all ranks {same_everywhere}", flush=True)
                ok = ok and eg < tol and et < 1e-3 and same_everywhere
            del tr
os.environ.pop("DUDF_DP_PEER", None)
os.environ.pop("DUDF_DP_GROUPS", None)
# grid query sharded by contiguous ranges of the flat index + all-gather == the single-GPU grid, bit for bit
from diffudf_b200.parallel import extract_fields_sharded  # noqa: E402
from diffudf_b200.render_mc import extract_fields  # noqa: E402
torch.manual_seed(123)
mg = SIREN(3, 1, [256] * 8, w0=30).cuda()
for prec in ("fp32", "tc16"):
    mg.precision = prec
    df, vecs = extract_fields_sharded(mg, 45, "tanh", 100.0, DataParallel())
    if rank == 0:
        df1, v1 = extract_fields(mg, None, 45, "tanh", torch.device("cuda", local), 100.0)
        same = bool(torch.equal(df, df1) and torch.equal(vecs, v1))
        print(f"{prec} sharded grid identical: {same}", flush=True)
        ok = ok and same
# sphere tracing and NDF projection sharded by contiguous ranges + all-gather == the single-GPU drivers, bit for bit
from diffudf_b200 import render_st  # noqa: E402
from diffudf_b200.parallel import project_points_sharded, propagate_rays_sharded  # noqa: E402
from diffudf_b200.render_pc import Sampler  # noqa: E402
rng = np.random.default_rng(3)
R = 5001
org = np.tile(np.array([[0.0, 0.0, 0.9]]), (R, 1))
dirs = rng.normal(size=(R, 3)) * 0.3 + np.array([[0.0, 0.0, -1.0]])
dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
ncfg, rcfg = dict(gt_mode="tanh", alpha=100.0), dict(surface_threshold=0.05, max_iterations=20)
mg.precision = "fp32"
try:
    t0a, ma = org.copy(), np.ones(R, bool)
    ha = propagate_rays_sharded(mg, dirs, t0a, ma, ncfg, rcfg, DataParallel())
    if rank == 0:
        t0b, mb = org.copy(), np.ones(R, bool)
        hb = render_st.propagate_rays(mg, dirs, t0b, mb, ncfg, rcfg, torch.device("cuda", local))
        same = bool(np.array_equal(ha, hb) and np.array_equal(t0a, t0b) and np.array_equal(ma, mb))
        print(f"sharded sphere tracing identical: {same} ({int(ha.sum())} hits of {R})", flush=True)
        ok = ok and same
except ValueError as exc:          # an untrained field may not be hit at all: both sides must agree on that too
    if rank == 0:
        print("sphere tracing: no hits on either side:", str(exc)[:60], flush=True)
smp = Sampler.__new__(Sampler)
smp.decoder, smp.device, smp.features = mg, torch.device("cuda", local), 3
pts = torch.from_numpy(rng.uniform(-1, 1, (7001, 3))).cuda()
pa = project_points_sharded(smp, pts, "tanh", 100.0, 3, DataParallel())
if rank == 0:
    pb = smp.project(pts, "tanh", 100.0, 3)
    same = all(torch.equal(torch.nan_to_num(a.to(b.dtype)), torch.nan_to_num(b)) for a, b in zip(pa, pb))
    print(f"sharded projection identical: {same}", flush=True)
    ok = ok and same
# the public epoch loop with config["dp"]: every rank iterates the same dataset and trains on its share of each row group; the
# per-epoch loss sums (all-reduced) equal the single-GPU loop's on the same batches
from diffudf_b200.train import train_model_tanh  # noqa: E402


class _DS:
    def __init__(self, batches, n_on):
        self.batches, self.batchesPerEpoch, self.samplesOnSurface = batches, len(batches), n_on

    def __iter__(self):
        for b in self.batches:
            yield tuple(torch.from_numpy(a) for a in b)


bs = [synthetic.make_batch(shape, sp, sn, 3000, (0.333, 0.666), np.random.default_rng(70 + i)) for i in range(2)]
cfg = dict(epochs=3, s1_epochs=2, warmup_epochs=1, warmup_lr=1e-5, lr_s1=1e-6, lr_s2=1e-7, loss_s1_weights=[1e4, 1e4, 1e4, 1e3],
           loss_s2_weights=[1e5, 1e5], alpha=100.0, precision="tcx3")
torch.manual_seed(123)
md = SIREN(3, 1, [256] * 8, w0=30).cuda()
ld, _, _ = train_model_tanh(_DS(bs, 999), md, torch.device("cuda", local), dict(cfg, dp=DataParallel()))
if rank == 0:
    torch.manual_seed(123)
    m1 = SIREN(3, 1, [256] * 8, w0=30).cuda()
    l1, _, _ = train_model_tanh(_DS(bs, 999), m1, torch.device("cuda", local), cfg)
    worst = max(abs(a - b) / max(abs(b), 1e-2) for k in l1 for a, b in zip(ld[k], l1[k]))
    print(f"epoch loop with dp: worst relative difference of the per-epoch loss terms {worst:.2e}", flush=True)
    ok = ok and worst < 5e-3
dist.barrier()
if rank == 0:
    print("DP_CHECK_OK" if ok else "DP_CHECK_FAILED", flush=True)
dist.destroy_process_group()
