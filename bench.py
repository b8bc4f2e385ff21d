#!/usr/bin/env python
"""Benchmark of the DUDF hot path (contract: see the task description / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): SIREN 3->256x8->1 training on a 200k-sample synthetic complex shape,
one step = loss_s1 (weights [1e4,1e4,1e4,1e3], alpha 100) forward jets + loss + reverse sweep + weight
gradients + Adam on a 29 970-row batch [9 990 on | 9 990 far | 9 990 near] per GPU (weak scaling), data
parallel with one all-reduce of the flat gradient.  Metric: train points/s (whole job).

`value` is measured in the CONFORMING arithmetic (--precision tcx3: split-precision tcgen05 forward, jets fp32-grade,
parameter gradients within north_star's 1e-3; tests/test_gpu_tcx3.py); the single-pass fp16 mode (tc16, 1e-3-class
jets, gradients 1e-2) is reported in `aux` only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F = {1: 1536 + 1 * 918016, 4: 1536 + 4 * 918016, 10: 1536 + 10 * 918016}   # algorithmic FLOP / point
W_S1 = [1e4, 1e4, 1e4, 1e3]
ALPHA = 100.0
LR = 1e-4
ROWS = 29970
# printed verbatim by BOTH arms (the driver compares the config objects of the two lines)
CONFIG = {"workload": "configs[1]: 200k-sample synthetic complex shape, SIREN 3->256x8->1, loss_s1 step (w=[1e4,1e4,1e4,1e3], alpha=100, "
                      "Adam lr 1e-4), 29 970 rows per GPU per step [9990 on|9990 far|9990 near], weak scaling",
          "rows": ROWS,
          "l2": "4 distinct batches cycled; the per-step working set of the GPU arm (operand images of the weight-gradient GEMM, ~1.3 GB "
                "written + read per step) exceeds the 126 MB L2"}


_JSON_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def make_batches(n_batches, seed, rows=30000):
    import numpy as np
    from diffudf_b200 import synthetic
    shape = synthetic.make_shape(0)
    surf_p, surf_n = shape.sample_surface(200000, np.random.default_rng(0))
    rng = np.random.default_rng(1000 + seed)
    return [synthetic.make_batch(shape, surf_p, surf_n, rows, (0.333, 0.666), rng) for _ in range(n_batches)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_run(steps, warmup, rows_cap=None):
    """The reference's CPU way of running the step (oracle/autograd_port.py) on the host cores."""
    import numpy as np
    import torch
    from oracle import autograd_port as AP
    from oracle import dudf_oracle as O
    torch.set_num_threads(os.cpu_count())
    params = AP.make_params(O.init_params(8, 256, 30.0, 123))
    opt = AP.make_optimizer(params, LR)
    x, n, d = make_batches(1, 0)[0]
    sample = "full 29 970-row batch per step"
    if rows_cap is not None and rows_cap < x.shape[1]:
        k = rows_cap // 3
        sel = np.concatenate([np.arange(k), 9990 + np.arange(k), 19980 + np.arange(k)])
        x, n, d = x[:, sel], n[:, sel], d[:, sel]
        sample = f"{3 * k}-row subsample [on|far|near] per step (linear in rows)"
    x, n, d = torch.from_numpy(x), torch.from_numpy(n), torch.from_numpy(d)
    for _ in range(warmup):
        AP.train_step(params, opt, x, n, d, "s1", W_S1, ALPHA)
    t0 = time.perf_counter()
    for _ in range(steps):
        AP.train_step(params, opt, x, n, d, "s1", W_S1, ALPHA)
    dt = time.perf_counter() - t0
    return x.shape[1] * steps / dt, dt / steps * 1e3, torch.get_num_threads(), sample


def run_reference(args, rank, world):
    if rank != 0:
        return
    # bound the run: ~2.3 s per full step on 8 cores; keep the whole arm within a few minutes
    t_probe = time.perf_counter()
    v1, ms1, cores, _ = cpu_port_run(1, 1, rows_cap=2997)
    est_full = ms1 * 10 / 1e3 * (args.steps + args.warmup)
    cap = None if est_full < 200 else max(2997, int(29970 * 200 / est_full) // 3 * 3)
    value, ms, cores, sample = cpu_port_run(args.steps, args.warmup, cap)
    line = {"impl": "reference", "metric": "train points/s", "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": CONFIG, "device": "host CPU",
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "probe_s": round(time.perf_counter() - t_probe, 1)}
    emit(line)


def aux_device_sampler(dev, trainer):
    """batches drawn on the device from a 200k-point oriented cloud (diffudf_b200.dataset.PointCloud,
    src/dataset.py:80-131 semantics) — the sampler alone, and sampler + training step back to back through the public loop."""
    import numpy as np
    import torch
    from diffudf_b200 import synthetic
    from diffudf_b200.dataset import PointCloud
    shape = synthetic.make_shape(0)
    pts, nrm = shape.sample_surface(200000, np.random.default_rng(0))
    ds = PointCloud(pts.astype(np.float32), nrm.astype(np.float32), 30000, [0.333, 0.666], 20, dev, seed=5)
    out = {}
    # the sampler alone: draws issued back to back in ONE stream (prefetch off), device time per batch
    ds.prefetch = False
    for _ in ds:
        break
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    nb = 0
    for _ in ds:
        nb += 1
    t.record()
    torch.cuda.synchronize()
    ds.prefetch = True
    ms = s.elapsed_time(t) / nb
    rows = ds.samplesOnSurface + ds.samplesFarSurface
    out["sampler_pc_200k_ms_per_batch"] = ms
    out["sampler_pc_200k_rows_per_s"] = rows / (ms * 1e-3)
    out["sampler_pc_200k_equivalent_pair_distances_per_s"] = (ds.samplesFarSurface // 2) * 200000 / (ms * 1e-3)
    # the far rows' distances alone: the prebuilt index against the tiled scan of the whole cloud (DUDF_NN_SCAN=1)
    from diffudf_b200.dataset import shortestDistance
    far = next(iter(ds))[0][0, ds.samplesOnSurface:ds.samplesOnSurface + ds.samplesFarSurface // 2].contiguous()
    for key, target, env in (("index", ds.index, None), ("scan", ds.surface_pc, "1")):
        if env:
            os.environ["DUDF_NN_SCAN"] = env
        shortestDistance(far, target)
        s.record()
        for _ in range(10):
            shortestDistance(far, target)
        t.record()
        torch.cuda.synchronize()
        os.environ.pop("DUDF_NN_SCAN", None)
        out[f"nearest_distance_{key}_9990x200k_ms"] = s.elapsed_time(t) / 10
    for x, n, d in ds:      # warm the step on this batch shape
        trainer.step("s1", x[0], n[0], d[0, :, 0], ds.samplesOnSurface, W_S1, ALPHA, LR)
        break
    s.record()
    nb = 0
    for x, n, d in ds:
        trainer.step("s1", x[0], n[0], d[0, :, 0], ds.samplesOnSurface, W_S1, ALPHA, LR)
        nb += 1
    t.record()
    torch.cuda.synchronize()
    out["train_with_device_sampler_points_per_s"] = rows * nb / (s.elapsed_time(t) * 1e-3)
    return out


def aux_other_losses(dev, batches, n_on, precision="tcx3"):
    """The other two loss configurations of the training loop on the same batches: loss_s2 (value-only
    forward, batch statistics, reverse sweep: train.py:174-191) and loss_siren (train.py:23-143), tensor-core path."""
    import torch
    from diffudf_b200 import SIREN
    from diffudf_b200.train import FusedTrainer
    out = {}
    # "_graph": the same step captured once as a CUDA graph and replayed (FusedTrainer(graph=True): learning rate and Adam's step
    # count on the device) — loss_s2's dozen small kernels take ~0.1 ms of device time, less than the host needs to launch them;
    # "loss_s1_graph" is the headline step through the same mechanism
    for mode, w, key, graph in (("s2", [1e5, 1e5], "loss_s2", False), ("s2", [1e5, 1e5], "loss_s2_graph", True),
                                ("siren", [3e3, 1e2, 1e2, 5e1], "loss_siren", False), ("s1", W_S1, "loss_s1_graph", True)):
        torch.manual_seed(123)
        tr = FusedTrainer(SIREN(3, 1, [256] * 8, w0=30).to(dev), precision=precision, graph=graph)
        for i in range(4):
            tr.step(mode, *batches[i % len(batches)], n_on, w, ALPHA, LR)
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(20):
            tr.step(mode, *batches[i % len(batches)], n_on, w, ALPHA, LR)
        t.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(t) / 20
        out[f"{key}_step_ms"] = ms
        out[f"{key}_train_points_per_s"] = batches[0][0].shape[0] / (ms * 1e-3)
    return out


def aux_full_size_queries(dev):
    """BASELINE configs 3-5 at full size on one GPU (split-precision tensor-core path = the conforming arithmetic, trained fixture
    weights when present): 512^3 grid for marching cubes (also in the single-pass fp16 mode), evaluate() with host buffers,
    1024^2 sphere tracing, 2 M-point NDF projection."""
    import numpy as np
    import torch
    from diffudf_b200 import SIREN, evaluate, render_st
    from diffudf_b200.render_mc import extract_fields
    from diffudf_b200.render_pc import Sampler
    out = {}
    m = SIREN(3, 1, [256] * 8, w0=30).to(dev)
    wpath = os.path.join(ROOT, "tests", "golden", "weights_trained.npz")
    if os.path.exists(wpath):
        z = np.load(wpath)
        m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(z[f"{c}{i}"]) for i in range(9) for k, c in (("weight", "W"), ("bias", "b"))})
        m.to(dev)
    m.precision = "tcx3"

    def timed(fn, reps=1):
        fn()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            r = fn()
        t.record()
        torch.cuda.synchronize()
        return s.elapsed_time(t) * 1e-3 / reps, r

    for prec in ("tc16", "tcx3"):
        m.precision = prec
        sec, _ = timed(lambda: extract_fields(m, None, 512, "tanh", dev, ALPHA))
        out[f"grid512_extract_fields_{prec}_queries_per_s"] = 512 ** 3 / sec
        out[f"grid512_extract_fields_{prec}_s"] = sec
    # CAP-UDF marching cubes on those 512^3 fields (src/render_mc.py:201-256; classify + scan + emit, HBM-bound).  Algorithmic
    # bytes: 4 B per grid point (distances; gradients are read for near-surface cells only), 2 B per cell (case byte written and
    # read back), 72 B per triangle
    from diffudf_b200.render_mc import cap_triangles
    u512, g512 = extract_fields(m, None, 512, "tanh", dev, ALPHA)
    sec, tri = timed(lambda: cap_triangles(u512, g512, 512), reps=3)
    out["cap_mc_512_ms"] = sec * 1e3
    out["cap_mc_512_triangles"] = int(tri.shape[0])
    out["cap_mc_512_algorithmic_gb_per_s"] = (4 * 512 ** 3 + 2 * 511 ** 3 + 72 * int(tri.shape[0])) / sec / 1e9
    del u512, g512, tri
    # evaluate(): host numpy in, float64 host arrays out (value + gradient), 2 M points
    xs = np.random.default_rng(0).uniform(-1, 1, (2_000_000, 3)).astype(np.float32)
    grads = np.zeros((xs.shape[0], 3))
    evaluate(m, xs, device=dev, gradients=grads, max_batch=1 << 20)          # warm-up: pinned staging, page faults of the result arrays
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        evaluate(m, xs, device=dev, gradients=grads, max_batch=1 << 20)
        times.append(time.perf_counter() - t0)
    out["evaluate_host_2M_value_grad_queries_per_s"] = xs.shape[0] / sorted(times)[1]
    # NDF projection (config 5): 2 M seeds, 3 steps (2 gradient queries + 1 Hessian query per point)
    smp = Sampler(decoder=m, device=dev)
    seeds = torch.from_numpy(xs.astype(np.float64)).to(dev)
    sec, _ = timed(lambda: smp.project(seeds, "tanh", ALPHA, 3))
    out["project_2M_x3steps_points_per_s"] = xs.shape[0] / sec
    for p in m.parameters():
        p.requires_grad_(True)
    # sphere tracing 1024 x 1024 (config 4) + eigen-normals and mean curvature at the hits
    R = 1024
    cam = np.array([0.8939, 0.7, 2.86]) * 0.45
    u, v = np.meshgrid(np.linspace(-0.6, 0.6, R), np.linspace(-0.6, 0.6, R))
    d = np.stack([u.ravel(), v.ravel(), -np.ones(R * R)], 1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    fwd = -cam / np.linalg.norm(cam)
    right = np.cross(fwd, [0, 1.0, 0]); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    rays = d[:, :1] * right + d[:, 1:2] * up - d[:, 2:3] * fwd
    start = np.tile(cam, (R * R, 1)) + rays * 0.35
    rays_d = torch.from_numpy(rays).to(dev)
    t0_d = torch.from_numpy(start).to(dev)
    idx = torch.arange(R * R, device=dev)
    render_st._march(m, rays_d, t0_d.clone(), idx, "tanh", ALPHA, 0.004, 100)        # untimed: allocates the driver workspace
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hits, _, nq = render_st._march(m, rays_d, t0_d.clone(), idx, "tanh", ALPHA, 0.004, 100)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    out["sphere_trace_1024_rays_per_s"] = R * R / sec
    out["sphere_trace_1024_value_queries_per_s"] = nq / sec
    out["sphere_trace_1024_hit_fraction"] = float(hits.float().mean())
    # the whole frame of BASELINE configs[3] through the reference's call (generate_st.py:127): numpy rays in, marching, hit
    # attributes with the third-order jet (mean curvature), percentile colour map, Blinn-Phong shading, (1024, 1024, 3) image out
    cfg = {"surface_threshold": 0.004, "max_iterations": 100, "gd_steps": 0, "height": R, "width": R, "light_position": [1, 2.38206, 10],
           "camera_position": cam.tolist(), "shininess": -1, "plot_curvatures": "mean", "curv_low_bound": 5, "curv_high_bound": 95,
           "reflection_method": "blinn-phong", "alpha1": 0.2, "alpha2": 0.2}
    net_cfg = {"gt_mode": "tanh", "alpha": ALPHA}
    for rep in range(2):
        t0_np, mask_np = start.copy(), np.ones(R * R, dtype=bool)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        img = render_st.create_projectional_image(m, rays, t0_np, mask_np, net_cfg, cfg, dev)
        sec = time.perf_counter() - t0
    out["st_image_1024_mean_curvature_s"] = sec
    out["st_image_1024_mean_curvature_rays_per_s"] = R * R / sec
    out["st_image_1024_hit_pixels"] = int((img.reshape(-1, 3) != 1.0).any(axis=1).sum())
    return out


def aux_reference_eager(dev, host_batch):
    """SURVEY 8d last row / VERDICT r1 item 3: the reference's own operation sequence (oracle/autograd_port.py: nn.Linear-style
    matmuls + torch.sin, autograd double-backward, torch.linalg.eigh, backward, torch.optim.Adam) as STOCK eager PyTorch on this
    B200, fp32 and with TF32 matmuls allowed — "the reference on the same box"."""
    import numpy as np
    import torch
    from oracle import autograd_port as AP
    from oracle import dudf_oracle as O
    out = {}
    x, n, d = (t.to(dev).unsqueeze(0) for t in host_batch)
    d = d.unsqueeze(-1)
    xq = np.random.default_rng(0).uniform(-1, 1, (1 << 17, 3)).astype(np.float32)
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        for tag, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            params = AP.make_params(O.init_params(8, 256, 30.0, 123), device=str(dev))
            opt = AP.make_optimizer(params, LR)
            for _ in range(2):
                AP.train_step(params, opt, x, n, d, "s1", W_S1, ALPHA)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                AP.train_step(params, opt, x, n, d, "s1", W_S1, ALPHA)
            torch.cuda.synchronize()
            out[f"reference_eager_b200_{tag}_train_points_per_s"] = x.shape[1] * 5 / (time.perf_counter() - t0)
            for key, wh in (("value_grad", False), ("value_grad_hess", True)):
                AP.evaluate(params, xq[:8192], True, wh)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                AP.evaluate(params, xq, True, wh)
                torch.cuda.synchronize()
                out[f"reference_eager_b200_{tag}_{key}_queries_per_s"] = xq.shape[0] / (time.perf_counter() - t0)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return out


def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from diffudf_b200 import SIREN, _lib
    from diffudf_b200.parallel import DataParallel, shard_batch
    from diffudf_b200.train import FusedTrainer
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dp = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dp = DataParallel()
    torch.manual_seed(123)
    model = SIREN(3, 1, [256] * 8, w0=30).to(dev)
    trainer = FusedTrainer(model, dp=dp, precision=args.precision)
    NB = 4
    host = []
    for x, n, d in make_batches(NB, rank):
        host.append(tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (x[0], n[0], d[0, :, 0])))
    resident = [tuple(t.to(dev) for t in b) for b in host]
    P = host[0][0].shape[0]
    assert P == ROWS
    n_on = 9990
    L = _lib.lib()

    def step_resident(i, tr=None):
        x, n, d = resident[i % NB]
        return (tr or trainer).step("s1", x, n, d, n_on, W_S1, ALPHA, LR)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(nsteps, fn=step_resident):
        """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over ranks -> ms"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nsteps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    # ---- device-resident timing (the headline `value`) ----
    clocks = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if clocks:
        clocks.start()
    l0 = L.dudf_launch_count()
    ms_total = timed_region(args.steps)
    launches = L.dudf_launch_count() - l0
    # ---- end to end: pinned host batch -> device, step, loss terms -> host, every step ----
    # the public loop (diffudf_b200.train): BatchFeeder copies batch i+1 from pinned host memory on a side stream while
    # step i computes; the 4 loss terms of every step are copied into pinned host memory; one sync at the end
    from diffudf_b200.train import BatchFeeder
    feeder = BatchFeeder(dev)
    host_terms = torch.zeros(args.steps, 4, dtype=torch.float64).pin_memory()

    def e2e_loop(nsteps, sink):
        for i, (x, n, d) in enumerate(feeder.feed(host[j % NB] for j in range(nsteps))):
            t = trainer.step("s1", x, n, d, n_on, W_S1, ALPHA, LR)
            if sink is not None:
                sink[i].copy_(t, non_blocking=True)

    e2e_loop(2, None)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_loop(args.steps, host_terms)
    e3.record()
    barrier()
    last = host_terms[-1]
    ms2 = torch.tensor([e2.elapsed_time(e3)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms2.item())
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    assert bool(torch.isfinite(last).all()), "non-finite loss terms"
    aux = {}
    # ---- timing hygiene (SURVEY 8d "Timing method"): 5 more repeats of the K-step region, and a >= 2 s back-to-back loop ----
    reps = sorted(timed_region(args.steps) / args.steps for _ in range(5))
    aux["repeat_ms_per_step_min"], aux["repeat_ms_per_step_median"] = reps[0], reps[2]
    per = max(reps[2], 1e-3)
    n_sus = int(2200.0 / per) + 1
    ms_sus = timed_region(n_sus)
    aux["sustained_ms_per_step"] = ms_sus / n_sus
    aux["sustained_steps"], aux["sustained_seconds"] = n_sus, ms_sus * 1e-3
    aux["sustained_train_points_per_s"] = world * P * n_sus / (ms_sus * 1e-3)
    clk = clocks.stop() if clocks else None
    # ---- strong scaling: the SAME global 29 970-row batch split over the ranks ----
    if world > 1:
        dps = DataParallel(rows_global=P)
        torch.manual_seed(123)
        tr_s = FusedTrainer(SIREN(3, 1, [256] * 8, w0=30).to(dev), dp=dps, precision=args.precision)
        shards = []
        for b in make_batches(NB, 0):       # every rank builds the same global batches and keeps its share of each row group
            xs, ns, ds, non = shard_batch(b[0][0], b[1][0], b[2][0, :, 0], n_on, 9990, rank, world)
            shards.append((torch.from_numpy(np.ascontiguousarray(xs)).to(dev), torch.from_numpy(np.ascontiguousarray(ns)).to(dev),
                           torch.from_numpy(np.ascontiguousarray(ds)).to(dev), non))

        def step_strong(i):
            xs, ns, ds, non = shards[i % NB]
            return tr_s.step("s1", xs, ns, ds, non, W_S1, ALPHA, LR)
        for i in range(4):
            step_strong(i)
        ms_st = timed_region(max(args.steps, 20), step_strong)
        aux["strong_scaling_ms_per_step"] = ms_st / max(args.steps, 20)
        aux["strong_scaling_points_per_s"] = P * max(args.steps, 20) / (ms_st * 1e-3)
        del tr_s, shards

    # ---- per-kernel timing of one step (CUDA events on the launching stream) for the roofline object ----
    prof = {}
    if True:      # every rank runs these steps (they contain the gradient all-reduce); rank 0's events are reported
        eng = model._engine_synced()
        names = ["jet_forward_multi", "loss", "jet_backward_multi", "jet_wgrad", "train_step_fused"]
        orig = {k: getattr(eng, k) for k in names}
        events = []

        def timed(name, fn, *a, **kw):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **kw)
            t.record()
            events.append((name, s, t))
            return r

        def wrap(name, fn):
            return lambda *a, **kw: timed(name, fn, *a, **kw)

        def fused_split(mode, segs, P_global, w, alpha, terms, amax_prev, amax_next, scratch, A, Zb, ld, gW, gb, flags=3):
            # the same two launches as the single C call, timed separately: fused forward+loss+reverse sweep, then the weight gradients
            timed("fused_fwd_loss_bwd", orig["train_step_fused"], mode, segs, P_global, w, alpha, terms, amax_prev, amax_next, scratch, A, Zb, ld,
                  gW, gb, flags | 4)
            timed("jet_wgrad", orig["jet_wgrad"], Zb, A, ld, ld, gW, "tc16", seed_absmax=amax_prev)
        for k in names:
            setattr(eng, k, fused_split if k == "train_step_fused" else wrap(k, orig[k]))
        nrep = 5
        for i in range(nrep):
            step_resident(i)
        torch.cuda.synchronize()
        for k in names:
            setattr(eng, k, orig[k])
        for tag, s, t in events:
            prof[tag] = prof.get(tag, 0.0) + s.elapsed_time(t) / nrep
    # ---- field queries (secondary metric: UDF+grad queries/s on a dense grid) ----
    if world > 1:
        aux["gradient_exchange"] = ("peer memory: reduction fused into the Adam kernel (dudf_adam_step_peers) behind a symmetric-memory barrier"
                                    if trainer.peer is not None else
                                    f"NCCL all-reduce in {len(trainer.groups) if trainer.groups else 1} layer group(s) under the weight-gradient GEMMs")
    if world > 1 and not args.light:
        # BASELINE's second metric at N GPUs: the 512^3 grid sharded by contiguous slabs of the flat index (no data-path
        # collective), then gathered (all_gather_into_tensor straight into the output); max over ranks, like the headline
        from diffudf_b200.parallel import extract_fields_sharded, shard_range
        from diffudf_b200.render_mc import extract_fields
        model.precision = "tcx3"
        Ng = 512
        lo, hi = shard_range(Ng ** 3, rank, world)
        extract_fields(model, None, 64, "tanh", dev, ALPHA)
        extract_fields_sharded(model, 64, "tanh", ALPHA, dp)        # untimed: NCCL's first all-gather sets up its channels (~0.2 s once)
        barrier()
        q0, q1, q2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        q0.record()
        df, vecs = extract_fields(model, None, Ng, "tanh", dev, ALPHA, first=lo, count=hi - lo)
        q1.record()
        del df, vecs
        df, vecs = extract_fields_sharded(model, Ng, "tanh", ALPHA, dp)
        q2.record()
        barrier()
        del df, vecs
        tq = torch.tensor([q0.elapsed_time(q1), q1.elapsed_time(q2)], device=dev, dtype=torch.float64)
        dist.all_reduce(tq, op=dist.ReduceOp.MAX)
        aux["grid512_tcx3_sharded_compute_queries_per_s"] = Ng ** 3 / (float(tq[0]) * 1e-3)
        aux["grid512_tcx3_sharded_compute_plus_gather_queries_per_s"] = Ng ** 3 / (float(tq[1]) * 1e-3)
    if rank == 0 and not args.light:
        eng = model._engine_synced()
        for prec, N in (("tcx3", 256), ("tc16", 256), ("fp32", 128)):
            cnt = N ** 3
            df = torch.empty(cnt, device=dev)
            vecs = torch.empty(cnt, 3, device=dev)
            eng.query_grid(N, 0, cnt, prec, 3, ALPHA, out=(df, vecs))
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(3):
                eng.query_grid(N, 0, cnt, prec, 3, ALPHA, out=(df, vecs))
            t.record()
            torch.cuda.synchronize()
            q = 3 * cnt / (s.elapsed_time(t) * 1e-3)
            aux[f"grid{N}_{prec}_queries_per_s"] = q
            aux[f"grid{N}_{prec}_tflops"] = q * F[4] / 1e12
        del df, vecs
        try:
            if args.precision != "tc16":       # the single-pass fp16 step (1e-3-class jets, gradients 1e-2): NOT the conforming arithmetic
                torch.manual_seed(123)
                tr16 = FusedTrainer(SIREN(3, 1, [256] * 8, w0=30).to(dev), precision="tc16")
                for i in range(4):
                    step_resident(i, tr16)
                s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for i in range(20):
                    step_resident(i, tr16)
                t.record()
                torch.cuda.synchronize()
                aux["tc16_single_pass_fused_step_ms"] = s.elapsed_time(t) / 20
                aux["tc16_single_pass_train_points_per_s"] = P * 20 / (s.elapsed_time(t) * 1e-3)
                del tr16
            aux.update(aux_reference_eager(dev, host[0]))
            aux.update(aux_full_size_queries(dev))
            aux.update(aux_other_losses(dev, resident, n_on, args.precision))
            aux.update(aux_device_sampler(dev, FusedTrainer(SIREN(3, 1, [256] * 8, w0=30).to(dev), precision=args.precision)))
        except Exception as exc:          # the secondary numbers must never cost the headline line
            aux["aux_error"] = repr(exc)[:300]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_burst = peaks.get("bf16_tflops", 1590.0)
    peak_sus = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = ("measured (MEASURED_PEAKS.json: bf16_tflops burst for the kernel timed alone, bf16_tflops_sustained for the >= 2 s loop)"
                if peaks else "fallback (B200_PROFILING.md: 1.59 PFLOP/s burst, 1.4 sustained)")
    third = n_on * F[10] + (P - n_on) * F[4]        # forward = reverse sweep = weight gradient in algorithmic FLOPs
    flops = {"jet_forward_multi": third, "jet_backward_multi": third, "jet_wgrad": third, "fused_fwd_loss_bwd": 2 * third}
    traffic = {}
    try:        # DRAM bytes per launch from the committed ncu --set full capture of the same kernels (tools/ncu_summary.py)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r3_ncu_traffic.json")))
    except Exception:
        pass
    dom = max((k for k in prof if k in flops), key=lambda k: prof[k]) if prof else None
    roof = None
    if dom:
        ach = flops[dom] / (prof[dom] * 1e-3) / 1e12
        step_gflop = third * 3 / 1e9
        ms_step = ms_total / args.steps
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak_burst, "unit": "TFLOP/s", "frac": ach / peak_burst,
                "frac_of_sustained_peak": ach / peak_sus,
                "traffic": traffic.get(dom), "traffic_source": traffic.get("source") if dom in traffic else None,
                "peak_source": peak_src, "ms": prof[dom],
                "step_share": prof[dom] / max(sum(prof.values()), 1e-9),
                "kernel_ms": {k: round(v, 4) for k, v in prof.items()},
                "kernel_tflops": {k: round(flops[k] / (prof[k] * 1e-3) / 1e12, 1) for k in prof if k in flops},
                "step_algorithmic_gflop": step_gflop,
                "whole_step": {"burst_region_tflops": step_gflop / ms_step, "burst_region_frac": step_gflop / ms_step / peak_burst,
                               "sustained_loop_tflops": step_gflop / aux["sustained_ms_per_step"],
                               "sustained_loop_frac": step_gflop / aux["sustained_ms_per_step"] / peak_sus},
                "note": "algorithmic FLOPs only: the split forward executes 3 MMAs per product (5/3 of the step's algorithmic MMA work)"
                        if args.precision == "tcx3" else None}
    cpu = None
    if world == 1:
        v, msc, cores, sample = cpu_port_run(3, 1)
        cpu = {"value": v, "unit": "points/s", "cores": cores, "kind": "port", "ms_per_step": msc,
               "sample": "3 timed loss_s1+backward+Adam steps (after 1 warm-up) of the " + sample + ", oracle/autograd_port.py"}
    value = world * P * args.steps / (ms_total * 1e-3)
    dtype = {"tcx3": "f16 hi+lo split operands (fp32-grade) / f32 accumulate", "tc16": "f16 operands / f32 accumulate", "fp32": "f32"}[args.precision]
    line = {"metric": "train points/s", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": CONFIG,
            "details": {"rows_per_gpu": P, "global_rows": world * P, "parallelism": f"dp{world}",
                        "precision": {"tcx3": "split-precision tcgen05 step (tcx3): forward 3 MMAs per product, reverse sweep / weight gradient "
                                              "single-pass; conforming to north_star's tolerances (tests/test_gpu_tcx3.py)",
                                      "tc16": "tcgen05 single-pass fp16-operand step (tc16); 1e-3-class, NOT conforming on derivatives",
                                      "fp32": "fp32 CUDA-core step"}[args.precision]},
            "e2e": {"value": world * P * args.steps / (ms_e2e * 1e-3), "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32,
                    "ms_per_step": ms_e2e / args.steps,
                    "mode": "diffudf_b200.train.BatchFeeder: batch i+1 copied from pinned host memory on a side stream during step i; "
                            "the 4 loss terms of every step copied to pinned host memory; one synchronisation at the end"},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "aux": aux}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tcx3", choices=["tcx3", "tc16", "fp32"],
                    help="arithmetic of the training step: split-precision tcgen05 (default, conforming), single-pass fp16 tcgen05, "
                         "or fp32 CUDA cores")
    ap.add_argument("--light", action="store_true", help="headline, e2e, sustained loop and roofline only (skips the secondary `aux` "
                                                         "measurements: full-size queries, other losses, eager reference, sampler)")
    args = ap.parse_args()
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    # rank 0 prints ONE JSON line on stdout: everything else that writes to fd 1 (NCCL's version banner comes from C) goes to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
