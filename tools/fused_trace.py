"""Event timeline of CTA 0 of the fused tensor-core training kernel (epilogue warp 0 and the MMA warp): per sub-tile pair, how many
clocks each layer's epilogue works and waits in the forward sweep, the loss phase and the reverse sweep.
    python tools/fused_trace.py [out.txt]"""
import os
import sys
from collections import defaultdict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ALPHA, LR, W_S1, make_batches  # noqa: E402
from diffudf_b200 import SIREN, _lib  # noqa: E402
from diffudf_b200.train import FusedTrainer  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
batches = [tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x[0], n[0], d[0, :, 0])) for x, n, d in make_batches(2, 0)]
torch.manual_seed(123)
model = SIREN(3, 1, [256] * 8, w0=30).cuda()
tr = FusedTrainer(model, precision="tc16", fused=True)
tr.core.fused_flags = int(os.environ.get("DUDF_FUSED_FLAGS", "2"))
for i in range(4):
    tr.step("s1", *batches[i % 2], 9990, W_S1, ALPHA, LR)
buf = torch.zeros(3 * 8192, dtype=torch.int64, device="cuda")
L = _lib.lib()
torch.cuda.synchronize()
L.dudf_debug_set_trace(buf.data_ptr())
tr.step("s1", *batches[0], 9990, W_S1, ALPHA, LR)
torch.cuda.synchronize()
L.dudf_debug_set_trace(None)
raw = buf.cpu().numpy().astype("uint64")


def events(region):
    ev = []
    for v in raw[region * 8192:(region + 1) * 8192]:
        v = int(v)
        if v == 0:
            break
        ev.append((v & 0xFFFFFFFFFFFF, v >> 56, (v >> 48) & 0xFF))
    return ev


epi, mma = events(1), events(0)
t0 = epi[0][0]
pairs, cur = [], None
for t, tag, aux in epi:
    if tag == 14:
        cur = {"nch": aux, "start": t, "ev": []}
    elif tag == 15:
        cur["end"] = t
        pairs.append(cur)
        cur = None
    elif cur is not None:
        cur["ev"].append((t, tag, aux))
print(f"{len(pairs)} pairs traced on CTA 0; kernel span {epi[-1][0] - t0} clk", file=out)
for k, p in enumerate(pairs):
    tot = p["end"] - p["start"]
    acc = defaultdict(int)
    per_layer = defaultdict(lambda: defaultdict(int))
    last = {}
    prev_t = p["start"]
    marks = {}
    for t, tag, aux in p["ev"]:
        if tag in (42, 43):
            last[("fw", tag - 42)] = t
        elif tag in (10, 11):
            s = tag - 10
            per_layer[aux][f"f_wait{s}"] = t - last[("fw", s)]
            acc["fwd wait"] += t - last[("fw", s)]
            last[("f", s)] = t
        elif tag in (12, 13):
            s = tag - 12
            per_layer[aux][f"f_work{s}"] = t - last[("f", s)]
            acc["fwd work"] += t - last[("f", s)]
        elif tag in (40, 41):
            last[("bw", tag - 40)] = t
        elif tag in (30, 31):
            s = tag - 30
            per_layer[aux][f"b_wait{s}"] = t - last[("bw", s)]
            acc["bwd wait"] += t - last[("bw", s)]
            last[("b", s)] = t
        elif tag in (32, 33):
            s = tag - 32
            per_layer[aux][f"b_work{s}"] = t - last[("b", s)]
            acc["bwd work"] += t - last[("b", s)]
        elif tag in (20, 21, 22):
            marks[tag] = t
    acc["output dot"] = marks[21] - marks[20]
    acc["loss rows"] = marks[22] - marks[21]
    rest = tot - sum(acc.values())
    print(f"\npair {k} (jet-{p['nch']}): {tot} clk   " + "  ".join(f"{n} {v} ({100 * v / tot:.0f}%)" for n, v in acc.items()) + f"  other {rest}", file=out)
    print("  layer   f_wait0 f_work0 f_wait1 f_work1 | b_wait0 b_work0 b_wait1 b_work1", file=out)
    for l in sorted(per_layer):
        d = per_layer[l]
        print(f"  {l:5d}   " + " ".join(f"{d.get(n, 0):7d}" for n in ("f_wait0", "f_work0", "f_wait1", "f_work1")) + " | " +
              " ".join(f"{d.get(n, 0):7d}" for n in ("b_wait0", "b_work0", "b_wait1", "b_work1")), file=out)
# MMA warp: time from 'activation tile seen' to 'group issued' per sub-tile
seen = {}
dur = []
for t, tag, aux in mma:
    if tag == 1:
        seen[aux] = t
    elif tag == 2 and (aux - 2) in seen:
        dur.append(t - seen[aux - 2])
if dur:
    dur = np.array(dur)
    print(f"\nMMA warp: act_ready -> 32 MMAs + image copy issued: median {np.median(dur):.0f} clk, p90 {np.percentile(dur, 90):.0f}, max {dur.max()} ({len(dur)} sub-tile phases)", file=out)
