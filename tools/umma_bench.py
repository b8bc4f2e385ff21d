"""Clocks per tcgen05.mma (K = 16, M = 128) for the operand placements / shapes the chain kernels could use."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from diffudf_b200 import _lib  # noqa: E402

torch.cuda.init()
L = _lib.lib()
names = ["A smem K-major, B smem MN-major, N=128 (current)", "A smem K-major, B smem K-major, N=128",
         "A smem K-major, B smem MN-major, N=256", "A smem K-major, B smem K-major, N=256",
         "A TMEM, B smem K-major, N=256", "A TMEM, B smem K-major, N=128",
         "as current, commit every 4 MMAs", "as current, commit every 8 MMAs", "as current, commit every 16 MMAs",
         "as current, alternate 2 accumulators", "as current, alternate 4 accumulators", "as current, warp-uniform issue + descriptor adds"]
for ctas in (148,):
    for v, name in enumerate(names):
        out = ctypes.c_float(0)
        _lib.check(L.dudf_bench_umma(v, ctas, 512, ctypes.byref(out)), "dudf_bench_umma")
        n = 256 if "N=256" in name else 128
        ideal = 2 * 128 * n * 16 / 8192
        print(f"ctas={ctas:4d} {name:52s}: {out.value:7.1f} clk/MMA  (ideal {ideal:.0f} at 8192 flop/clk/SM -> {100 * ideal / out.value:5.1f} %)", flush=True)
