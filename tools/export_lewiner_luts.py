"""Export Lewiner's marching-cubes look-up tables (Lewiner, Lopes, Vieira, Tavares, "Efficient implementation of Marching
Cubes' cases with topological guarantees", JGT 8(2) 2003 — published with the paper as LookUpTable.h and shipped by the
reference, base64-encoded, in src/marching_cubes/_marching_cubes_lewiner_luts.py) into the binary blob the C++ MeshUDF mesher
loads as DATA: diffudf_b200/data/lewiner_luts.bin.  The tables are constants of the published algorithm; nothing else of the
reference is copied.  Needs /root/reference (build container); the blob is committed.

blob := "DUDFLUT1" u32 n_tables { char name[16]; u32 ndim; u32 shape[3]; u32 nbytes; i8 data[nbytes] (padded to 4) }*
"""
import base64
import importlib.util
import os
import struct
import sys

import numpy as np

SRC = "/root/reference/src/marching_cubes/_marching_cubes_lewiner_luts.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffudf_b200", "data", "lewiner_luts.bin")

spec = importlib.util.spec_from_file_location("ref_luts", SRC)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
names = [n for n in dir(mod) if n.isupper() and isinstance(getattr(mod, n), tuple) and len(getattr(mod, n)) == 2]
# the reference keeps the edge -> corner offsets in its Python wrapper (_marching_cubes_lewiner.py:148-150): cube geometry,
# restated here from the edge numbering of the paper (edges 0-3 bottom face, 4-7 top face, 8-11 vertical)
corner = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]          # (x, y, z) of v0..v7
edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
tables = []
for axis, nm in enumerate(("EDGESRELX", "EDGESRELY", "EDGESRELZ")):
    tables.append((nm, np.array([[corner[a][axis], corner[b][axis]] for a, b in edges], np.int8)))
for n in sorted(names):
    shape, text = getattr(mod, n)
    ar = np.frombuffer(base64.decodebytes(text.encode("utf-8")), dtype=np.int8).reshape(shape)
    tables.append((n, ar))
with open(OUT, "wb") as fh:
    fh.write(b"DUDFLUT1")
    fh.write(struct.pack("<I", len(tables)))
    for n, ar in tables:
        shape = list(ar.shape) + [1] * (3 - ar.ndim)
        raw = np.ascontiguousarray(ar).tobytes()
        fh.write(n.encode().ljust(16, b"\0"))
        fh.write(struct.pack("<IIIII", ar.ndim, *shape, len(raw)))
        fh.write(raw + b"\0" * (-len(raw) % 4))
print(f"{len(tables)} tables -> {OUT} ({os.path.getsize(OUT)} bytes)")
