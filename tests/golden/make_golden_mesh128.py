"""Golden MeshUDF mesh of the trained fixture network at 128^3 (SURVEY 8d config 3 sub-problem): the oracle's fp32 fields
(oracle/dudf_oracle.extract_fields, pinned to the unmodified reference by fields_trained.npz) meshed by the reference's own
marching cubes (oracle/_ref, built from /root/reference by oracle/build_ref_mc.py) exactly as extract_mesh_MESHUDF calls it
(src/render_mc.py:127-133).  The fp64-numpy field evaluation takes ~8 minutes on 16 cores, which is why the GPU test reads
this fixture instead of recomputing it.

    python tests/golden/make_golden_mesh128.py        # writes tests/golden/mesh128_trained.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import build_ref_mc, dudf_oracle as O  # noqa: E402

N = 128
params = O.load_params(os.path.join(HERE, "weights_trained.npz"))
df, vecs = O.extract_fields(params, N, "tanh", 100.0, chunk=1 << 15)
assert build_ref_mc.built() or build_ref_mc.build()
mc = build_ref_mc.load()
d = np.array(df, np.float32)
d[d < 0] = 0
v, f, _, _ = mc(d, np.ascontiguousarray(vecs, np.float32), spacing=[2.0 / (N - 1)] * 3, avg_thresh=1.05, max_thresh=1.75)
v = v - 1
print(v.shape, f.shape)
np.savez_compressed(os.path.join(HERE, "mesh128_trained.npz"), N=N, verts=v.astype(np.float32), faces=f.astype(np.int32),
                    df_sample_idx=np.arange(0, N ** 3, 4099), df_sample=df.reshape(-1)[::4099], vecs_sample=vecs.reshape(-1, 3)[::4099])
