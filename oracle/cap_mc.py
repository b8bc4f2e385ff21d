"""TEST INFRASTRUCTURE ONLY (never imported by the product path): CPU restatement of extract_mesh_CAP, the CAP-UDF marching cubes
of /root/reference/src/render_mc.py:201-256, with numpy.

The reference hands every signed 2x2x2 block to mcubes.marching_cubes (PyMCubes 0.1.4, dudf.yml:314, call site
render_mc.py:231); PyMCubes is neither installed nor vendored under /root/reference, so that one step — the per-cell
triangulation of a signed block at iso-value 0 — restates the published marching-cubes algorithm with the case table of
tools/gen_mc_table.py (oracle/mc_table.py).  PARITY UNPINNED for PyMCubes' choice of polygon diagonals, of the ambiguous-face
rule and of its vertex de-duplication inside a cell; everything around it follows the reference line by line: cell order, the
0.008 threshold on the smallest of the 8 distances (:208-215), the pseudo-sign dot(grad[0,0,0], grad[ii,jj,kk]) < 0 (:222-229), the
res.min() < 0 test (:231), vertices shifted by (i, j, k) (:237-239) and mapped by v / (resolution - 1) * 2 - 1 (:251).
Output: triangle soup (T, 3, 3) float64 in the reference's cell order."""
import numpy as np

from .mc_table import EDGE_CORNERS, TRIS


def cell_triangles(res):
    """Marching cubes at iso-value 0 of one float64 block res[ii][jj][kk]; corner c = 4*ii + 2*jj + kk is negative when res < 0."""
    val = res.reshape(8)
    case = 0
    for c in range(8):
        if val[c] < 0:
            case |= 1 << c
    out = []
    for tri in TRIS[case]:
        pts = []
        for e in tri:
            c0, c1 = EDGE_CORNERS[e]
            mu = (0.0 - val[c0]) / (val[c1] - val[c0])
            p0 = np.array([c0 >> 2 & 1, c0 >> 1 & 1, c0 & 1], dtype=np.float64)
            p1 = np.array([c1 >> 2 & 1, c1 >> 1 & 1, c1 & 1], dtype=np.float64)
            pts.append(p0 + mu * (p1 - p0))
        out.append(pts)
    return np.array(out, dtype=np.float64).reshape(-1, 3, 3)


def extract_mesh_CAP(ndf, grad, resolution, threshold=0.008):
    ndf = np.asarray(ndf)
    grad = np.asarray(grad)
    N = resolution
    # candidate cells (smallest corner distance <= threshold), found with shifted views instead of the triple loop; np.argwhere
    # returns them in the loop's order (i, then j, then k)
    cmin = ndf[:-1, :-1, :-1].copy()
    for di in (0, 1):
        for dj in (0, 1):
            for dk in (0, 1):
                cmin = np.minimum(cmin, ndf[di:N - 1 + di, dj:N - 1 + dj, dk:N - 1 + dk])
    tris = []
    for i, j, k in np.argwhere(~(cmin > threshold)):
        ndf_loc = ndf[i:i + 2, j:j + 2, k:k + 2]
        grad_loc = grad[i:i + 2, j:j + 2, k:k + 2]
        res = np.ones((2, 2, 2))
        for ii in range(2):
            for jj in range(2):
                for kk in range(2):
                    val = ndf_loc[ii][jj][kk]
                    g0, g1 = grad_loc[0][0][0], grad_loc[ii][jj][kk]
                    d = np.float32(np.float32(np.float32(g0[0] * g1[0]) + np.float32(g0[1] * g1[1])) + np.float32(g0[2] * g1[2]))
                    res[ii][jj][kk] = -val if d < 0 else val
        if res.min() < 0:
            t = cell_triangles(res)
            t = t + np.array([i, j, k], dtype=np.float64)
            tris.append(t)
    if not tris:
        return np.zeros((0, 3, 3))
    v = np.concatenate(tris)
    return v / (resolution - 1.0) * 2.0 + (-1.0)
