"""fp64 numpy restatement of the mesh half of the batch sampler — TEST INFRASTRUCTURE ONLY (tests/ import it as the checker).

Follows: src/dataset.py:14-70 (sampleTrainingData: far rows uniform in the domain, near rows = surface rows displaced along
their normal by N(0, 0.01), distance of every off-surface row to the mesh) and src/preprocess_mesh.py:5-15,29-40 (normalise by
the vertex mean and 1.1 max|coord|; area-weighted surface samples with triangle normals).

Third-party arithmetic: the reference calls Open3D 0.17 (RaycastingScene.compute_signed_distance, sample_points_uniformly),
whose source is not under /root/reference — PARITY UNPINNED for Open3D's sign convention and random stream; what is
restated is the published geometry: |d| = min over triangles of the Euclidean point-triangle distance (closest-point
regions of a triangle, Ericson, "Real-Time Collision Detection" 5.1.5) and Open3D's documented barycentric map
(1 - sqrt r1, sqrt r1 (1 - r2), sqrt r1 r2)."""
import numpy as np


def read_obj(path):
    V, F = [], []
    for line in open(path):
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            V.append([float(t) for t in p[1:4]])
        elif p[0] == "f":
            idx = [int(t.split("/")[0]) - 1 for t in p[1:]]
            for k in range(1, len(idx) - 1):
                F.append([idx[0], idx[k], idx[k + 1]])
    return np.array(V, np.float64), np.array(F, np.int64)


def normalize_vertices(V):
    """preprocess_mesh.normalizeMesh: centre = mean of the vertices (Open3D get_center), scale 1 / (1.1 max|coord|)."""
    V = np.asarray(V, np.float64)
    V = V - V.mean(0)
    return V / (1.1 * np.abs(V).max())


def point_triangle_distance(P, tri, chunk=256):
    """P (n,3), tri (m,3,3) -> (n,) unsigned distances, fp64, by region classification of the closest point."""
    P = np.asarray(P, np.float64)
    tri = np.asarray(tri, np.float64)
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    ab, ac = b - a, c - a
    out = np.empty(P.shape[0])
    for i0 in range(0, P.shape[0], chunk):
        p = P[i0:i0 + chunk, None, :]                       # (q,1,3)
        ap = p - a[None]
        d1 = np.einsum("qmk,mk->qm", ap, ab)
        d2 = np.einsum("qmk,mk->qm", ap, ac)
        bp = p - b[None]
        d3 = np.einsum("qmk,mk->qm", bp, ab)
        d4 = np.einsum("qmk,mk->qm", bp, ac)
        cp = p - c[None]
        d5 = np.einsum("qmk,mk->qm", cp, ab)
        d6 = np.einsum("qmk,mk->qm", cp, ac)
        vc = d1 * d4 - d3 * d2
        vb = d5 * d2 - d1 * d6
        va = d3 * d6 - d5 * d4
        q = p.shape[0]
        m = a.shape[0]
        closest = np.zeros((q, m, 3))
        done = np.zeros((q, m), bool)

        def put(mask, pts):
            sel = mask & ~done
            closest[sel] = pts[sel]
            done[sel] = True

        A = np.broadcast_to(a[None], (q, m, 3))
        B = np.broadcast_to(b[None], (q, m, 3))
        C = np.broadcast_to(c[None], (q, m, 3))
        AB = np.broadcast_to(ab[None], (q, m, 3))
        AC = np.broadcast_to(ac[None], (q, m, 3))
        with np.errstate(divide="ignore", invalid="ignore"):
            put((d1 <= 0) & (d2 <= 0), A)
            put((d3 >= 0) & (d4 <= d3), B)
            v = d1 / (d1 - d3)
            put((vc <= 0) & (d1 >= 0) & (d3 <= 0), A + v[..., None] * AB)
            put((d6 >= 0) & (d5 <= d6), C)
            w = d2 / (d2 - d6)
            put((vb <= 0) & (d2 >= 0) & (d6 <= 0), A + w[..., None] * AC)
            w2 = (d4 - d3) / ((d4 - d3) + (d5 - d6))
            put((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), B + w2[..., None] * (C - B))
            den = va + vb + vc
            vv, ww = vb / den, vc / den
            put(np.ones((q, m), bool), A + vv[..., None] * AB + ww[..., None] * AC)
        dist = np.linalg.norm(p - closest, axis=-1)
        out[i0:i0 + chunk] = np.nanmin(dist, axis=1)
    return out


def area_cdf(tri):
    tri = np.asarray(tri, np.float64)
    cr = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    area = 0.5 * np.linalg.norm(cr, axis=1)
    return np.cumsum(area) / area.sum(), cr


def sample_surface(tri, draws):
    """draws (n,3) = (u_triangle, r1, r2) in [0,1): area-weighted samples with triangle normals."""
    tri = np.asarray(tri, np.float64)
    cdf, cr = area_cdf(tri)
    t = np.minimum(np.searchsorted(cdf.astype(np.float32), draws[:, 0].astype(np.float32), side="right"), len(tri) - 1)
    sr = np.sqrt(draws[:, 1].astype(np.float64))
    r2 = draws[:, 2].astype(np.float64)
    pts = (1 - sr)[:, None] * tri[t, 0] + (sr * (1 - r2))[:, None] * tri[t, 1] + (sr * r2)[:, None] * tri[t, 2]
    nrm = cr[t] / np.maximum(np.linalg.norm(cr[t], axis=1, keepdims=True), 1e-30)
    return pts, nrm, t


def sample_training_data(surf_pts, surf_nrm, tri, n_on, n_off, on_idx, far, near_idx, near_off):
    """sampleTrainingData with the draws supplied: returns coords (P,3), normals (P,3), |sdf| (P,)."""
    sp = np.asarray(surf_pts, np.float64)[on_idx]
    sn = np.asarray(surf_nrm, np.float64)[on_idx]
    close = sp[near_idx] + sn[near_idx] * np.asarray(near_off, np.float64)[:, None]
    far = np.asarray(far, np.float64)
    coords = np.concatenate([sp, far, close])
    normals = np.concatenate([sn, np.zeros((n_off, 3))])
    d = np.concatenate([np.zeros(n_on), point_triangle_distance(far, tri), point_triangle_distance(close, tri)])
    return coords, normals, d
