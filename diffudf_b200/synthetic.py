"""Synthetic "complex shape" used by BASELINE config 2: the union of
closed-form primitives (spheres, tori, open discs).  The unsigned distance to a union of
surfaces is the minimum of the parts' distances, so ground truth is exact and no Open3D
is needed.  Batches follow the layout of the reference sampler
(/root/reference/src/dataset.py:14-70,162-163): rows ordered [on | far | near], normals
zero for off-surface rows, distance 0 for on-surface rows.

Host-side numpy only; the device-side sampler is a "next" row.
"""
import numpy as np


def _frame(axis):
    axis = axis / np.linalg.norm(axis)
    h = np.array([1.0, 0, 0]) if abs(axis[0]) < 0.9 else np.array([0, 1.0, 0])
    u = np.cross(axis, h)
    u /= np.linalg.norm(u)
    v = np.cross(axis, u)
    return np.stack([u, v, axis], 0)            # rows: local x, y, z in world coords


class Shape:
    """Union of primitives.  Each part: dict(kind, c, R(3x3 rows=local axes), r0, r1)."""

    def __init__(self, parts):
        self.parts = parts
        self.areas = np.array([self._area(p) for p in parts])

    @staticmethod
    def _area(p):
        if p["kind"] == "sphere":
            return 4 * np.pi * p["r0"] ** 2
        if p["kind"] == "torus":
            return 4 * np.pi ** 2 * p["r0"] * p["r1"]
        return np.pi * p["r0"] ** 2

    def udf(self, pts):
        pts = np.asarray(pts, np.float64)
        best = np.full(pts.shape[0], np.inf)
        for p in self.parts:
            q = (pts - p["c"]) @ p["R"].T
            if p["kind"] == "sphere":
                d = np.abs(np.linalg.norm(q, axis=1) - p["r0"])
            elif p["kind"] == "torus":
                rho = np.hypot(q[:, 0], q[:, 1]) - p["r0"]
                d = np.abs(np.hypot(rho, q[:, 2]) - p["r1"])
            else:
                rho = np.hypot(q[:, 0], q[:, 1])
                d = np.where(rho <= p["r0"], np.abs(q[:, 2]), np.hypot(rho - p["r0"], q[:, 2]))
            best = np.minimum(best, d)
        return best

    def sample_surface(self, n, rng):
        which = rng.choice(len(self.parts), size=n, p=self.areas / self.areas.sum())
        pts = np.empty((n, 3))
        nrm = np.empty((n, 3))
        for i, p in enumerate(self.parts):
            idx = np.nonzero(which == i)[0]
            m = idx.size
            if m == 0:
                continue
            if p["kind"] == "sphere":
                v = rng.normal(size=(m, 3))
                v /= np.linalg.norm(v, axis=1, keepdims=True)
                ql, nl = v * p["r0"], v
            elif p["kind"] == "torus":
                u = rng.uniform(0, 2 * np.pi, m)
                vv = np.empty(m)
                filled = 0
                while filled < m:                     # area-uniform minor angle by rejection
                    cand = rng.uniform(0, 2 * np.pi, 2 * (m - filled) + 8)
                    acc = cand[rng.uniform(0, 1, cand.size) < (p["r0"] + p["r1"] * np.cos(cand)) / (p["r0"] + p["r1"])]
                    k = min(acc.size, m - filled)
                    vv[filled:filled + k] = acc[:k]
                    filled += k
                nl = np.stack([np.cos(vv) * np.cos(u), np.cos(vv) * np.sin(u), np.sin(vv)], 1)
                ql = np.stack([p["r0"] * np.cos(u), p["r0"] * np.sin(u), np.zeros(m)], 1) + p["r1"] * nl
            else:
                rr = p["r0"] * np.sqrt(rng.uniform(0, 1, m))
                th = rng.uniform(0, 2 * np.pi, m)
                ql = np.stack([rr * np.cos(th), rr * np.sin(th), np.zeros(m)], 1)
                nl = np.tile(np.array([0, 0, 1.0]), (m, 1))
            pts[idx] = ql @ p["R"] + p["c"]
            nrm[idx] = nl @ p["R"]
        return pts.astype(np.float32), nrm.astype(np.float32)


def make_shape(seed=0, n_parts=8):
    rng = np.random.default_rng(seed)
    kinds = ["torus", "sphere", "disc"]
    parts = []
    for i in range(n_parts):
        kind = kinds[i % 3]
        c = rng.uniform(-0.45, 0.45, 3)
        R = _frame(rng.normal(size=3))
        if kind == "sphere":
            r0, r1 = rng.uniform(0.12, 0.3), 0.0
        elif kind == "torus":
            r0 = rng.uniform(0.18, 0.3)
            r1 = rng.uniform(0.04, 0.09)
        else:
            r0, r1 = rng.uniform(0.15, 0.35), 0.0
        parts.append(dict(kind=kind, c=c, R=R, r0=r0, r1=r1))
    return Shape(parts)


def make_batch(shape, surf_pts, surf_nrm, batch_size=30000, percentiles=(0.333, 0.666), rng=None):
    """One training batch (coords (1,P,3), normals (1,P,3), dist (1,P,1)) fp32 numpy.

    Mirrors sampleTrainingData (/root/reference/src/dataset.py:14-70): on = int(bs*p0) draws
    with replacement from the surface cloud; off = int(bs*p1) split into far (uniform in
    [-1,1]^3) and near (surface subset + normal * N(0, 0.01))."""
    rng = np.random.default_rng(0) if rng is None else rng
    n_on = int(batch_size * percentiles[0])
    n_off = int(batch_size * percentiles[1])
    n_far = n_off // 2
    n_near = n_off - n_far
    sel = rng.integers(0, surf_pts.shape[0], n_on)
    on_p, on_n = surf_pts[sel], surf_nrm[sel]
    far = rng.uniform(-1, 1, (n_far, 3)).astype(np.float32)
    sub = rng.integers(0, n_on, n_near)
    near = (on_p[sub] + on_n[sub] * rng.normal(0, 0.01, (n_near, 1))).astype(np.float32)
    x = np.concatenate([on_p, far, near], 0).astype(np.float32)
    nrm = np.concatenate([on_n, np.zeros((n_off, 3), np.float32)], 0).astype(np.float32)
    d = np.concatenate([np.zeros(n_on), shape.udf(far), shape.udf(near)]).astype(np.float32)
    return x[None], nrm[None], d[None, :, None]
