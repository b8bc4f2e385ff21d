"""Drop-in for /root/reference/src/dataset.py: `PointCloud` (:134-185) in both its modes, `sampleTrainingDataPC` (:80-131),
`sampleTrainingData` (:14-70) and `shortestDistance` (:72-78), produced on the device by `dudf_sample_batch_pc` /
`dudf_sample_batch_mesh` / `dudf_nearest_distance` / `dudf_mesh_distance` (include/dudf_b200.h) — no host sampling, no
host->device copy.

Open3D (`.ply` IO, RaycastingScene) is a third-party consumer that is not rebuilt: construct from arrays (points + normals,
and the triangle list for the mesh mode).  The mesh mode measures the UNSIGNED distance of every off-surface row to the
triangles by brute force; Open3D's ray-parity sign is not reproduced (the losses are even in the distance, and the sign is
undefined for the open surfaces DUDF targets).  Iterating yields the reference's triple
`(coords (1, P, 3), normals (1, P, 3), sdf (1, P, 1))`, fp32, as CUDA tensors ordered [on | far | near]."""
import ctypes

import numpy as np
import torch

from . import _lib


def _as_dev(t, device, dtype=torch.float32):
    return torch.as_tensor(t).to(device=device, dtype=dtype).contiguous()


class CloudIndex:
    """Spatial index of a cloud (dudf_cloud_index_build): Morton-sorted points under a 32-ary box hierarchy in one device buffer.
    The reference measures every batch's far rows against the same cloud (src/dataset.py:116-118); `PointCloud` builds this once
    and every batch walks it instead of scanning the cloud.  Distances are exact in fp32."""

    def __init__(self, X):
        if X.device.type != "cuda":
            raise RuntimeError("diffudf_b200.dataset.CloudIndex needs a CUDA (sm_100) tensor; there is no CPU fallback")
        self.cloud = X.detach().to(torch.float32).contiguous()
        self.n = int(self.cloud.shape[0])
        nbytes = int(_lib.lib().dudf_cloud_index_bytes(self.n))
        if nbytes < 0:
            raise ValueError("CloudIndex: 1 .. 2^30 points")
        self.buffer = torch.empty(nbytes, device=self.cloud.device, dtype=torch.uint8)
        with torch.cuda.device(self.cloud.device):
            _lib.check(_lib.lib().dudf_cloud_index_build(self.cloud.data_ptr(), self.n, self.buffer.data_ptr(), _lib.current_stream()),
                       "dudf_cloud_index_build")

    def distance(self, P):
        """(n, 3) CUDA queries -> (n,) fp32 distances to the nearest cloud point."""
        P = P.detach().to(device=self.cloud.device, dtype=torch.float32).contiguous()
        out = torch.empty(P.shape[0], device=P.device, dtype=torch.float32)
        if P.shape[0]:
            with torch.cuda.device(P.device):
                _lib.check(_lib.lib().dudf_nearest_distance_indexed(P.data_ptr(), P.shape[0], self.buffer.data_ptr(), self.n, out.data_ptr(),
                                                                    _lib.current_stream()), "dudf_nearest_distance_indexed")
        return out


def shortestDistance(P, X):
    """Distance from each row of P to its nearest row of X (CUDA tensors (n, 3), (m, 3)) -> (n,) fp32.  X may be a `CloudIndex`."""
    if isinstance(X, CloudIndex):
        return X.distance(P)
    if P.device.type != "cuda":
        raise RuntimeError("diffudf_b200.dataset.shortestDistance needs CUDA (sm_100) tensors; there is no CPU fallback")
    P = P.detach().to(torch.float32).contiguous()
    X = X.detach().to(device=P.device, dtype=torch.float32).contiguous()
    out = torch.empty(P.shape[0], device=P.device, dtype=torch.float32)
    if P.shape[0] == 0:
        return out
    with torch.cuda.device(P.device):
        _lib.check(_lib.lib().dudf_nearest_distance(P.data_ptr(), P.shape[0], X.data_ptr(), X.shape[0], out.data_ptr(),
                                                    _lib.current_stream()), "dudf_nearest_distance")
    return out


def _triangles(triangles, device):
    """(n_tri, 3, 3) float32 CUDA tensor from a (n_tri, 3, 3) array or a (vertices (n,3), faces (m,3)) pair."""
    if isinstance(triangles, (tuple, list)) and len(triangles) == 2:
        V, F = triangles
        V = torch.as_tensor(np.asarray(V) if not torch.is_tensor(V) else V)
        F = torch.as_tensor(np.asarray(F) if not torch.is_tensor(F) else F).long()
        triangles = V[F]
    t = _as_dev(np.asarray(triangles) if not torch.is_tensor(triangles) else triangles, device)
    if t.ndim != 3 or t.shape[1:] != (3, 3):
        raise ValueError("triangles must have shape (n_tri, 3, 3)")
    return t


def meshDistance(P, triangles):
    """Unsigned distance from each row of P (n, 3) to the triangle mesh (what |scene.compute_signed_distance| returns,
    src/dataset.py:35,50) -> (n,) fp32."""
    if P.device.type != "cuda":
        raise RuntimeError("diffudf_b200.dataset.meshDistance needs CUDA (sm_100) tensors; there is no CPU fallback")
    P = P.detach().to(torch.float32).contiguous()
    T = _triangles(triangles, P.device)
    out = torch.empty(P.shape[0], device=P.device, dtype=torch.float32)
    if P.shape[0] == 0:
        return out
    with torch.cuda.device(P.device):
        _lib.check(_lib.lib().dudf_mesh_distance(P.data_ptr(), P.shape[0], T.data_ptr(), T.shape[0], out.data_ptr(),
                                                 _lib.current_stream()), "dudf_mesh_distance")
    return out


def sampleTrainingData(surface_pc, surface_normals, samplesOnSurface, samplesOffSurface, triangles, domainBounds=([-1, -1, -1], [1, 1, 1]),
                       seed=0, batch_index=0, draws=None, sigma=0.01):
    """One batch of the mesh mode (reference :14-70; `scene` becomes the triangle list): like sampleTrainingDataPC, with the
    distance of the far AND the near rows measured to the mesh."""
    dev = surface_pc.device
    if dev.type != "cuda":
        raise RuntimeError("diffudf_b200.dataset.sampleTrainingData needs the cloud on a CUDA (sm_100) device; there is no CPU fallback")
    X = surface_pc.detach().to(torch.float32).contiguous()
    N = surface_normals.detach().to(device=dev, dtype=torch.float32).contiguous()
    T = _triangles(triangles, dev)
    n_on, n_off = int(samplesOnSurface), int(samplesOffSurface)
    n_far = n_off // 2
    n_near = n_off - n_far
    P = n_on + n_off
    coords = torch.empty(1, P, 3, device=dev, dtype=torch.float32)
    normals = torch.empty(1, P, 3, device=dev, dtype=torch.float32)
    sdf = torch.empty(1, P, 1, device=dev, dtype=torch.float32)
    d = draws or {}
    keep = [_as_dev(d[k], dev, torch.int64 if k.endswith("idx") else torch.float32) if d.get(k) is not None else None
            for k in ("on_idx", "far", "near_idx", "near_off")]
    lo = (ctypes.c_float * 3)(*[float(v) for v in domainBounds[0]])
    hi = (ctypes.c_float * 3)(*[float(v) for v in domainBounds[1]])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dudf_sample_batch_mesh(X.data_ptr(), N.data_ptr(), X.shape[0], T.data_ptr(), T.shape[0], n_on, n_far, n_near,
                                                     float(sigma), lo, hi, int(seed) & (2 ** 64 - 1), int(batch_index), _lib.ptr(keep[0]),
                                                     _lib.ptr(keep[1]), _lib.ptr(keep[2]), _lib.ptr(keep[3]), coords.data_ptr(),
                                                     normals.data_ptr(), sdf.data_ptr(), _lib.current_stream()), "dudf_sample_batch_mesh")
    return coords, normals, sdf


def sampleTrainingDataPC(surface_pc, surface_normals, samplesOnSurface, samplesOffSurface, domainBounds=([-1, -1, -1], [1, 1, 1]),
                         seed=0, batch_index=0, draws=None, sigma=0.01, index=None):
    """One batch (reference :80-131).  `draws` (optional): dict with any of on_idx (n_on,) int64, far (n_far, 3),
    near_idx (n_near,) int64, near_off (n_near,) — replaces the corresponding Philox draws (parity tests).
    `index` (optional): a `CloudIndex` of surface_pc, built once by the caller, that serves the far rows' distances."""
    dev = surface_pc.device
    if dev.type != "cuda":
        raise RuntimeError("diffudf_b200.dataset.sampleTrainingDataPC needs the cloud on a CUDA (sm_100) device; there is no CPU fallback")
    X = surface_pc.detach().to(torch.float32).contiguous()
    N = surface_normals.detach().to(device=dev, dtype=torch.float32).contiguous()
    n_on, n_off = int(samplesOnSurface), int(samplesOffSurface)
    n_far = n_off // 2
    n_near = n_off - n_far
    P = n_on + n_off
    coords = torch.empty(1, P, 3, device=dev, dtype=torch.float32)
    normals = torch.empty(1, P, 3, device=dev, dtype=torch.float32)
    sdf = torch.empty(1, P, 1, device=dev, dtype=torch.float32)
    d = draws or {}
    keep = [_as_dev(d[k], dev, torch.int64 if k.endswith("idx") else torch.float32) if d.get(k) is not None else None
            for k in ("on_idx", "far", "near_idx", "near_off")]
    lo = (ctypes.c_float * 3)(*[float(v) for v in domainBounds[0]])
    hi = (ctypes.c_float * 3)(*[float(v) for v in domainBounds[1]])
    with torch.cuda.device(dev):
        if index is not None:
            if index.n != X.shape[0] or index.buffer.device != dev:
                raise ValueError("sampleTrainingDataPC: the index was built for another cloud")
            _lib.check(_lib.lib().dudf_sample_batch_pc_indexed(X.data_ptr(), N.data_ptr(), X.shape[0], index.buffer.data_ptr(), n_on, n_far, n_near,
                                                               float(sigma), lo, hi, int(seed) & (2 ** 64 - 1), int(batch_index),
                                                               _lib.ptr(keep[0]), _lib.ptr(keep[1]), _lib.ptr(keep[2]), _lib.ptr(keep[3]),
                                                               coords.data_ptr(), normals.data_ptr(), sdf.data_ptr(), _lib.current_stream()),
                       "dudf_sample_batch_pc_indexed")
        else:
            _lib.check(_lib.lib().dudf_sample_batch_pc(X.data_ptr(), N.data_ptr(), X.shape[0], n_on, n_far, n_near, float(sigma), lo, hi,
                                                       int(seed) & (2 ** 64 - 1), int(batch_index), _lib.ptr(keep[0]), _lib.ptr(keep[1]),
                                                       _lib.ptr(keep[2]), _lib.ptr(keep[3]), coords.data_ptr(), normals.data_ptr(),
                                                       sdf.data_ptr(), _lib.current_stream()), "dudf_sample_batch_pc")
    return coords, normals, sdf


class PointCloud(torch.utils.data.IterableDataset):
    """Iterable of device-generated batches; same attributes as the reference dataset (`batchesPerEpoch`,
    `samplesOnSurface`, `samplesFarSurface`) so the training loops take it unchanged."""

    def __init__(self, points, normals, batchSize, samplingPercentiles, batchesPerEpoch, device, seed=0, triangles=None, prefetch=True):
        """triangles=None is the reference's onlyPCloud=True mode (distances to the cloud); with a triangle list
        ((n_tri,3,3) or (vertices, faces)) the off-surface distances are measured to the mesh (onlyPCloud=False)."""
        super().__init__()
        self.device = torch.device(device)
        self.onlyPCloud = triangles is None
        self.triangles = None if triangles is None else _triangles(triangles, self.device)
        self.surface_pc = _as_dev(np.asarray(points) if not torch.is_tensor(points) else points, self.device)
        self.surface_normals = _as_dev(np.asarray(normals) if not torch.is_tensor(normals) else normals, self.device)
        if self.surface_pc.ndim != 2 or self.surface_pc.shape[1] != 3 or self.surface_pc.shape != self.surface_normals.shape:
            raise ValueError("PointCloud: points and normals must both be (n, 3)")
        self.batchSize = batchSize
        self.samplesOnSurface = int(batchSize * samplingPercentiles[0])
        self.samplesFarSurface = int(batchSize * samplingPercentiles[1])
        self.batchesPerEpoch = batchesPerEpoch
        self.seed = seed
        self.batches_drawn = 0
        self.prefetch = prefetch
        self._side = None
        self.index = CloudIndex(self.surface_pc) if self.onlyPCloud else None      # built once; every batch's far rows walk it

    def _draw(self):
        if self.onlyPCloud:
            out = sampleTrainingDataPC(self.surface_pc, self.surface_normals, self.samplesOnSurface, self.samplesFarSurface,
                                       seed=self.seed, batch_index=self.batches_drawn, index=self.index)
        else:
            out = sampleTrainingData(self.surface_pc, self.surface_normals, self.samplesOnSurface, self.samplesFarSurface,
                                     self.triangles, seed=self.seed, batch_index=self.batches_drawn)
        self.batches_drawn += 1
        return out

    def __iter__(self):
        """Batch i + 1 is drawn on a side stream while the consumer works on batch i (the reference draws synchronously inside the
        loop, src/dataset.py:176-185): the sampler's kernels fill the SMs the persistent training kernels leave idle at their
        tails.  `prefetch=False` draws in the consumer's stream."""
        if not self.prefetch:
            for _ in range(self.batchesPerEpoch):
                yield self._draw()
            return
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        side = self._side
        side.wait_stream(main)                 # the cloud / triangle uploads of the constructor

        def draw_ahead():
            with torch.cuda.stream(side):
                out = self._draw()
                ev = torch.cuda.Event()
                ev.record(side)
            return out, ev
        nxt = draw_ahead() if self.batchesPerEpoch > 0 else None
        for i in range(self.batchesPerEpoch):
            cur, ev = nxt
            nxt = draw_ahead() if i + 1 < self.batchesPerEpoch else None
            main = torch.cuda.current_stream(self.device)
            main.wait_event(ev)
            for t in cur:                       # allocated on the side stream, consumed on this one
                t.record_stream(main)
            yield cur
