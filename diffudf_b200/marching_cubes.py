"""MeshUDF marching cubes: drop-in for /root/reference/src/marching_cubes/_marching_cubes_lewiner.py `udf_mc_lewiner`
(:80-141) over the C++17 restatement of the reference's Cython mesher (csrc/meshudf_mc.cpp -> dudf_meshudf_mc; SURVEY.md
§8f row 4).  Host code: the search is serial by construction (visit order decides the pseudo-signs).  Lewiner's look-up tables
are data (data/lewiner_luts.bin, tools/export_lewiner_luts.py).

`extract_mesh_MESHUDF`'s numerical part (src/render_mc.py:103-134: fields -> clamp -> udf_mc_lewiner -> shift by -1) is
`meshudf_from_fields`; the trimesh clean-up loops and the border smoothing that follow in the reference (:136-199) are mesh
hygiene on third-party code and stay out."""
import ctypes
import os

import numpy as np

from . import _lib

_LUTS = None


def _luts():
    global _LUTS
    if _LUTS is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "lewiner_luts.bin")
        with open(path, "rb") as fh:
            _LUTS = fh.read()
    return _LUTS


def udf_mc_lewiner(volume, grads, spacing=(1., 1., 1.), gradient_direction='descent', step_size=1, allow_degenerate=True,
                   use_classic=False, avg_thresh=1.05, max_thresh=1.75, mask=None):
    """Same arguments, checks, outputs and post-processing as the reference wrapper: returns (vertices (V,3) in z-y-x order
    times spacing, faces (F,3), normals (V,3), values (V,))."""
    if not isinstance(volume, np.ndarray) or (volume.ndim != 3):
        raise ValueError('Input volume should be a 3D numpy array.')
    if volume.shape[0] < 2 or volume.shape[1] < 2 or volume.shape[2] < 2:
        raise ValueError("Input array must be at least 2x2x2.")
    volume = np.ascontiguousarray(volume, np.float32)
    grads = np.ascontiguousarray(grads, np.float32)
    if grads.shape != volume.shape + (3,):
        raise ValueError('grads must have shape volume.shape + (3,).')
    if len(spacing) != 3:
        raise ValueError("`spacing` must consist of three floats.")
    step_size = int(step_size)
    if step_size < 1:
        raise ValueError('step_size must be at least one.')
    if mask is not None:
        if not mask.shape == volume.shape:
            raise ValueError('volume and mask must have the same shape.')
        mask = np.ascontiguousarray(mask, np.uint8)
    if not allow_degenerate:
        raise NotImplementedError("remove_degenerate_faces is not part of the hot path (extract_mesh_MESHUDF keeps degenerate faces)")
    L = _lib.lib()
    luts = _luts()
    res = _lib.MeshResult()
    nz, ny, nx = volume.shape
    _lib.check(L.dudf_meshudf_mc(volume.ctypes.data, grads.ctypes.data, nz, ny, nx, step_size, float(avg_thresh), float(max_thresh),
                                 None if mask is None else mask.ctypes.data, luts, len(luts), ctypes.byref(res)), "dudf_meshudf_mc")
    try:
        nv, nf = int(res.n_vertices), int(res.n_faces)

        def take(ptr, count, dtype):
            if count == 0:
                return np.empty(0, dtype)
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float if dtype == np.float32 else ctypes.c_int32)),
                                         shape=(count,)).astype(dtype, copy=True)
        vertices = take(res.vertices, nv * 3, np.float32).reshape(-1, 3)
        normals = take(res.normals, nv * 3, np.float32).reshape(-1, 3)
        values = take(res.values, nv, np.float32)
        faces = take(res.faces, nf * 3, np.int32).reshape(-1, 3)
    finally:
        L.dudf_meshudf_free(ctypes.byref(res))
    if not len(vertices):
        raise RuntimeError('No surface found at the given iso value.')
    vertices = np.fliplr(vertices)          # z-y-x order, as skimage / the reference
    normals = np.fliplr(normals)
    if gradient_direction == 'descent':
        faces = np.fliplr(faces)
    elif not gradient_direction == 'ascent':
        raise ValueError("Incorrect input %s in `gradient_direction`, see docstring." % (gradient_direction))
    if not np.array_equal(spacing, (1, 1, 1)):
        vertices = vertices * np.r_[spacing]
    return vertices, faces, normals, values


def meshudf_from_fields(df, vecs, N=None):
    """The mesher call of extract_mesh_MESHUDF (src/render_mc.py:127-134): df (N,N,N) distances (clamped at 0), vecs (N,N,N,3)
    pseudo-normals -> (vertices (V,3) in [-1,1]^3, faces (F,3)).  Accepts torch tensors or numpy arrays."""
    to_np = lambda t: t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    df = np.array(to_np(df), np.float32)
    df[df < 0] = 0
    N = df.shape[0] if N is None else N
    voxel = 2.0 / (N - 1)
    v, f, _, _ = udf_mc_lewiner(df, to_np(vecs), spacing=[voxel] * 3, avg_thresh=1.05, max_thresh=1.75)
    return v - 1, f
