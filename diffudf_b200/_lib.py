"""Loader / builder of libdudf_b200.so (the C-ABI library declared in include/dudf_b200.h).

The library is built in-tree with nvcc for sm_100a only and bound with ctypes: plain pointers and
sizes cross the boundary, torch only provides device memory and the current stream.  There is no
CPU fallback — `lib()` raises when the library is missing and every entry point raises when the
C call reports a failure.
"""
import ctypes
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libdudf_b200.so")
SOURCES = ["dudf_api.cu", "dudf_simt.cu", "dudf_tc.cu", "dudf_tc_train.cu", "dudf_tcx.cu", "dudf_mesh.cu", "dudf_misc.cu", "dudf_sampler.cu", "dudf_cloud_index.cu", "dudf_drivers.cu", "dudf_capmc.cu"]
HOST_SOURCES = ["meshudf_mc.cpp"]            # plain C++17 (no CUDA): g++, strict IEEE arithmetic (no contraction)
HOST_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off"]
HEADERS = ["dudf_common.cuh", "dudf_kernels.h", "dudf_device.cuh", "dudf_umma.cuh", "dudf_tc_common.cuh", "dudf_loss.cuh", "dudf_mc_table.h",
           os.path.join(ROOT, "include", "dudf_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]

_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

class Segment(ctypes.Structure):
    """struct dudf_segment of include/dudf_b200.h"""
    _fields_ = [("x", c_void_p), ("rows", c_int64), ("order", c_int), ("packed", c_void_p), ("seeds", c_void_p), ("col0", c_int64)]


class MeshResult(ctypes.Structure):
    """struct dudf_meshudf_result of include/dudf_b200.h"""
    _fields_ = [("vertices", c_void_p), ("normals", c_void_p), ("values", c_void_p), ("faces", c_void_p), ("n_vertices", c_int64),
                ("n_faces", c_int64)]


class TrainSegment(ctypes.Structure):
    """struct dudf_train_segment of include/dudf_b200.h"""
    _fields_ = [("x", c_void_p), ("normals", c_void_p), ("dist", c_void_p), ("rows", c_int64), ("order", c_int), ("packed", c_void_p)]


# name -> (argtypes) ; every function returns int except the two noted below
SIGNATURES = {
    "dudf_jet_forward_multi": [c_void_p, ctypes.POINTER(Segment), c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "dudf_jet_backward_multi": [c_void_p, ctypes.POINTER(Segment), c_int, c_void_p, c_void_p, c_void_p, c_int64,
                                ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_int, c_void_p],
    "dudf_fused_scratch_bytes": [c_void_p],
    "dudf_train_step_fused": [c_void_p, c_int, ctypes.POINTER(TrainSegment), c_int, c_int64, ctypes.POINTER(c_float), c_float, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, ctypes.POINTER(c_void_p),
                              ctypes.POINTER(c_void_p), c_int, c_void_p],
    "dudf_bench_umma": [c_int, c_int, c_int, ctypes.POINTER(c_float)],
    "dudf_sample_batch_pc": [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, ctypes.POINTER(c_float),
                             ctypes.POINTER(c_float), ctypes.c_uint64, ctypes.c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p],
    "dudf_cloud_index_bytes": [c_int64],
    "dudf_cloud_index_build": [c_void_p, c_int64, c_void_p, c_void_p],
    "dudf_nearest_distance_indexed": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p],
    "dudf_sample_batch_pc_indexed": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_float, ctypes.POINTER(c_float),
                                     ctypes.POINTER(c_float), ctypes.c_uint64, ctypes.c_uint64, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p],
    "dudf_mesh_distance": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p],
    "dudf_sample_batch_mesh": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, ctypes.POINTER(c_float),
                               ctypes.POINTER(c_float), ctypes.c_uint64, ctypes.c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p],
    "dudf_mesh_sample_surface": [c_void_p, c_void_p, c_int64, c_int64, ctypes.c_uint64, c_void_p, c_void_p, c_void_p, c_void_p],
    "dudf_nearest_distance": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p],
    "dudf_march_rays": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_int, c_int,
                        ctypes.POINTER(c_int64), c_void_p],
    "dudf_project_points": [c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "dudf_cap_mesh": [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_int64, ctypes.POINTER(c_int64), c_void_p],
    "dudf_shade_hits": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(ctypes.c_double),
                        ctypes.POINTER(ctypes.c_double), c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_void_p, c_void_p],
    "dudf_debug_set_trace": [c_void_p],
    "dudf_version": [],
    "dudf_launch_count": [],
    "dudf_create": [c_int, c_float, c_float, ctypes.POINTER(c_void_p)],
    "dudf_destroy": [c_void_p],
    "dudf_set_weights": [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_void_p],
    "dudf_bind_weights": [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p)],
    "dudf_refresh_weights": [c_void_p, c_int, c_void_p],
    "dudf_query_points": [c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "dudf_query_grid": [c_void_p, c_int, c_int64, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "dudf_eig_normals": [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p],
    "dudf_curvature": [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "dudf_field_vectors": [c_void_p, c_void_p, c_int64, c_void_p, c_void_p],
    "dudf_evaluate_host": [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "dudf_stash_columns": [c_int, c_int64, c_int],
    "dudf_jet_forward": [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p],
    "dudf_jet_backward": [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                          ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_int, c_void_p],
    "dudf_jet_wgrad": [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, ctypes.POINTER(c_void_p), c_int, c_void_p],
    "dudf_jet_wgrad_layers": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_void_p],
    "dudf_loss": [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, ctypes.POINTER(c_float), c_float,
                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "dudf_loss_s2_stats": [c_void_p, c_void_p, c_int64, c_void_p, c_void_p],
    "dudf_loss_s2_finish": [c_void_p, c_float, c_float, c_void_p, c_void_p],
    "dudf_adam_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_int64, c_void_p],
    "dudf_adam_step_dev": [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float, c_float, c_void_p],
    "dudf_adam_step_guarded": [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_int64, c_void_p,
                               c_void_p, c_void_p],
    "dudf_scale_guard": [c_void_p, c_void_p, c_float, c_void_p, c_void_p],
    "dudf_mean_curvature": [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p],
    "dudf_adam_step_peers": [c_void_p, ctypes.POINTER(c_void_p), c_int, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                             c_int64, c_int, c_void_p, c_void_p, c_void_p],
    "dudf_meshudf_mc": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_int64, c_void_p],
    "dudf_meshudf_free": [c_void_p],
    "dudf_selftest_umma": [c_int, ctypes.POINTER(c_float)],
}


def _needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HOST_SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into diffudf_b200/libdudf_b200.so (nvcc cross-compiles
    without a GPU).  Objects are compiled in parallel, then linked."""
    if not force and not _needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v", "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s in HOST_SOURCES:
        obj = os.path.join(objdir, s.replace(".cpp", ".o"))
        cmd = [os.environ.get("CXX", "g++")] + HOST_FLAGS + ["-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        with open(os.path.join(objdir, s + ".log"), "w") as fh:
            fh.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"compiler failed for {s}:\n{out}")
        if verbose:
            print(out)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


def lib():
    """The loaded library (ctypes.CDLL) with argtypes set.  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(diffudf_b200 has no CPU fallback)")
            L = ctypes.CDLL(LIB_PATH)
            for name, args in SIGNATURES.items():
                fn = getattr(L, name)
                fn.argtypes = args
                fn.restype = c_int64 if name in ("dudf_stash_columns", "dudf_launch_count", "dudf_fused_scratch_bytes", "dudf_cloud_index_bytes") else c_int
            L.dudf_last_error.argtypes = []
            L.dudf_last_error.restype = ctypes.c_char_p
            _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().dudf_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device (or host) pointer of a tensor / numpy array, None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
