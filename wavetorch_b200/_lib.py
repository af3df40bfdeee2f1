"""ctypes binding of lib/libwavetorch_b200.so (C ABI declared in include/wavetorch_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails this module raises.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WT_LIB") or os.path.join(_HERE, "lib", "libwavetorch_b200.so")   # WT_LIB: a debug build

WT_F_ZERO_INIT = 1
WT_F_FORCE_STREAM = 2
WT_F_FORCE_RESIDENT = 4
WT_F_NEED_GRAD_B = 8
WT_F_NO_SPECIALIZE = 16
WT_F_NO_PLAIN_WARPS = 32
WT_PATH_STREAM = 0
WT_PATH_RESIDENT = 1

EXPORTS = ("wt_abi_version", "wt_last_error", "wt_query_plan", "wt_validate_pixels", "wt_forward", "wt_backward", "wt_step_forward",
           "wt_step_backward", "wt_geom_forward", "wt_geom_backward", "wt_loss_forward", "wt_loss_backward",
           "wt_peer_allreduce", "wt_slab_forward", "wt_slab_backward", "wt_slab_exchange", "wt_query_plan_f64", "wt_forward_f64",
           "wt_backward_f64")


class WtProblem(ctypes.Structure):
    _fields_ = [("Nx", ctypes.c_int32), ("Ny", ctypes.c_int32), ("B", ctypes.c_int32), ("T", ctypes.c_int32),
                ("n_src", ctypes.c_int32), ("n_prb", ctypes.c_int32), ("flags", ctypes.c_uint32),
                ("device", ctypes.c_int32), ("dt", ctypes.c_double), ("h", ctypes.c_double),
                ("b0", ctypes.c_double), ("uth", ctypes.c_double), ("c_nl", ctypes.c_double),
                ("cluster", ctypes.c_int32), ("rows_per_thread", ctypes.c_int32), ("field_every", ctypes.c_int32),
                ("checkpoint_every", ctypes.c_int32), ("reserved", ctypes.c_int32 * 4)]


class WtPlan(ctypes.Structure):
    _fields_ = [("path", ctypes.c_int32), ("cluster", ctypes.c_int32), ("rows_per_thread", ctypes.c_int32),
                ("threads", ctypes.c_int32), ("rows_per_cta", ctypes.c_int32), ("n_clusters", ctypes.c_int32),
                ("smem_fwd", ctypes.c_int32), ("smem_bwd", ctypes.c_int32), ("nonlinear", ctypes.c_int32),
                ("launches_fwd", ctypes.c_int32), ("launches_bwd", ctypes.c_int32), ("reserved", ctypes.c_int32 * 5),
                ("history_bytes", ctypes.c_uint64), ("workspace_fwd_bytes", ctypes.c_uint64),
                ("workspace_bwd_bytes", ctypes.c_uint64)]


_lock = threading.Lock()
_lib = None
launch_count = 0   # kernels launched through this binding (bench.py reports it as gpu_launches)


def load():
    """Load the shared library once; raise with build instructions when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "wavetorch_b200: %s not found. Build it with `make -C wavetorch_b200/csrc` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU/PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        vp, i32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        lib.wt_abi_version.restype = ctypes.c_int
        lib.wt_last_error.restype = ctypes.c_char_p
        lib.wt_query_plan.argtypes = [ctypes.POINTER(WtProblem), ctypes.POINTER(WtPlan)]
        lib.wt_validate_pixels.argtypes = [ctypes.POINTER(WtProblem), vp, vp, ctypes.POINTER(ctypes.c_int32)]
        lib.wt_forward.argtypes = [ctypes.POINTER(WtProblem)] + [vp] * 12 + [vp, sz, vp, sz, vp]
        lib.wt_backward.argtypes = [ctypes.POINTER(WtProblem)] + [vp] * 9 + [vp, sz] + [vp] * 6 + [vp, sz, vp]
        lib.wt_step_forward.argtypes = [ctypes.POINTER(WtProblem), vp, i32, vp, i32, vp, vp, vp, vp]
        lib.wt_step_backward.argtypes = [ctypes.POINTER(WtProblem), vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.wt_geom_forward.argtypes = [i32, i32, i32, i32] + [vp] * 8 + [i32, vp]
        lib.wt_geom_backward.argtypes = [i32, i32, i32, i32] + [vp] * 9 + [i32, vp]
        lib.wt_loss_forward.argtypes = [i32, i32, i32, i32] + [vp] * 6 + [i32, vp]
        lib.wt_loss_backward.argtypes = [i32, i32, i32] + [vp] * 3 + [i32, vp]
        lib.wt_peer_allreduce.argtypes = [i32, i32, i32, i32, ctypes.c_float, vp, vp, ctypes.POINTER(ctypes.c_uint64),
                                          ctypes.c_uint64, vp, i32, vp]
        lib.wt_slab_forward.argtypes = [ctypes.POINTER(WtProblem), vp] + [vp] * 11 + [vp, sz, vp, sz, vp]
        lib.wt_slab_backward.argtypes = [ctypes.POINTER(WtProblem), vp] + [vp] * 8 + [vp, sz] + [vp] * 4 + [vp, sz, vp]
        lib.wt_slab_exchange.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp]
        lib.wt_query_plan_f64.argtypes = [ctypes.POINTER(WtProblem), ctypes.POINTER(WtPlan)]
        lib.wt_forward_f64.argtypes = [ctypes.POINTER(WtProblem)] + [vp] * 12 + [vp, sz, vp, sz, vp]
        lib.wt_backward_f64.argtypes = [ctypes.POINTER(WtProblem)] + [vp] * 8 + [vp, sz] + [vp] * 4 + [vp, sz, vp]
        for name in ("wt_slab_forward", "wt_slab_backward", "wt_slab_exchange", "wt_query_plan_f64", "wt_forward_f64",
                     "wt_backward_f64"):
            getattr(lib, name).restype = ctypes.c_int
        for name in ("wt_query_plan", "wt_validate_pixels", "wt_forward", "wt_backward", "wt_step_forward", "wt_step_backward",
                     "wt_geom_forward", "wt_geom_backward", "wt_loss_forward", "wt_loss_backward",
                     "wt_peer_allreduce"):
            getattr(lib, name).restype = ctypes.c_int
        if lib.wt_abi_version() != 1:
            raise RuntimeError("wavetorch_b200: ABI version mismatch (library %d, binding 1)" % lib.wt_abi_version())
        _lib = lib
    return _lib


def check(status, what):
    if status != 0:
        msg = load().wt_last_error().decode("utf-8", "replace")
        raise RuntimeError("wavetorch_b200.%s failed (status %d): %s" % (what, status, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be contiguous."""
    if t is None:
        return None
    assert t.is_contiguous(), "wavetorch_b200: non-contiguous tensor at the C boundary"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def make_problem(Nx, Ny, B, T, n_src, n_prb, dt, h, b0=0.0, uth=0.0, c_nl=0.0, flags=0, device=0, cluster=0,
                 rows_per_thread=0):
    p = WtProblem()
    p.Nx, p.Ny, p.B, p.T, p.n_src, p.n_prb = int(Nx), int(Ny), int(B), int(T), int(n_src), int(n_prb)
    if os.environ.get("WT_RES_NOSPEC", "0") == "1":     # A/B switch: generic instead of shape-specialised kernels
        flags |= WT_F_NO_SPECIALIZE
    if os.environ.get("WT_RES_NOPLAIN", "0") == "1":    # A/B switch: no plain-warp instantiation of the time step
        flags |= WT_F_NO_PLAIN_WARPS
    p.flags, p.device = int(flags), int(device)
    p.dt, p.h, p.b0, p.uth, p.c_nl = float(dt), float(h), float(b0), float(uth), float(c_nl)
    p.cluster = int(os.environ.get("WT_CLUSTER", cluster))
    p.rows_per_thread = int(os.environ.get("WT_ROWS", rows_per_thread))
    return p


def query_plan(problem):
    plan = WtPlan()
    check(load().wt_query_plan(ctypes.byref(problem), ctypes.byref(plan)), "wt_query_plan")
    return plan


WT_MAX_SRC_LISTINGS = 2


def validate_pixels(Nx, Ny, src_ij_host, prb_ij_host):
    """wt_validate_pixels on HOST int32 tensors [n,2]: raises on out-of-range coordinates, returns the largest number of
    listings of one source pixel."""
    p = make_problem(Nx, Ny, 1, 1, src_ij_host.shape[0], prb_ij_host.shape[0], 1.0, 1.0)
    s = src_ij_host.contiguous()
    q = prb_ij_host.contiguous()
    most = ctypes.c_int32(0)
    st = load().wt_validate_pixels(ctypes.byref(p), ctypes.c_void_p(s.data_ptr()) if s.numel() else None,
                                   ctypes.c_void_p(q.data_ptr()) if q.numel() else None, ctypes.byref(most))
    if st != 0:
        raise IndexError(load().wt_last_error().decode("utf-8", "replace"))
    return int(most.value)


def count_launches(n):
    global launch_count
    launch_count += int(n)
