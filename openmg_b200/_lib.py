"""ctypes binding of libomg_b200.so (include/omg_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100a
device is visible, every compute entry point raises RuntimeError.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libomg_b200.so")

OK, EINVAL, ECUDA, ENODEV, ESHAPE, EDIM, EINDEX, ESINGULAR, ENOMEM, ENCCL, EUNSUPPORTED = range(11)
SMOOTHERS = {"jacobi": 0, "rbgs": 1, "gs": 2, "lexgs": 2}
KINDS = {0: "band", 1: "band+exc", 2: "csr"}
FLAG_FORCE_CSR, FLAG_NO_GRAPH, FLAG_NO_FUSED, FLAG_KEEP_CSR, FLAG_FACTOR = 1, 2, 4, 8, 16

_lib = None
_inited = False

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_h = ctypes.c_void_p

# name -> (restype, argtypes); the CPU test-suite checks that every name is exported
SIGNATURES = {
    "omg_init": (ctypes.c_int, [ctypes.c_int]),
    "omg_finalize": (None, []),
    "omg_last_error": (ctypes.c_char_p, []),
    "omg_device_info": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, c_i32p, c_i64p]),
    "omg_nccl_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "omg_dist_init": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_char_p]),
    "omg_dist_rank": (ctypes.c_int, [c_i32p, c_i32p]),
    "omg_partition": (ctypes.c_int, [ctypes.c_int, c_i64p, c_i64p, c_i32p, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int64, c_i32p, c_i64p, c_i64p]),
    "omg_level_partition": (ctypes.c_int, [c_h, ctypes.c_int, c_i64p, c_i64p, c_i32p]),
    "omg_host_alloc": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64]),
    "omg_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "omg_hierarchy_create_csr": (ctypes.c_int, [ctypes.POINTER(c_h), ctypes.c_int, c_i64p, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int64, c_i32p, c_i32p, c_f64p,
                                                ctypes.c_int]),
    "omg_hierarchy_create_band": (ctypes.c_int, [ctypes.POINTER(c_h), ctypes.c_int, c_i64p, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_int,
                                                 c_i64p, c_f64p, ctypes.c_int]),
    "omg_operator_create_csr": (ctypes.c_int, [ctypes.POINTER(c_h), ctypes.c_int64, c_i32p, c_i32p, c_f64p,
                                               ctypes.c_int]),
    "omg_hierarchy_destroy": (None, [c_h]),
    "omg_level_count": (ctypes.c_int, [c_h, c_i32p]),
    "omg_level_info": (ctypes.c_int, [c_h, ctypes.c_int, c_i64p, c_i64p, c_i64p, c_i32p, c_i64p]),
    "omg_level_band": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_i32p, c_i64p, c_f64p]),
    "omg_level_export_A": (ctypes.c_int, [c_h, ctypes.c_int, c_i32p, c_i32p, c_f64p]),
    "omg_level_export_R": (ctypes.c_int, [c_h, ctypes.c_int, c_i32p, c_i32p, c_f64p]),
    "omg_setup_times": (ctypes.c_int, [c_h, c_f64p, c_f64p, c_f64p]),
    "omg_restriction": (ctypes.c_int, [ctypes.c_int, c_i64p, c_i64p, c_i64p, c_i32p, c_i32p, c_f64p]),
    "omg_solve": (ctypes.c_int, [c_h, c_f64p, c_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_double, ctypes.c_int, ctypes.c_double, c_i32p, c_f64p, c_f64p,
                                 ctypes.c_int]),
    "omg_solve_stats": (ctypes.c_int, [c_h, c_i64p, c_f64p]),
    "omg_cycle": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_double, c_f64p]),
    "omg_set_rhs": (ctypes.c_int, [c_h, c_f64p]),
    "omg_bench_cycles": (ctypes.c_int, [c_h, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                        ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), c_i64p]),
    "omg_profile_cycle": (ctypes.c_int, [c_h, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                         ctypes.c_int, ctypes.c_char_p, ctypes.c_int]),
    "omg_get_solution": (ctypes.c_int, [c_h, c_f64p]),
    "omg_current_norm": (ctypes.c_int, [c_h, c_f64p]),
    "omg_smooth": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_f64p, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_double]),
    "omg_smooth_to_threshold": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_f64p, ctypes.c_double, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_double, c_i32p, c_f64p]),
    "omg_residual_restrict": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_f64p, c_f64p]),
    "omg_smooth_residual_restrict": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_f64p, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_double, c_f64p]),
    "omg_prolong_correct": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_f64p]),
    "omg_prolong_correct_smooth": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_f64p, c_f64p, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_double]),
    "omg_coarse_solve": (ctypes.c_int, [c_h, c_f64p, c_f64p]),
    "omg_residual_norm": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_f64p, c_f64p]),
    "omg_residual": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_f64p, c_f64p]),
    "omg_matvec": (ctypes.c_int, [c_h, ctypes.c_int, c_f64p, c_f64p]),
}


class SingularMatrixError(np.linalg.LinAlgError, ZeroDivisionError):
    pass


_EXC = {
    EINVAL: ValueError, ECUDA: RuntimeError, ENODEV: RuntimeError, ESHAPE: ValueError, EDIM: ValueError,
    EINDEX: IndexError, ESINGULAR: SingularMatrixError, ENOMEM: MemoryError, ENCCL: RuntimeError,
    EUNSUPPORTED: NotImplementedError,
}


def load(path=None):
    """dlopen the library and attach signatures (no device needed)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            "libomg_b200.so not found at %s — build it with `python -m openmg_b200.build` "
            "(nvcc, sm_100a). openmg_b200 has no CPU fallback." % p)
    lib = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(rc):
    if rc != OK:
        msg = _lib.omg_last_error().decode("utf-8", "replace")
        raise _EXC.get(rc, RuntimeError)(msg)


def lib():
    """The library with the device initialised (raises RuntimeError without a B200)."""
    global _inited
    L = load()
    if not _inited:
        check(L.omg_init(-1))
        _inited = True
    return L


def f64(a):
    return a.ctypes.data_as(c_f64p)


def i32(a):
    return a.ctypes.data_as(c_i32p)


def i64(a):
    return a.ctypes.data_as(c_i64p)


def device_info():
    L = lib()
    buf = ctypes.create_string_buffer(256)
    sm = ctypes.c_int32()
    mem = ctypes.c_int64()
    check(L.omg_device_info(buf, 256, ctypes.byref(sm), ctypes.byref(mem)))
    return {"name": buf.value.decode(), "sm_count": sm.value, "mem_bytes": mem.value}


def pinned_empty(n, dtype=np.float64):
    """numpy array over cudaMallocHost memory (freed when the array is collected)."""
    L = lib()
    nbytes = int(n) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    check(L.omg_host_alloc(ctypes.byref(p), nbytes))
    buf = (ctypes.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(n))

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                L.omg_host_free(self.ptr)
            except Exception:  # noqa: BLE001
                pass
    owner = _Owner(p)
    # keep owner alive as long as any view of the buffer is
    arr = arr.view(_PinnedArray)
    arr._owner = owner
    return arr


class _PinnedArray(np.ndarray):
    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)
