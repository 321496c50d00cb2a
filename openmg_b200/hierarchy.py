"""Device-resident multigrid hierarchy (A_l, R_l, P_l = R_l^T) — the object that
replaces the `R` and `A` lists of openmg.mgSolve (openmg/__init__.py:103-109).
Everything here is a thin ctypes wrapper over include/omg_b200.h."""
import ctypes

import numpy as np
import scipy.sparse as sparse

from . import _lib
from ._lib import SMOOTHERS, check, f64, i32, i64


class BandMatrix(object):
    """A_0 = diag*I + sum_k coeff_k (S^{+o_k} + S^{-o_k}) truncated at the global ends:
    the closed form of operators.poisson (openmg/operators.py:191-256),
    carried as a descriptor so that 512^3+ problems never materialise a host CSR."""

    def __init__(self, n, diag, offsets, coeffs, problemShape=None):
        self.n = int(n)
        self.shape = (self.n, self.n)
        self.diag = float(diag)
        self.offsets = [int(o) for o in offsets]
        self.coeffs = [float(c) for c in coeffs]
        self.problemShape = problemShape
        self.dtype = np.dtype(np.float64)

    def tocsr(self):
        diags, offs = [np.full(self.n, self.diag)], [0]
        for o, c in zip(self.offsets, self.coeffs):
            if o < self.n:
                diags += [np.full(self.n - o, c), np.full(self.n - o, c)]
                offs += [o, -o]
        A = sparse.diags(diags, offs, shape=self.shape, format='csr')
        A.sum_duplicates()
        A.sort_indices()
        return sparse.csr_matrix(A)

    def toarray(self):
        return self.tocsr().toarray()

    def _operator(self):
        op = getattr(self, "_op", None)
        if op is None:
            op = self._op = Operator(self)
        return op

    def dot(self, x):
        x = np.asarray(x, dtype=np.float64)
        return self._operator().matvec(x.ravel()).reshape(x.shape)

    __mul__ = dot
    __matmul__ = dot


def as_csr(A):
    """Canonical int32/float64 CSR (sorted, duplicates summed) of any input matrix."""
    if isinstance(A, BandMatrix):
        A = A.tocsr()
    if not sparse.issparse(A):
        A = sparse.csr_matrix(np.asarray(A, dtype=np.float64))
    A = sparse.csr_matrix(A, dtype=np.float64)
    if A.shape[0] != A.shape[1]:
        raise ValueError("coefficient matrix must be square, got %r" % (A.shape,))
    if not A.has_canonical_format:
        A = A.copy()
        A.sum_duplicates()
    if A.indices.dtype != np.int32 or A.indptr.dtype != np.int32:
        if A.nnz >= 2 ** 31:
            raise NotImplementedError("matrices with >= 2^31 stored entries need the BandMatrix path")
        A = sparse.csr_matrix((A.data, A.indices.astype(np.int32), A.indptr.astype(np.int32)), shape=A.shape)
    return A


def _vec(v, n, name):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
    if a.size != n:
        raise ValueError("%s has %d entries, expected %d" % (name, a.size, n))
    return a


class _Handle(object):
    """Owner of an omg_hierarchy*."""

    def __init__(self):
        self._h = ctypes.c_void_p()
        self._L = _lib.lib()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.omg_hierarchy_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---- introspection
    @property
    def nlevels(self):
        n = ctypes.c_int32()
        check(self._L.omg_level_count(self._h, ctypes.byref(n)))
        return n.value

    def level_info(self, level):
        n, nnzA, nnzR, nexc = (ctypes.c_int64() for _ in range(4))
        kind = ctypes.c_int32()
        check(self._L.omg_level_info(self._h, level, ctypes.byref(n), ctypes.byref(nnzA), ctypes.byref(nnzR),
                                     ctypes.byref(kind), ctypes.byref(nexc)))
        return {"n": n.value, "nnzA": nnzA.value, "nnzR": nnzR.value, "kind": _lib.KINDS[kind.value],
                "nexc": nexc.value}

    def local_range(self, level=0):
        """(row0, nloc, is_slab): the rows of `level` this rank owns (everything on one GPU)."""
        r0, nl = ctypes.c_int64(), ctypes.c_int64()
        slab = ctypes.c_int32()
        check(self._L.omg_level_partition(self._h, level, ctypes.byref(r0), ctypes.byref(nl), ctypes.byref(slab)))
        return r0.value, nl.value, bool(slab.value)

    def level_band(self, level):
        diag = ctypes.c_double()
        nb = ctypes.c_int32()
        offs = np.zeros(16, np.int64)
        coef = np.zeros(16, np.float64)
        check(self._L.omg_level_band(self._h, level, ctypes.byref(diag), ctypes.byref(nb), i64(offs), f64(coef)))
        return diag.value, offs[:nb.value].copy(), coef[:nb.value].copy()

    def export_A(self, level):
        info = self.level_info(level)
        n, nnz = info["n"], info["nnzA"]
        indptr = np.empty(n + 1, np.int32)
        indices = np.empty(max(nnz, 1), np.int32)
        data = np.empty(max(nnz, 1), np.float64)
        check(self._L.omg_level_export_A(self._h, level, i32(indptr), i32(indices), f64(data)))
        return sparse.csr_matrix((data[:nnz], indices[:nnz], indptr), shape=(n, n))

    def export_R(self, level):
        info = self.level_info(level)
        nf = info["n"]
        nc = self.level_info(level + 1)["n"]
        nnz = info["nnzR"]
        indptr = np.empty(nc + 1, np.int32)
        indices = np.empty(max(nnz, 1), np.int32)
        data = np.empty(max(nnz, 1), np.float64)
        check(self._L.omg_level_export_R(self._h, level, i32(indptr), i32(indices), f64(data)))
        return sparse.csr_matrix((data[:nnz], indices[:nnz], indptr), shape=(nc, nf))

    def setup_times(self):
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        check(self._L.omg_setup_times(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"upload_ms": a.value, "galerkin_ms": b.value, "coarse_factor_ms": c.value}

    # ---- unit operations (one reference operation each)
    def n(self, level=0):
        return self.level_info(level)["n"]

    def smooth(self, level, b, x, sweeps, smoother="jacobi", omega=0.8):
        n = self.n(level)
        bb, xx = _vec(b, n, "b"), _vec(x, n, "x").copy()
        check(self._L.omg_smooth(self._h, level, f64(bb), f64(xx), int(sweeps), SMOOTHERS[smoother], float(omega)))
        return xx

    def smooth_to_threshold(self, level, b, x, threshold, smoother="jacobi", omega=0.8, max_sweeps=10 ** 7):
        n = self.n(level)
        bb, xx = _vec(b, n, "b"), _vec(x, n, "x").copy()
        it, nv = ctypes.c_int32(), ctypes.c_double()
        check(self._L.omg_smooth_to_threshold(self._h, level, f64(bb), f64(xx), float(threshold), int(max_sweeps),
                                              SMOOTHERS[smoother], float(omega), ctypes.byref(it), ctypes.byref(nv)))
        return xx, it.value, nv.value

    def residual_restrict(self, level, b, x):
        n = self.n(level)
        bb, xx = _vec(b, n, "b"), _vec(x, n, "x")
        rc = np.empty(self.n(level + 1), np.float64)
        check(self._L.omg_residual_restrict(self._h, level, f64(bb), f64(xx), f64(rc)))
        return rc

    def smooth_residual_restrict(self, level, b, x, sweeps, smoother="jacobi", omega=0.8):
        """The descent step of a cycle on one level (openmg/__init__.py:201,209-210): returns (smoothed x, R (b - A x))."""
        n = self.n(level)
        bb, xx = _vec(b, n, "b"), _vec(x, n, "x").copy()
        rc = np.empty(self.n(level + 1), np.float64)
        check(self._L.omg_smooth_residual_restrict(self._h, level, f64(bb), f64(xx), int(sweeps), SMOOTHERS[smoother],
                                                   float(omega), f64(rc)))
        return xx, rc

    def prolong_correct(self, level, ec, x):
        xx = _vec(x, self.n(level), "x").copy()
        ee = _vec(ec, self.n(level + 1), "e")
        check(self._L.omg_prolong_correct(self._h, level, f64(ee), f64(xx)))
        return xx

    def prolong_correct_smooth(self, level, b, ec, x, sweeps, smoother="jacobi", omega=0.8):
        n = self.n(level)
        bb, xx = _vec(b, n, "b"), _vec(x, n, "x").copy()
        ee = _vec(ec, self.n(level + 1), "e")
        check(self._L.omg_prolong_correct_smooth(self._h, level, f64(bb), f64(ee), f64(xx), int(sweeps),
                                                 SMOOTHERS[smoother], float(omega)))
        return xx

    def coarse_solve(self, b):
        n = self.n(self.nlevels - 1)
        bb = _vec(b, n, "b")
        x = np.empty(n, np.float64)
        check(self._L.omg_coarse_solve(self._h, f64(bb), f64(x)))
        return x

    def residual_norm(self, level, b, x):
        n = self.n(level)
        bb, xx = _vec(b, n, "b"), _vec(x, n, "x")
        out = ctypes.c_double()
        check(self._L.omg_residual_norm(self._h, level, f64(bb), f64(xx), ctypes.byref(out)))
        return out.value

    def residual(self, level, b, x):
        n = self.n(level)
        bb, xx = _vec(b, n, "b"), _vec(x, n, "x")
        r = np.empty(n, np.float64)
        check(self._L.omg_residual(self._h, level, f64(bb), f64(xx), f64(r)))
        return r

    def matvec(self, x, level=0):
        xx = _vec(x, self.n(level), "x")
        y = np.empty_like(xx)
        check(self._L.omg_matvec(self._h, level, f64(xx), f64(y)))
        return y


class Operator(_Handle):
    """One matrix on the device (no hierarchy): backs the standalone solvers.* / tools.* calls."""

    def __init__(self, A, factor=False, flags=0):
        _Handle.__init__(self)
        A = as_csr(A)
        self._keep = A
        if factor:
            flags |= _lib.FLAG_FACTOR
        check(self._L.omg_operator_create_csr(ctypes.byref(self._h), A.shape[0], i32(A.indptr), i32(A.indices),
                                              f64(A.data), int(flags)))


class Hierarchy(_Handle):
    """Hierarchy(A_in, problemShape, coarsestLevel, minSize=8, flags=0)

    `coarsestLevel` and `minSize` follow operators.restrictionList
    (openmg/operators.py:92-141); len(R) = nlevels - 1."""

    def __init__(self, A_in, problemShape, coarsestLevel, minSize=8, flags=0):
        _Handle.__init__(self)
        shape = np.ascontiguousarray(np.array([int(s) for s in problemShape], dtype=np.int64))
        self.problemShape = tuple(int(s) for s in problemShape)
        if isinstance(A_in, BandMatrix):
            offs = np.ascontiguousarray(np.array(A_in.offsets, dtype=np.int64))
            coef = np.ascontiguousarray(np.array(A_in.coeffs, dtype=np.float64))
            check(self._L.omg_hierarchy_create_band(ctypes.byref(self._h), len(shape), i64(shape), int(coarsestLevel),
                                                    int(minSize), A_in.n, A_in.diag, len(offs), i64(offs), f64(coef),
                                                    int(flags)))
        else:
            A = as_csr(A_in)
            check(self._L.omg_hierarchy_create_csr(ctypes.byref(self._h), len(shape), i64(shape), int(coarsestLevel),
                                                   int(minSize), A.shape[0], i32(A.indptr), i32(A.indices),
                                                   f64(A.data), int(flags)))

    def solve(self, b, x0=None, pre=1, post=0, smoother="rbgs", omega=0.8, cycles=0, threshold=0.1,
              want_history=False, out=None):
        """The V-cycle loop of mgSolve (openmg/__init__.py:112-138).  Returns (x, cycles_done, norm, history)."""
        n = self.n(0)
        bb = _vec(b, n, "b")
        if out is None:
            out = np.zeros(n, np.float64)    # multi-GPU: only this rank's rows are filled (see openmg_b200.dist)
        has_initial = 0
        if x0 is not None:
            out[:] = _vec(x0, n, "x0")
            has_initial = 1
        cap = max(int(cycles), 1) if cycles > 0 else 100000
        hist = np.zeros(cap, np.float64) if (want_history or threshold > 0) else None
        done, norm = ctypes.c_int32(), ctypes.c_double()
        check(self._L.omg_solve(self._h, f64(bb), f64(out), has_initial, int(pre), int(post), SMOOTHERS[smoother],
                                float(omega), int(cycles), float(threshold), ctypes.byref(done), ctypes.byref(norm),
                                f64(hist) if hist is not None else None, cap if hist is not None else 0))
        return out, done.value, norm.value, (hist[:done.value].copy() if hist is not None else None)

    def solve_stats(self):
        """{'host_syncs': stream synchronisations of the last solve's cycle loop, 'coarse_defect': max |A_L Ainv - I|}"""
        syncs, defect = ctypes.c_int64(), ctypes.c_double()
        check(self._L.omg_solve_stats(self._h, ctypes.byref(syncs), ctypes.byref(defect)))
        return {"host_syncs": syncs.value, "coarse_defect": defect.value}

    def cycle(self, b, x0=None, level=0, pre=1, post=0, smoother="rbgs", omega=0.8):
        """One mgCycle entered at `level` (openmg/__init__.py:151-236). Returns (uOut, norm)."""
        n = self.n(level)
        bb = _vec(b, n, "b")
        x = np.zeros(n, np.float64) if x0 is None else _vec(x0, n, "x0").copy()
        norm = ctypes.c_double()
        check(self._L.omg_cycle(self._h, int(level), f64(bb), f64(x), 0 if x0 is None else 1, int(pre), int(post),
                                SMOOTHERS[smoother], float(omega), ctypes.byref(norm)))
        return x, norm.value

    # ---- multi-GPU: per-rank slices instead of global host vectors
    def _virt(self, local_arr, level=0):
        """Pointer p with p[row0 + i] == local_arr[i]: what the C-ABI (which takes GLOBAL host
        vectors and touches only this rank's rows) needs when only the local slice exists."""
        row0, nloc, _ = self.local_range(level)
        a = local_arr
        if a.dtype != np.float64 or not a.flags.c_contiguous or a.size != nloc:
            raise ValueError("local slice must be a contiguous float64 array of %d entries" % nloc)
        return ctypes.cast(ctypes.c_void_p(a.ctypes.data - 8 * row0), _lib.c_f64p)

    def matvec_local(self, x_loc, level=0):
        y = np.empty_like(x_loc)
        check(self._L.omg_matvec(self._h, level, self._virt(x_loc, level), self._virt(y, level)))
        return y

    def set_rhs_local(self, b_loc):
        check(self._L.omg_set_rhs(self._h, self._virt(b_loc)))

    def solve_local(self, b_loc, out_loc, pre=1, post=1, smoother="jacobi", omega=0.8, cycles=1, threshold=0.0):
        """omg_solve on this rank's rows only (zero initial iterate). Returns (cycles_done, global norm)."""
        done, norm = ctypes.c_int32(), ctypes.c_double()
        check(self._L.omg_solve(self._h, self._virt(b_loc), self._virt(out_loc), 0, int(pre), int(post),
                                SMOOTHERS[smoother], float(omega), int(cycles), float(threshold),
                                ctypes.byref(done), ctypes.byref(norm), None, 0))
        return done.value, norm.value

    # ---- device-resident benchmarking
    def set_rhs(self, b):
        bb = _vec(b, self.n(0), "b")
        check(self._L.omg_set_rhs(self._h, f64(bb)))

    def bench_cycles(self, ncycles, pre=1, post=1, smoother="jacobi", omega=0.8, with_norm=False):
        ms = ctypes.c_float()
        launches = ctypes.c_int64()
        check(self._L.omg_bench_cycles(self._h, int(pre), int(post), SMOOTHERS[smoother], float(omega), int(ncycles),
                                       1 if with_norm else 0, ctypes.byref(ms), ctypes.byref(launches)))
        return ms.value, launches.value

    def profile_cycle(self, reps=3, pre=1, post=1, smoother="jacobi", omega=0.8):
        """Per-kernel CUDA-event timings (list of dicts) of direct-launched cycles."""
        import json
        buf = ctypes.create_string_buffer(1 << 16)
        check(self._L.omg_profile_cycle(self._h, int(pre), int(post), SMOOTHERS[smoother], float(omega), int(reps),
                                        buf, 1 << 16))
        return json.loads(buf.value.decode())

    def solution(self):
        x = np.empty(self.n(0), np.float64)
        check(self._L.omg_get_solution(self._h, f64(x)))
        return x

    def current_norm(self):
        out = ctypes.c_double()
        check(self._L.omg_current_norm(self._h, ctypes.byref(out)))
        return out.value
