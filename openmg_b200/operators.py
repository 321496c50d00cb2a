'''Operator matrices of the reference's `openmg.operators` (openmg/operators.py):
restriction / restrictionList / coeffecientList are built on the device and
exported as scipy CSR; the Poisson generators are the problem source (host,
closed form — bit-identical to the reference's dense builds).'''
import ctypes

import numpy as np
import scipy.sparse

from . import _lib
from . import tools
from .hierarchy import BandMatrix, Hierarchy, as_csr


def restriction(shape, dense=False):
    """Restriction matrix of openmg/operators.py:15-89 (unweighted average over the
    2^alpha-cell aggregates, value 1/2^alpha), built by the device closed form.
    Raises ValueError (0/1 coarse rows, >3-D) and IndexError exactly where the
    reference does."""
    L = _lib.lib()
    shp = np.ascontiguousarray(np.array([int(s) for s in shape], dtype=np.int64))
    n, nnz = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(L.omg_restriction(len(shp), _lib.i64(shp), ctypes.byref(n), ctypes.byref(nnz), None, None, None))
    indptr = np.empty(n.value + 1, np.int32)
    indices = np.empty(nnz.value, np.int32)
    data = np.empty(nnz.value, np.float64)
    _lib.check(L.omg_restriction(len(shp), _lib.i64(shp), ctypes.byref(n), ctypes.byref(nnz), _lib.i32(indptr),
                                 _lib.i32(indices), _lib.f64(data)))
    N = tools.product([int(s) for s in shape])
    R = scipy.sparse.csr_matrix((data, indices, indptr), shape=(n.value, N))
    R.has_sorted_indices = True
    R._omg_shape = tuple(int(s) for s in shape)
    if dense:
        return R.toarray()
    return R


def interpolation(shape, dense=False):
    """Prolongation P = R^T with the same 1/2^alpha weights (openmg/__init__.py:214)."""
    R = restriction(shape, dense=dense)
    return R.T if dense else R.T.tocsr()


def restrictionList(problemShape, coarsestLevel, minSize, dense=False, verbose=False):
    """List of restriction matrices, one per level transition; depth rule of
    openmg/operators.py:128-140 (first R unconditional, later ones while
    level < coarsestLevel and rows > minSize)."""
    if verbose:
        print("Generating restriction matrices; dense=%s" % dense)
    fine = np.array(problemShape)
    R = [restriction(tuple(fine), dense=dense)]                    # the first transition is unconditional
    for level in range(1, coarsestLevel + 1):
        candidate = restriction(tuple(fine // (2 ** level)), dense=dense)
        if candidate.shape[0] <= minSize:                          # coarse grid would be too small: stop above it
            break
        R.append(candidate)
    return _RList(R, tuple(int(s) for s in problemShape))


class _RList(list):
    """A plain list that remembers the problemShape it was built for."""

    def __init__(self, items, problemShape):
        list.__init__(self, items)
        self.problemShape = problemShape


def coeffecientList(A_in, R, dense=False, verbose=False):
    """Galerkin coarse operators A[l] = R[l-1] A[l-1] R[l-1]^T (openmg/operators.py:144-188),
    built on the device; returned in canonical (sorted, zero-free) CSR.  `R` must come
    from restrictionList/restriction of this package (it carries its problemShape)."""
    if verbose:
        print("Generating coefficient matrices; dense=%s ..." % dense, end=' ')
    shape = getattr(R, "problemShape", None)
    if shape is None and len(R) > 0:
        shape = getattr(R[0], "_omg_shape", None)
    if shape is None:
        raise NotImplementedError("coeffecientList needs restriction matrices built by openmg_b200.operators "
                                  "(they carry the problemShape the device closed form uses)")
    h = Hierarchy(A_in, shape, len(R) - 1, minSize=0)
    nlev = h.nlevels
    if nlev != len(R) + 1:
        raise ValueError("restriction list does not match problemShape %r" % (shape,))
    A = [h.export_A(l) for l in range(nlev)]
    if not isinstance(A_in, BandMatrix):
        A[0] = scipy.sparse.csr_matrix(A_in) if not dense else A_in
    if dense:
        A = [a.todense() if scipy.sparse.issparse(a) else a for a in A]
    if verbose:
        print('made %i A matrices' % len(A))
    return A


# ---------------------------------------------------------------------------
# Poisson generators (problem source; openmg/operators.py:191-279)
# ---------------------------------------------------------------------------

def _bands(shape, sparse_1d):
    if len(shape) == 1:
        return (4.0, [(1, -1.0)]) if sparse_1d else (2.0, [(1, -1.0)])     # :196 vs :211-213
    if len(shape) == 2:
        return -4.0, [(1, 1.0), (shape[0] + 1, 1.0)]                        # :226-241
    NX, NY = shape[0], shape[1]
    taps = {}
    for o in (1, NX, NX * NY):                                              # :252-254 (assignment: coincident taps stay 1)
        taps[o] = 1.0
    return -12.0, sorted(taps.items())                                      # -6 doubled by `A += A.T` (:251,255)


def poisson_band(shape, sparse_1d=None):
    """BandMatrix descriptor of poisson(shape): what mgSolve takes for problems whose CSR
    would not fit on the host (512^3 upward).  1-D: `sparse_1d=True` selects the diag-4
    matrix of poisson1Dsparse (openmg/operators.py:191-203), default the diag-2 dense one."""
    if isinstance(shape, int):
        shape = (shape,)
    shape = tuple(int(s) for s in shape)
    if len(shape) > 3 or len(shape) < 1:
        raise ValueError('Only 1, 2 or 3 dimensions are allowed.')
    d, bands = _bands(shape, bool(sparse_1d))
    N = tools.product(shape)
    bands = [(o, c) for o, c in bands if o < N]
    return BandMatrix(N, d, [o for o, _ in bands], [c for _, c in bands], problemShape=shape)


def poisson1Dsparse(N):
    '''Sparse square coefficient matrix for the 1D Poisson equation: tridiag(-1, 4, -1)
    (openmg/operators.py:191-203).'''
    return poisson_band((N,), sparse_1d=True).tocsr()


def poisson1D(shape, sparse=False):
    N = shape[0]
    if sparse:
        return poisson1Dsparse(N)
    if isinstance(N, tuple):
        N = N[0]
    return poisson_band((N,), sparse_1d=False).toarray()


def poisson2D(shape, sparse=False):
    '''Coefficient matrix of the reference's 2D generator (openmg/operators.py:221-241):
    diag -4, +1 at offsets +-1 and +-(NX+1), no row-boundary breaks.  Dense by default;
    `sparse=True` (NotImplementedError in the reference) returns the same matrix as CSR.'''
    A = poisson_band(tuple(shape))
    return A.tocsr() if sparse else A.toarray()


def poisson3D(shape, sparse=False):
    '''Coefficient matrix of the reference's 3D generator (openmg/operators.py:244-256):
    diag -12, +1 at offsets +-1, +-NX, +-NX*NY, no boundary breaks.'''
    A = poisson_band(tuple(shape))
    return A.tocsr() if sparse else A.toarray()


def poissonnd(shape, sparse=False):
    '''Using a 1-, 2-, or 3-element tuple (or an int) for the shape, return the
    reference's Poisson matrix (openmg/operators.py:259-276).'''
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    generators = {1: poisson1D, 2: poisson2D, 3: poisson3D}
    if len(shape) not in generators:
        raise ValueError('Only 1, 2 or 3 dimensions are allowed.')
    M = generators[len(shape)](shape, sparse)
    return scipy.sparse.csr_matrix(M) if sparse else M


poisson = poissonnd
