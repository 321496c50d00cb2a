"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.

CPU restatement (numpy/scipy, Python 3) of tsbertalan/openmg's V-cycle hot
path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module; `openmg_b200` never does.

Every function cites the reference file:line it restates (paths relative to
the reference root, i.e. `openmg/...`).

Pinning status
--------------
* Everything the reference itself implements (restriction, restrictionList,
  coeffecientList, the Poisson generators, lexicographic Gauss-Seidel,
  coarseSolve, mgCycle control flow, mgSolve stop logic) is PINNED: it is
  checked against outputs of the reference's own code (run in the build
  container through tools/ref_shim.py) committed as tests/golden/*.npz by
  tools/make_golden.py, and live against the reference when /root/reference
  is present (tests/test_oracle_vs_reference.py).
* The reference has no numeric golden vectors of its own (unseeded RNG,
  inequality asserts only — openmg/tests.py:81,515,531,570); the goldens above
  are the pin.
* Weighted Jacobi and two-colour Gauss-Seidel do not exist in the reference
  (openmg/solvers.py:28-29 offers lexicographic GS only).  Their definitions
  live here; their V-cycle goldens were produced by the REFERENCE's own
  `mgCycle` (openmg/__init__.py:151-236) with its documented plug-in point
  `openmg.smooth` replaced by the smoothers below, so the cycle control flow
  around them is pinned, the smoother arithmetic itself is "parity unpinned"
  (defined by this file).
"""
import ctypes
import os

import numpy as np
import scipy.sparse as sparse
import scipy.sparse.linalg as splinalg

# --------------------------------------------------------------------------
# tools  (openmg/tools.py)
# --------------------------------------------------------------------------


def product(iterableThing):
    """openmg/tools.py:56-60."""
    out = 1
    for thing in iterableThing:
        out *= thing
    return out


def flexibleMmult(x, y):
    """openmg/tools.py:18-26: np.dot for dense·dense, overloaded `*` otherwise."""
    if (not sparse.issparse(x)) and (not sparse.issparse(y)):
        return np.dot(x, y)
    return x * y


def getresidual(b, A, x, N):
    """openmg/tools.py:12-15: (N,1) column b - A x."""
    return b.reshape((N, 1)) - flexibleMmult(A, x.reshape((N, 1)))


def dictAddNoClobber(dictionary, key, value):
    """openmg/tools.py:43-53."""
    if key not in dictionary:
        dictionary[key] = value
    return dictionary


def dictUpdateNoClobber(updateDict, targetDict):
    """openmg/tools.py:29-40."""
    for key, value in updateDict.items():
        dictAddNoClobber(targetDict, key, value)
    return targetDict


# --------------------------------------------------------------------------
# operators  (openmg/operators.py)
# --------------------------------------------------------------------------

def _restriction_offsets(shape):
    """Column offsets written per coarse row, openmg/operators.py:75-84.
    NX = shape[0], NY = shape[1] (openmg/operators.py:46-49)."""
    alpha = len(shape)
    NX = shape[0]
    offs = [0, 1]
    if alpha >= 2:
        offs += [NX, NX + 1]
        if alpha == 3:
            NY = shape[1]
            offs += [NX * NY, NX * NY + 1, NX * NY + NX, NX * NY + NX + 1]
    return offs


def restriction_loop(shape, dense=False):
    """Literal restatement of openmg/operators.py:15-89 (Python loop into a
    lil_matrix).  Slow; used to validate `restriction` on small shapes."""
    alpha = len(shape)
    NX = shape[0]
    if alpha >= 2:
        NY = shape[1]
    N = product(shape)
    n = N // (2 ** alpha)                                   # :52 (py2 int division)
    if n in (0, 1):                                          # :53-56
        raise ValueError('New restriction matrix would have shape ' + str((n, N))
                         + '. ' + 'Coarse set would have %d point(s)! ' % n +
                         'Try a larger problem or fewer gridLevels.')
    R = np.zeros((n, N)) if dense else sparse.lil_matrix((n, N))
    if alpha == 1:                                           # :63-71
        coarseColumns = np.arange(N).reshape(shape)[::2].ravel()
    elif alpha == 2:
        coarseColumns = np.arange(N).reshape(shape)[::2, ::2].ravel()
    elif alpha == 3:
        coarseColumns = np.arange(N).reshape(shape)[::2, ::2, ::2].ravel()
    else:
        raise ValueError("restriction(): Greater than 3 dimensions is not"
                         "implemented. (shape was" + str(shape) + " .)")
    each = 1.0 / (2 ** alpha)                                # :73
    for r, c in zip(range(n), coarseColumns):                # :74-84
        R[r, c] = each
        R[r, c + 1] = each
        if alpha >= 2:
            R[r, c + NX] = each
            R[r, c + NX + 1] = each
            if alpha == 3:
                R[r, c + NX * NY] = each
                R[r, c + NX * NY + 1] = each
                R[r, c + NX * NY + NX] = each
                R[r, c + NX * NY + NX + 1] = each
    return R if dense else R.tocsr()                         # :86-89


def restriction(shape, dense=False):
    """Closed form of openmg/operators.py:15-89: same matrix,
    same exceptions, no Python loop over rows.  Validated against
    `restriction_loop` and the reference goldens."""
    shape = tuple(int(s) for s in shape)
    alpha = len(shape)
    N = product(shape)
    n = N // (2 ** alpha)
    if n in (0, 1):
        raise ValueError('New restriction matrix would have shape ' + str((n, N))
                         + '. ' + 'Coarse set would have %d point(s)! ' % n +
                         'Try a larger problem or fewer gridLevels.')
    if alpha > 3:
        raise ValueError("restriction(): Greater than 3 dimensions is not"
                         "implemented. (shape was" + str(shape) + " .)")
    sl = (slice(None, None, 2),) * alpha
    cc = np.arange(N, dtype=np.int64).reshape(shape)[sl].ravel()[:n]   # zip truncation :74
    nrows = cc.size                      # may be < n only if the coarse set is smaller
    offs = np.array(_restriction_offsets(shape), dtype=np.int64)
    cols = cc[:, None] + offs[None, :]
    if cols.size and cols.max() >= N:
        # lil_matrix.__setitem__ raises IndexError in the reference (:75-84)
        raise IndexError('column index out of range in restriction for shape ' + str(shape))
    each = 1.0 / (2 ** alpha)
    rows = np.repeat(np.arange(nrows, dtype=np.int64), offs.size)
    # duplicates inside a row collapse by overwrite in the lil_matrix -> value stays `each`
    key = rows * N + cols.ravel()
    key = np.unique(key)
    rows_u, cols_u = key // N, key % N
    R = sparse.csr_matrix((np.full(key.size, each), (rows_u, cols_u)), shape=(n, N))
    R.sort_indices()
    if dense:
        return R.toarray()
    return R


def interpolation(shape, dense=False):
    """Prolongation = R^T with the same weights, openmg/__init__.py:214."""
    R = restriction(shape, dense=dense)
    return R.T if dense else R.T.tocsr()


def restrictionList(problemShape, coarsestLevel, minSize, dense=False, verbose=False):
    """openmg/operators.py:92-141: first R unconditional (:130-132), later ones
    while level < coarsestLevel and n > minSize (:133-140)."""
    levels = coarsestLevel + 1
    R = []
    level = 0
    nextR = restriction(tuple(np.array(problemShape) // (2 ** level)), dense=dense)
    R.append(nextR)
    while level < levels - 1:
        level += 1
        nextR = restriction(tuple(np.array(problemShape) // (2 ** level)), dense=dense)
        nNext = nextR.shape[0]
        if nNext <= minSize:
            break
        R.append(nextR)
    return R


def coeffecientList(A_in, R, dense=False, verbose=False):
    """openmg/operators.py:144-188: A[0]=csr(A_in) (:178) or dense (:172-176);
    A[l] = (R[l-1]*A[l-1])*R[l-1].T (:184-186)."""
    levels = len(R) + 1
    A = list(range(levels))
    if dense:
        A[0] = A_in.todense() if sparse.issparse(A_in) else A_in
    else:
        A[0] = sparse.csr_matrix(A_in)
    for level in range(1, levels):
        A[level] = flexibleMmult(flexibleMmult(R[level - 1], A[level - 1]), R[level - 1].T)
    return A


def canonical_csr(M):
    """Sorted-index, duplicate-free, explicit-zero-free CSR copy (int32/float64),
    the form in which patterns are compared."""
    M = sparse.csr_matrix(M).copy()
    M.sum_duplicates()
    M.eliminate_zeros()
    M.sort_indices()
    return M


def poisson_dense(shape):
    """Literal restatement of the dense generators openmg/operators.py:206-256."""
    if isinstance(shape, int):
        shape = (shape,)
    if len(shape) == 1:                                    # poisson1D :206-218
        N = shape[0]
        return (np.diag(-np.ones(N - 1), -1) + np.diag(2 * np.ones(N))
                + np.diag(-np.ones(N - 1), 1))
    if len(shape) == 2:                                    # poisson2D :221-241
        NX, NY = shape
        N = NX * NY
        main = np.eye(N) * -4
        oneup = np.eye(N, k=1)                              # :227-233
        twoup = np.eye(N, k=1 + NX)                         # :234-240
        return main + oneup + twoup + oneup.T + twoup.T
    if len(shape) == 3:                                    # poisson3D :244-256
        NX, NY, NZ = shape
        N = NX * NY * NZ
        A = np.zeros((N, N))
        for i in range(N):
            A[i, i] = -6
            for index in (i + 1, i + NX, i + NX * NY):
                if index < N:
                    A[i, index] = 1
        A += A.T
        return A
    raise ValueError('Only 1, 2 or 3 dimensions are allowed.')   # :273


def poisson_bands(shape, sparse_1d=False):
    """(diag, [(offset, coeff), ...]) of the reference's generators, SURVEY §A.2:
    1-D dense (2,-1,[1]) :206-218; 1-D sparse (4,-1,[1]) :191-203;
    2-D (-4,+1,[1,NX+1]) :226-241; 3-D (-12,+1,[1,NX,NX*NY]) :249-255."""
    if isinstance(shape, int):
        shape = (shape,)
    if len(shape) == 1:
        return (4.0, [(1, -1.0)]) if sparse_1d else (2.0, [(1, -1.0)])
    if len(shape) == 2:
        return -4.0, [(1, 1.0), (shape[0] + 1, 1.0)]
    if len(shape) == 3:
        NX, NY = shape[0], shape[1]
        offs = {}
        for o in (1, NX, NX * NY):            # coincident offsets add (dense `A += A.T`
            offs[o] = 1.0                     # overwrites, it does not accumulate: :252-254)
        return -12.0, sorted(offs.items())
    raise ValueError('Only 1, 2 or 3 dimensions are allowed.')


def poisson(shape, sparse=False):
    """openmg/operators.py:259-279 (`poisson = poissonnd`).  `sparse=True` for
    1-D returns the reference's diag-4 matrix (:191-203); for 2-D/3-D the
    reference raises NotImplementedError (:224,247) and so does this."""
    import scipy.sparse as sp
    if isinstance(shape, int):
        shape = (shape,)
    if len(shape) > 3 or len(shape) == 0:
        raise ValueError('Only 1, 2 or 3 dimensions are allowed.')
    if sparse:
        if len(shape) > 1:
            raise NotImplementedError("Sparse poisson for alpha>1 is not yet implemented.")
        return poisson_csr(shape, sparse_1d=True)
    return poisson_dense(shape)


def poisson_csr(shape, sparse_1d=False):
    """Closed-form CSR equal (pattern and values) to
    csr_matrix(reference.poisson(shape)) — SURVEY.md §A.2."""
    if isinstance(shape, int):
        shape = (shape,)
    N = product(shape)
    d, bands = poisson_bands(shape, sparse_1d=sparse_1d)
    diags = [np.full(N, d)]
    offs = [0]
    for o, c in bands:
        if o < N:
            diags += [np.full(N - o, c), np.full(N - o, c)]
            offs += [o, -o]
    A = sparse.diags(diags, offs, shape=(N, N), format='csr')
    A.sort_indices()
    return A


# --------------------------------------------------------------------------
# solvers  (openmg/solvers.py)
# --------------------------------------------------------------------------

def coarseSolve(A, b):
    """openmg/solvers.py:16-26."""
    if sparse.issparse(A):
        toreturn = splinalg.spsolve(sparse.csc_matrix(A), np.asarray(b).ravel())
    else:
        toreturn = np.linalg.solve(A, b)
    return np.ravel(toreturn)


def gaussSeidel(A, b, x, iterations=None, threshold=None, verbose=False):
    """Literal restatement of openmg/solvers.py:34-75 — lexicographic forward
    Gauss-Seidel, in place, pure-Python loop over rows (that loop IS the
    reference's speed; see `gaussSeidel_c` for the compiled restatement)."""
    if iterations is None and threshold is None:
        iterations = 1
    N = x.size
    bf = np.asarray(b).ravel()

    def stop(iteration, x):
        iterStatus = threshStatus = False
        if iterations is not None:
            iterStatus = (iteration >= iterations)
        if threshold is not None:
            norm = np.linalg.norm(getresidual(bf, A, x, N))
            threshStatus = (norm < threshold)
        return iterStatus or threshStatus

    iteration = 0
    stopping = stop(iteration, x)
    xf = x.reshape(-1)  # view; in-place like the reference
    while not stopping:
        if sparse.issparse(A):
            indptr, indices, data = A.indptr, A.indices, A.data
            for i in range(N):
                rs, re = indptr[i], indptr[i + 1]
                Aix = np.dot(data[rs:re], xf[indices[rs:re]])        # :63-65
                xf[i] = xf[i] + (bf[i] - Aix) / A[i, i]               # :68
        else:
            Ad = np.asarray(A)
            for i in range(N):
                xf[i] = xf[i] + (bf[i] - np.dot(Ad[i, :], xf)) / Ad[i, i]   # :70-71
        iteration += 1
        stopping = stop(iteration, x)
    return x


_CLIB = None


def _clib():
    """Compiled restatement (oracle/csrc/oracle_kernels.c -> oracle/_build/)."""
    global _CLIB
    if _CLIB is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "_build", "liboracle_kernels.so")
        if not os.path.exists(path):
            from . import build as _b
            _b.build()
        lib = ctypes.CDLL(path)
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
        f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
        lib.oracle_gs_sweeps.argtypes = [ctypes.c_int64, i32p, i32p, f64p, f64p, f64p, ctypes.c_int]
        lib.oracle_gs_sweeps.restype = ctypes.c_int
        lib.oracle_jacobi_sweeps.argtypes = [ctypes.c_int64, i32p, i32p, f64p, f64p, f64p, f64p,
                                             ctypes.c_double, ctypes.c_int]
        lib.oracle_jacobi_sweeps.restype = ctypes.c_int
        lib.oracle_rbgs_sweeps.argtypes = [ctypes.c_int64, i32p, i32p, f64p, f64p, f64p, f64p, u8p,
                                           ctypes.c_int]
        lib.oracle_rbgs_sweeps.restype = ctypes.c_int
        _CLIB = lib
    return _CLIB


def _csr32(A):
    A = sparse.csr_matrix(A)
    return (np.ascontiguousarray(A.indptr, np.int32), np.ascontiguousarray(A.indices, np.int32),
            np.ascontiguousarray(A.data, np.float64))


def gaussSeidel_c(A, b, x, iterations=1):
    """Same row loop as openmg/solvers.py:56-68, compiled C (one thread)."""
    indptr, indices, data = _csr32(A)
    bf = np.ascontiguousarray(np.asarray(b).ravel(), np.float64)
    xf = x.reshape(-1)
    assert xf.flags.c_contiguous and xf.dtype == np.float64
    rc = _clib().oracle_gs_sweeps(xf.size, indptr, indices, data, bf, xf, int(iterations))
    if rc != 0:
        raise ZeroDivisionError("zero/missing diagonal in Gauss-Seidel")
    return x


def jacobi(A, b, x, iterations=1, omega=0.8, verbose=False):
    """Weighted Jacobi, x <- x + omega*(b - A x)/diag(A); in place, returns x
    (same calling convention as openmg/solvers.py:28 `smooth`).  Not in the
    reference; update formula is the per-row GS update openmg/solvers.py:68
    applied simultaneously to all rows with weight omega."""
    A = sparse.csr_matrix(A)
    d = A.diagonal()
    bf = np.asarray(b).ravel()
    xf = x.reshape(-1)
    for _ in range(int(iterations)):
        xf += omega * (bf - A.dot(xf)) / d
    return x


def level_shape(problemShape, level):
    """Grid shape on `level`: openmg/operators.py:131,136."""
    return tuple(int(s) for s in (np.array(problemShape) // (2 ** level)))


def colouring(problemShape, level, n):
    """Two-colouring of the n rows of level `level` (uint8 0/1).
    Rule: flat-index parity for 1-D/2-D level 0 (exact
    red-black there), C-order grid-coordinate parity elsewhere.  The leading
    coordinate is taken without wrap so any n is covered."""
    shape = level_shape(problemShape, level)
    alpha = len(shape)
    i = np.arange(n, dtype=np.int64)
    if alpha == 1 or (alpha == 2 and level == 0):
        return (i & 1).astype(np.uint8)
    s = np.zeros(n, dtype=np.int64)
    rem = i
    for d in range(alpha - 1, 0, -1):
        sd = max(int(shape[d]), 1)
        s += rem % sd
        rem = rem // sd
    s += rem
    return (s & 1).astype(np.uint8)


def rbgs(A, b, x, iterations=1, colours=None, verbose=False):
    """Two-colour Gauss-Seidel ("red-black" where the graph is bipartite).
    One sweep = for c in (0,1): x_i <- x_i + (b_i - A_i·x)/a_ii for all rows of
    colour c simultaneously, using x as it stood before the half-sweep (so
    same-colour couplings are lagged).  Row update = openmg/solvers.py:68.
    In place, returns x."""
    A = sparse.csr_matrix(A)
    d = A.diagonal()
    bf = np.asarray(b).ravel()
    xf = x.reshape(-1)
    if colours is None:
        colours = (np.arange(xf.size) & 1).astype(np.uint8)
    masks = [colours == 0, colours == 1]
    for _ in range(int(iterations)):
        for m in masks:
            r = (bf - A.dot(xf)) / d
            xf[m] += r[m]
    return x


def make_smoother(kind, problemShape=None, omega=0.8, fast=True):
    """Returns smooth(A, b, x, iterations, verbose=False, level=None) for
    kind in {'gs','jacobi','rbgs'}; `level` selects the colouring for 'rbgs'."""
    if kind == 'gs':
        def smooth(A, b, x, iterations, verbose=False, level=None):
            if fast and sparse.issparse(A):
                return gaussSeidel_c(A, b, x, iterations)
            return gaussSeidel(A, b, x, iterations=iterations)
    elif kind == 'jacobi':
        def smooth(A, b, x, iterations, verbose=False, level=None):
            return jacobi(A, b, x, iterations, omega)
    elif kind == 'rbgs':
        def smooth(A, b, x, iterations, verbose=False, level=None):
            n = x.size
            col = colouring(problemShape, level, n) if (problemShape is not None and level is not None) \
                else (np.arange(n) & 1).astype(np.uint8)
            return rbgs(A, b, x, iterations, col)
    else:
        raise ValueError("unknown smoother %r" % (kind,))
    return smooth


# --------------------------------------------------------------------------
# cycle driver  (openmg/__init__.py)
# --------------------------------------------------------------------------

defaults = {                      # openmg/__init__.py:16-27
    'problemShape': (200,),
    'gridLevels': 2,
    'verbose': False,
    'threshold': 0.1,
    'cycles': 0,
    'preIterations': 1,
    'postIterations': 0,
    'dense': False,
    'giveInfo': False,
    'minSize': 8,
}


def mgCycle(A, b, level, R, parameters, initial=None, smooth=None, norms=None):
    """openmg/__init__.py:151-236.  `smooth(A,b,x,iterations,level=...)` is the
    plug-in point (the reference resolves the module-global `smooth` at
    :201,218).  The wasted `R*b` of :205 is not recomputed (only its length is
    used there)."""
    if smooth is None:
        smooth = make_smoother('gs')
    b = np.asarray(b, dtype=np.float64).ravel()
    if initial is None:
        initial = np.zeros((b.size,))                         # :191-192
    N = b.size
    if level < parameters['coarsestLevel']:                  # :199
        uApx = smooth(A[level], b, initial, parameters['preIterations'], level=level)   # :201
        NH = R[level].shape[0]                                # :205-206
        residual = np.asarray(getresidual(b, A[level], uApx, N)).ravel()      # :209
        coarseResidual = np.asarray(flexibleMmult(R[level], residual.reshape((N, 1)))).reshape((NH,))  # :210
        coarseCorrection = mgCycle(A, coarseResidual, level + 1, R, parameters, smooth=smooth)[0]      # :213
        correction = np.asarray(flexibleMmult(R[level].transpose(),
                                              coarseCorrection.reshape((NH, 1)))).reshape((N,))        # :214
        if parameters['postIterations'] > 0:                  # :216-222
            uOut = smooth(A[level], b, uApx + correction, parameters['postIterations'], level=level)
        else:
            uOut = uApx + correction                          # :224
        norm = np.linalg.norm(getresidual(b, A[level], uOut, N))               # :227
    else:
        norm = 0                                              # :232
        uOut = coarseSolve(A[level], b.reshape((N, 1)))       # :234
    return uOut, {'norm': norm}


def mgSolve(A_in, b, parameters, smooth=None):
    """openmg/__init__.py:28-148 (dict mutation :93-96,106; >=1 cycle :112;
    ValueError after the first cycle :118-119; stop rule :120-138).
    Adds info key 'norms' (per-cycle history) — harmless extra."""
    problemShape = parameters['problemShape']
    gridLevels = parameters['gridLevels']
    defaults['coarsestLevel'] = gridLevels - 1
    dictUpdateNoClobber(defaults, parameters)
    if smooth is None:
        smooth = make_smoother(parameters.get('smoother', 'gs'), problemShape,
                               parameters.get('omega', 0.8))
    R = restrictionList(problemShape, parameters['coarsestLevel'], parameters['minSize'],
                        dense=False, verbose=parameters['verbose'])
    parameters['coarsestLevel'] = len(R)
    A = coeffecientList(A_in, R, dense=False, verbose=parameters['verbose'])
    result, infoDict = mgCycle(A, b, 0, R, parameters, smooth=smooth)
    norm = infoDict['norm']
    norms = [norm]
    cycle = 1
    if parameters['threshold'] <= 0 and parameters['cycles'] <= 0:
        raise ValueError("Either parameters['threshold'] or parameters['cycles'] must be > 0.")

    def stop(cycle, norm):
        cycleStop = thresholdStop = False
        if 'cycles' in parameters and parameters['cycles'] > 0:
            if cycle >= parameters['cycles']:
                cycleStop = True
        if 'threshold' in parameters:
            if norm < parameters['threshold'] and parameters['threshold'] > 0:
                thresholdStop = True
        return cycleStop or thresholdStop

    stopping = stop(cycle, norm)
    while not stopping:
        cycle += 1
        result, infoDict = mgCycle(A, b, 0, R, parameters, initial=result, smooth=smooth)
        norm = infoDict['norm']
        norms.append(norm)
        stopping = stop(cycle, norm)
    infoDict['cycle'] = cycle
    infoDict['norm'] = norm
    infoDict['norms'] = norms
    infoDict['R'] = R
    infoDict['A'] = A
    if parameters["giveInfo"]:
        return result, infoDict
    return result


# --------------------------------------------------------------------------
# matrix-free band restatement (scales to 256^3..512^3 on the host for the CPU
# baseline; validated against the CSR path above at small sizes)
# --------------------------------------------------------------------------

def band_matvec(diag, bands, x):
    """y = A x for A = diag*I + sum_k c_k (S^{o_k} + S^{-o_k}), truncated at the
    two global ends."""
    y = diag * x
    n = x.size
    for o, c in bands:
        if o < n:
            y[:-o] += c * x[o:]
            y[o:] += c * x[:-o]
    return y
