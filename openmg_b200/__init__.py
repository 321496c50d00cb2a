"""
openmg_b200 — a B200-native (sm_100a CUDA, behind a C-ABI) implementation of the
multigrid V-cycle hot path of tsbertalan/openmg, exposed under the reference's own
Python API:

    import openmg_b200 as openmg
    u = openmg.mgSolve(A, b, {'problemShape': (N,), 'gridLevels': 3, 'cycles': 10})

Same names, argument meaning, return shapes, dict mutation and exceptions as
`openmg/__init__.py`; scipy.sparse / numpy objects in and out.  Additional optional
parameters: 'smoother' in {'rbgs' (default), 'jacobi', 'gs'}, 'omega' (Jacobi weight,
default 0.8), 'hierarchy' (a prebuilt openmg_b200.Hierarchy to reuse).  The key
'iterations' is accepted and ignored, exactly as in the reference.
"""
import numpy as np
import scipy.sparse

from .solvers import smooth, smoothToThreshold, coarseSolve
from . import tools
from . import operators
from . import solvers
from .hierarchy import Hierarchy, Operator, BandMatrix
from . import dist as _dist

tools.poisson = operators.poisson          # north-star alias

# openmg/__init__.py:16-27 — kept identical; mgSolve mutates it like the reference (:95)
defaults = {
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

_ORIGINAL_SMOOTH = smooth


def _smoother_of(parameters):
    s = parameters.get('smoother', solvers.DEFAULT_SMOOTHER)
    if s not in ('rbgs', 'jacobi', 'gs', 'lexgs'):
        raise ValueError("parameters['smoother'] must be 'rbgs', 'jacobi' or 'gs', not %r" % (s,))
    return s, float(parameters.get('omega', solvers.DEFAULT_OMEGA))


class _LazyLevels(object):
    """infoDict['A'] / infoDict['R'] (openmg/__init__.py:142-143): list-like, exported
    from the device on first access (a 512^3 hierarchy is GBs of CSR)."""

    def __init__(self, hierarchy, which, count, dense=False):
        self._h, self._which, self._n, self._dense = hierarchy, which, count, dense
        self._cache = {}

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        if i not in self._cache:
            M = self._h.export_A(i) if self._which == 'A' else self._h.export_R(i)
            self._cache[i] = M.todense() if self._dense else M
        return self._cache[i]

    def __iter__(self):
        return (self[i] for i in range(self._n))

    def __repr__(self):
        return "<%d device-resident %s matrices>" % (self._n, self._which)


def mgSolve(A_in, b, parameters):
    """
    The main externally-usable function (openmg/__init__.py:28-148).

    Parameters
    ----------
        A_in : ndarray, scipy.sparse matrix, or openmg_b200.BandMatrix
            The (square) coefficient matrix.
        b : ndarray
            The right-hand-side vector, shape (N,) or (N,1).
        parameters : dictionary
            Required: problemShape (tuple of ints), gridLevels (int: number of
            restriction transitions; gridLevels=k gives up to k+1 grids).
            Optional (defaults as in the reference): coarsestLevel=gridLevels-1, minSize=8,
            verbose=False, threshold=.1, cycles=0, preIterations=1, postIterations=0,
            dense=False, giveInfo=False.  New: smoother='rbgs', omega=0.8, hierarchy=None.

    Returns
    -------
        result : ndarray of shape (N,)
        infoDict : dict with 'cycle', 'norm', 'R', 'A' (+ 'norms', 'hierarchy'), only if giveInfo.
    """
    problemShape = parameters['problemShape']
    gridLevels = parameters['gridLevels']
    defaults['coarsestLevel'] = gridLevels - 1                     # :95 (mutates the module global)
    tools.dictUpdateNoClobber(defaults, parameters)                # :96 (mutates the caller's dict)

    verbose = parameters['verbose']
    dense = parameters['dense']
    smoother, omega = _smoother_of(parameters)

    if verbose:
        print("Generating restriction matrices; dense=%s" % dense)
    # under torchrun (one process per GPU, torch.distributed initialised) the fine levels are row slabs across the
    # ranks: every rank passes the same global b, uploads only the rows it owns, and gets the full x back
    dist = _dist.active()
    if dist is not None:
        _dist.init_from_torch(dist)
    h = parameters.get('hierarchy', None)
    if h is None:
        h = Hierarchy(A_in, problemShape, parameters['coarsestLevel'], parameters['minSize'])
    nR = h.nlevels - 1
    parameters['coarsestLevel'] = nR                               # :106
    if verbose:
        print("Generating coefficient matrices; dense=%s ..." % dense, end=' ')
        print('made %i A matrices' % (nR + 1))

    N = h.n(0)
    if smooth is not _ORIGINAL_SMOOTH:
        # the reference's smoother plug-in point: `openmg.smooth = f` (:201,218 resolve the
        # module global at call time).  Drive the cycle from Python, one device call per step.
        return _mgSolve_pluggable(h, b, parameters)

    cycles, threshold = parameters['cycles'], parameters['threshold']
    try:
        result, cycle, norm, hist = h.solve(b, None, parameters['preIterations'], parameters['postIterations'],
                                            smoother, omega, cycles, threshold, want_history=bool(verbose))
    except ValueError:
        if verbose and threshold <= 0 and cycles <= 0:
            _verbose_cycle_trace(nR)
        raise
    if dist is not None and h.local_range(0)[2]:
        result = _dist.gather_solution(dist, h, result)
    if verbose:
        for c in range(1, cycle + 1):
            _verbose_cycle_trace(nR)
            if hist is not None:
                print("Residual norm from cycle %d is %f." % (c, hist[c - 1]))
            if c < cycle:
                print('cycle %i < cycles %i' % (c, parameters['cycles']))
    infoDict = {'norm': norm, 'cycle': cycle}
    infoDict['R'] = _LazyLevels(h, 'R', nR, dense)
    infoDict['A'] = _LazyLevels(h, 'A', nR + 1, dense)
    infoDict['A']._hierarchy = h           # mgCycle(info['A'], ...) reuses the device hierarchy
    if hist is not None:
        infoDict['norms'] = hist
    infoDict['hierarchy'] = h
    if verbose:
        print('Returning mgSolve after %i cycle(s) with norm %f' % (cycle, norm))
    if parameters["giveInfo"]:
        return result, infoDict
    return result


mg_solve = mgSolve


def _verbose_cycle_trace(nR):
    for level in range(nR):
        print(level * " " + "calling mgCycle at level %i" % level)          # :208
    print(nR * " " + "direct solving at level %i" % nR)                     # :233


def _stop_rule(parameters, cycle, norm):
    """The reference's stop test (openmg/__init__.py:121-130): a positive `cycles` caps the count, a positive
    `threshold` stops on the residual norm; either suffices.  (omg_solve applies the same rule on the device path.)"""
    ncyc = parameters.get('cycles', 0)
    thr = parameters.get('threshold', 0)
    hit_cap = ncyc > 0 and cycle >= ncyc
    hit_norm = thr > 0 and norm < thr
    return bool(hit_cap or hit_norm)


def _mgSolve_pluggable(h, b, parameters):
    """mgSolve's loop (:112-148) around the Python-driven mgCycle, used when `smooth` was replaced."""
    verbose = parameters['verbose']
    nR = h.nlevels - 1
    A = _LazyLevels(h, 'A', nR + 1)
    R = _LazyLevels(h, 'R', nR)
    A._hierarchy = h
    result, infoDict = mgCycle(A, b, 0, R, parameters)
    norm = infoDict['norm']
    cycle = 1
    if verbose:
        print("Residual norm from cycle %d is %f." % (cycle, norm))
    if parameters['threshold'] <= 0 and parameters['cycles'] <= 0:
        raise ValueError("Either parameters['threshold'] or parameters['cycles'] must be > 0.")

    norms = [norm]
    while not _stop_rule(parameters, cycle, norm):
        if verbose:
            print('cycle %i < cycles %i' % (cycle, parameters['cycles']))
        cycle += 1
        result, infoDict = mgCycle(A, b, 0, R, parameters, initial=result)
        norm = infoDict['norm']
        norms.append(norm)
        if verbose:
            print("Residual norm from cycle %d is %f." % (cycle, norm))
    infoDict = {'cycle': cycle, 'norm': norm, 'R': R, 'A': A, 'norms': np.array(norms), 'hierarchy': h}
    if verbose:
        print('Returning mgSolve after %i cycle(s) with norm %f' % (cycle, norm))
    if parameters["giveInfo"]:
        return result, infoDict
    return result


def _lists_match_hierarchy(h, A, R, shape):
    """Do the caller's A / R lists describe the hierarchy the device holds (the Galerkin products of A[0] with the
    closed-form aggregation restrictions)?  The reference's mgCycle uses whatever lists it is handed
    (openmg/__init__.py:199-214); the device path can only honour lists equal to its own."""
    if len(R) < h.nlevels - 1 or len(A) < h.nlevels:
        return False
    for l in range(h.nlevels - 1):
        if isinstance(R, _LazyLevels):
            break
        want = operators.restriction(tuple(int(s) // (2 ** l) for s in shape))
        got = scipy.sparse.csr_matrix(R[l])
        if got.shape != want.shape or (got != want).nnz != 0:
            return False
    for l in range(1, h.nlevels):
        if isinstance(A, _LazyLevels):
            break
        mine = h.export_A(l)
        theirs = scipy.sparse.csr_matrix(A[l])
        if theirs.shape != mine.shape:
            return False
        d = abs(theirs - mine)
        if d.nnz and d.max() > 1e-12 * max(abs(mine).max(), 1e-300):
            return False
    return True


def _hierarchy_for(A, R, parameters):
    """The device hierarchy behind the lists handed to mgCycle: the one that produced them (infoDict['A']), the one
    in parameters['hierarchy'], or one built from A[0] — cached per (A, R) pair, keyed by identity AND by the
    content fingerprint of A[0], so an in-place edit of the matrix is not served from the cache."""
    h = getattr(A, '_hierarchy', None)
    if h is None:
        h = parameters.get('hierarchy', None)
    if h is not None:
        return h
    shape = getattr(R, 'problemShape', None) or parameters.get('problemShape')
    depth = min(int(parameters['coarsestLevel']), len(R))       # a shallower cycle than the list allows is legal
    A0 = A[0]
    fp = tools.fingerprint(A0) if (scipy.sparse.issparse(A0) or isinstance(A0, BandMatrix)) else None
    key = (id(A), id(R), depth)
    cached = _hierarchy_for._cache.get(key)
    if cached is not None and cached[0] is A and cached[1] is R:
        same = (fp is not None and cached[3] == fp) or (fp is None and np.array_equal(cached[4], np.asarray(A0)))
        if same:
            return cached[2]
    if depth < 1:
        raise ValueError("mgCycle needs at least one restriction (coarsestLevel >= 1)")
    h = Hierarchy(A0, shape, depth - 1, minSize=0)
    if h.nlevels != depth + 1:
        raise ValueError("R list does not match problemShape %r" % (shape,))
    if not _lists_match_hierarchy(h, A, R, shape):
        raise NotImplementedError(
            "mgCycle: the A / R lists differ from the Galerkin hierarchy of A[0] with the aggregation restrictions of "
            "problemShape %r; the device path only runs its own hierarchy (build the lists with "
            "operators.restrictionList / coeffecientList, or pass infoDict['A'], infoDict['R'])" % (shape,))
    _hierarchy_for._cache = {key: (A, R, h, fp, None if fp is not None else np.array(A0, copy=True))}
    return h


_hierarchy_for._cache = {}


def mgCycle(A, b, level, R, parameters, initial=None):
    """
    One recursive V-cycle entered at `level` (openmg/__init__.py:151-236).  `A` and `R`
    are the lists made by operators.coeffecientList / restrictionList (or infoDict['A'],
    infoDict['R']); the level operators used are the device-resident Galerkin hierarchy of
    A[0].  (Unlike the reference, whose in-place smoother overwrites `initial` with the
    pre-smoothed iterate, `initial` is left untouched.)  Returns (uOut, {'norm': norm}).
    """
    verbose = parameters['verbose']
    h = _hierarchy_for(A, R, parameters)
    if parameters['coarsestLevel'] != h.nlevels - 1:
        raise NotImplementedError("mgCycle: parameters['coarsestLevel'] = %r does not match the %d-level hierarchy "
                                  "behind these lists" % (parameters['coarsestLevel'], h.nlevels))
    bf = np.asarray(b, dtype=np.float64).ravel()
    N = bf.size
    pre, post = parameters['preIterations'], parameters['postIterations']
    if smooth is _ORIGINAL_SMOOTH:
        smoother, omega = _smoother_of(parameters)
        if verbose:
            for l in range(level, h.nlevels - 1):
                print(l * " " + "calling mgCycle at level %i" % l)
            print((h.nlevels - 1) * " " + "direct solving at level %i" % (h.nlevels - 1))
        uOut, norm = h.cycle(bf, initial, level, pre, post, smoother, omega)
        return uOut, {'norm': norm}
    # ---- user-supplied smoother: the reference's control flow, device call per step
    if initial is None:
        initial = np.zeros((N,))                                            # :191-192
    if level < parameters['coarsestLevel']:
        uApx = smooth(A[level], bf, initial, pre, verbose=verbose)          # :201
        if verbose:
            print(level * " " + "calling mgCycle at level %i" % level)
        coarseResidual = h.residual_restrict(level, bf, uApx)               # :209-210
        coarseCorrection = mgCycle(A, coarseResidual, level + 1, R, parameters)[0]   # :213
        corrected = h.prolong_correct(level, coarseCorrection, np.asarray(uApx).ravel())   # :214,:224
        if post > 0:
            uOut = smooth(A[level], bf, corrected, post, verbose=verbose)   # :216-222
        else:
            uOut = corrected
        norm = h.residual_norm(level, bf, np.asarray(uOut).ravel())         # :227
    else:
        norm = 0
        if verbose:
            print(level * " " + "direct solving at level %i" % level)
        uOut = h.coarse_solve(bf)                                           # :234
    return uOut, {'norm': norm}
