'''Helper functions of the reference's `openmg.tools` (openmg/tools.py), same
names and calling conventions.  Matrix-vector work runs on the device.'''
import functools
import operator

import numpy as np
import scipy.sparse as sparse

from . import hierarchy as _hier

_OP_CACHE = {}


def _operator_for(A, factor=False):
    """Device copy of A, cached by object identity (+ nnz/shape so a mutated
    pattern is re-uploaded)."""
    if isinstance(A, _hier.BandMatrix):
        key = (id(A), A.n, factor)
    elif sparse.issparse(A):
        key = (id(A), A.shape, A.nnz, factor)
    else:
        A = np.asarray(A)
        key = (id(A), A.shape, None, factor)
    ent = _OP_CACHE.get(key)
    if ent is not None and ent[0] is A:
        if sparse.issparse(A) or isinstance(A, _hier.BandMatrix):
            return ent[1]
        if np.array_equal(ent[2], A):      # dense arrays are mutable in place: verify
            return ent[1]
    op = _hier.Operator(A, factor=factor)
    if len(_OP_CACHE) > 16:
        _OP_CACHE.clear()
    _OP_CACHE[key] = (A, op, None if (sparse.issparse(A) or isinstance(A, _hier.BandMatrix)) else A.copy())
    return op


def getresidual(b, A, x, N):
    '''b - A x as an (N,1) column (openmg/tools.py:12-15), computed on the device.'''
    op = _operator_for(A)
    r = op.residual(0, np.asarray(b, dtype=np.float64).ravel(), np.asarray(x, dtype=np.float64).ravel())
    return r.reshape((N, 1))


def flexibleMmult(x, y):
    '''Dot two 2D arrays (openmg/tools.py:18-26).  matrix @ vector runs on the device;
    matrix @ matrix (only used by the reference to form R A R^T, which
    operators.coeffecientList does on the device instead) is left to numpy/scipy.'''
    yv = None
    if not sparse.issparse(y) and not isinstance(y, _hier.BandMatrix):
        ya = np.asarray(y)
        if ya.ndim == 1 or (ya.ndim == 2 and ya.shape[1] == 1):
            yv = ya
    is_mat = sparse.issparse(x) or isinstance(x, _hier.BandMatrix) or (np.asarray(x).ndim == 2)
    if yv is not None and is_mat:
        xs = x.shape
        if xs[0] == xs[1] and xs[1] == yv.shape[0]:
            out = _operator_for(x).matvec(np.asarray(yv, dtype=np.float64).ravel())
            return out.reshape(yv.shape)
    if isinstance(x, _hier.BandMatrix):
        x = x.tocsr()
    if isinstance(y, _hier.BandMatrix):
        y = y.tocsr()
    if sparse.issparse(x) or sparse.issparse(y):
        return x * y                    # scipy matrix semantics: a matrix product
    return np.dot(x, y)


def dictUpdateNoClobber(updateDict, targetDict):
    """Copy the entries of `updateDict` into `targetDict` unless the key is already there; returns `targetDict`
    (the reference's helper of the same name, openmg/tools.py:29-40).
    >>> opts = {'cycles': 3}
    >>> dictUpdateNoClobber({'cycles': 10, 'minSize': 8}, opts) is opts
    True
    >>> opts == {'cycles': 3, 'minSize': 8}
    True
    """
    for name in updateDict:
        targetDict.setdefault(name, updateDict[name])
    return targetDict


def dictAddNoClobber(dictionary, key, value):
    """dictionary[key] = value only if `key` is missing; returns the dictionary (openmg/tools.py:43-53).
    >>> dictAddNoClobber({'omega': 0.8}, 'omega', 1.0)
    {'omega': 0.8}
    """
    dictionary.setdefault(key, value)
    return dictionary


def product(iterableThing):
    """Product of the entries, 1 for an empty iterable (openmg/tools.py:56-60)."""
    return functools.reduce(operator.mul, iterableThing, 1)
