"""Direct solver and relaxation methods of the reference's `openmg.solvers`
(openmg/solvers.py), same names and in-place conventions, running on the device.

The reference's only smoother is a lexicographic Gauss-Seidel (a Python loop over rows).
Here `smooth` runs the package default (two-colour Gauss-Seidel), `gaussSeidel` keeps the
reference's lexicographic semantics (sequential device kernel — exact, for small systems),
and `jacobi` / `rbgs` expose the two parallel smoothers directly."""
import numpy as np
import scipy.sparse as sparse

from . import tools

DEFAULT_SMOOTHER = 'rbgs'
DEFAULT_OMEGA = 0.8


def coarseSolve(A, b):
    """Direct solve of A x = b (openmg/solvers.py:16-26): dense inverse formed once on the
    device (Gauss-Jordan with partial pivoting), applied as a GEMV.  Returns a flat array."""
    op = tools._operator_for(A, factor=True)
    return np.ravel(op.coarse_solve(np.asarray(b, dtype=np.float64).ravel()))


def _writeback(x, result):
    """The reference mutates x in place and returns the same object (openmg/solvers.py:68,75)."""
    if isinstance(x, np.ndarray) and x.dtype == np.float64:
        x[...] = result.reshape(x.shape)
        return x
    return result.reshape(np.shape(x))


def _run(A, b, x, iterations, threshold, smoother, omega):
    if iterations is None and threshold is None:       # openmg/solvers.py:39-40
        iterations = 1
    op = tools._operator_for(A)
    bf = np.asarray(b, dtype=np.float64).ravel()
    xf = np.asarray(x, dtype=np.float64).ravel()
    if threshold is None:
        out = op.smooth(0, bf, xf, int(iterations), smoother, omega)
    else:
        cap = int(iterations) if iterations is not None else 10 ** 7
        out, _, _ = op.smooth_to_threshold(0, bf, xf, threshold, smoother, omega, max_sweeps=cap)
    return _writeback(x, out)


def smooth(A, b, x, iterations, verbose=False, smoother=None, omega=None):
    """`iterations` sweeps of the default smoother, in place (openmg/solvers.py:28-29)."""
    return _run(A, b, x, iterations, None, smoother or DEFAULT_SMOOTHER, omega or DEFAULT_OMEGA)


def smoothToThreshold(A, b, x, threshold, verbose=False, smoother=None, omega=None):
    """Sweep until ||b - A x||_2 < threshold (openmg/solvers.py:31-32)."""
    return _run(A, b, x, None, threshold, smoother or DEFAULT_SMOOTHER, omega or DEFAULT_OMEGA)


def gaussSeidel(A, b, x, iterations=None, threshold=None, verbose=False):
    """Lexicographic forward Gauss-Seidel with the reference's stop rule
    (openmg/solvers.py:34-75), on the device."""
    return _run(A, b, x, iterations, threshold, 'gs', 1.0)


def jacobi(A, b, x, iterations=1, omega=DEFAULT_OMEGA, threshold=None):
    """Weighted Jacobi x <- x + omega (b - A x)/diag(A), in place."""
    return _run(A, b, x, iterations, threshold, 'jacobi', omega)


def rbgs(A, b, x, iterations=1, threshold=None):
    """Two-colour Gauss-Seidel (flat-index parity colouring), in place."""
    return _run(A, b, x, iterations, threshold, 'rbgs', 1.0)
