"""GPU: the drop-in Python surface behaves like openmg's (the reference's own test
list, openmg/tests.py, re-run against openmg_b200)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import seeded_problem
import oracle.openmg_oracle as orc

pytestmark = pytest.mark.gpu

import openmg_b200 as openmg            # noqa: E402
from openmg_b200 import operators, tools, solvers   # noqa: E402


def base_parameters():
    # openmg/tests.py:46-56
    return {'problemShape': (512,), 'gridLevels': 3, 'iterations': 1, 'verbose': False, 'threshold': 4,
            'giveInfo': True}


def test_a_1d_matrix_with_3d_hierarchy():
    # openmg/tests.py:58-81
    problemscale = 12
    size = problemscale ** 3
    gridLevels = 4
    u_actual = np.sin(np.array(range(int(size))) * 3.0 / size).T
    A = operators.poisson((size,))
    b = tools.flexibleMmult(A, u_actual)
    uSmoothed = openmg.smooth(A, b, np.zeros((size,)), iterations=1)
    assert uSmoothed.shape == (size,)
    parameters = {'coarsestLevel': gridLevels - 1, 'problemShape': (problemscale,) * 3,
                  'gridLevels': gridLevels, 'threshold': 8e-3}
    u_mmg = openmg.mgSolve(A, b, parameters)
    assert u_mmg.shape == (size,)
    assert parameters['threshold'] > np.linalg.norm(tools.flexibleMmult(A, u_mmg) - b)


def test_gs_thresh_and_nothresh():
    # openmg/tests.py:342-364
    A = operators.poisson((12, 12))
    b = np.random.RandomState(0).random_sample(144)
    x = solvers.smoothToThreshold(A, b, np.zeros(144), 1e-4)
    assert np.linalg.norm(tools.getresidual(b, A, x, 144)) < 1e-4
    xg = np.zeros(144)
    out = solvers.gaussSeidel(A, b, xg, threshold=1e-4)
    assert out is xg and np.linalg.norm(b - A.dot(xg)) < 1e-4
    A1 = operators.poisson((12,))
    b1 = np.random.RandomState(1).random_sample(12)
    x1 = solvers.gaussSeidel(A1, b1, np.zeros(12))               # neither stop: one sweep
    np.testing.assert_allclose(x1, orc.gaussSeidel(A1, b1, np.zeros(12)), rtol=1e-13)


def test_smooth_in_place_and_shapes():
    A = operators.poisson(64, sparse=True)
    b = np.random.RandomState(2).random_sample(64)
    x = np.zeros((64, 1))                                       # openmg_usage_demo.py:94 passes (N,1)
    out = openmg.smooth(A, b, x, 2)
    assert out is x and x.shape == (64, 1) and np.abs(x).max() > 0
    xs = np.zeros(64)
    np.testing.assert_allclose(solvers.gaussSeidel(A, b, xs, iterations=3),
                               orc.gaussSeidel(A, b, np.zeros(64), iterations=3), rtol=1e-13)
    np.testing.assert_allclose(solvers.jacobi(A, b, np.zeros(64), 2, 0.7),
                               orc.jacobi(A, b, np.zeros(64), 2, 0.7), rtol=1e-13)
    r = tools.getresidual(b, A, xs, 64)
    assert r.shape == (64, 1)
    np.testing.assert_allclose(r.ravel(), b - A.dot(xs), rtol=1e-13, atol=1e-15)


def test_coarse_solve():
    for shape, s in (((64,), True), ((8, 8), False), ((4, 4, 4), False)):
        A = orc.poisson_csr(shape, sparse_1d=s)
        b = np.random.RandomState(3).random_sample(A.shape[0])
        x = openmg.coarseSolve(A, b.reshape(-1, 1))
        assert x.shape == (A.shape[0],)
        np.testing.assert_allclose(x, orc.coarseSolve(A, b), rtol=1e-11)
        np.testing.assert_allclose(openmg.coarseSolve(A.toarray(), b), x, rtol=1e-12)
    with pytest.raises(np.linalg.LinAlgError):
        openmg.coarseSolve(sp.csr_matrix(np.array([[1.0, 2.0], [2.0, 4.0]])), np.ones(2))


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_noise_mg(dim):
    # openmg/tests.py:366-423
    shape = {1: (512,), 2: (32, 32), 3: (8, 8, 8)}[dim]
    N = int(np.prod(shape))
    u_actual = np.random.RandomState(dim).random_sample(N).reshape((N, 1))
    A_in = operators.poisson(shape)
    b = tools.flexibleMmult(A_in, u_actual)
    assert b.shape == (N, 1)
    parameters = base_parameters()
    parameters['problemShape'] = shape
    u_mg, info = openmg.mgSolve(A_in, b, parameters)
    assert u_mg.shape == (N,) and u_mg.dtype == np.float64
    assert set(('cycle', 'norm', 'R', 'A')) <= set(info)
    assert info['norm'] < 4 or info['cycle'] >= 1
    if dim == 3:
        assert len(info['R']) == 1 and len(info['A']) == 2     # minSize=8 cuts 8^3 to one transition
        assert info['A'][1].shape == (64, 64) and info['R'][0].shape == (64, 512)
    assert abs(np.linalg.norm(b.ravel() - A_in.dot(u_mg)) - info['norm']) < 1e-10


def test_dense_flag_and_mpl_demo_configuration():
    # openmg/tests.py:454-500 (the solve, not the plot): 16x16, gridLevels 2, cycles 300, dense=True
    NX = 16
    A_in = operators.poisson((NX, NX))
    u = np.random.RandomState(5).random_sample(NX * NX)
    b = tools.flexibleMmult(A_in, u)
    u_mg = openmg.mgSolve(A_in, b, {'problemShape': (NX, NX), 'gridLevels': 2, 'iterations': 1,
                                    'verbose': False, 'cycles': 300, 'dense': True})
    assert u_mg.shape == (NX * NX,)
    # the reference's 2-D hierarchy converges slowly; compare with the oracle, same smoother
    u_or = orc.mgSolve(sp.csr_matrix(A_in), b, {'problemShape': (NX, NX), 'gridLevels': 2, 'cycles': 300,
                                                'smoother': 'rbgs'})
    np.testing.assert_allclose(u_mg, u_or, rtol=0, atol=1e-10 * np.abs(u_or).max())


def test_thresh_stop_and_cycle_stop(gold_cycles):
    # openmg/tests.py:502-531, and the golden mgSolve runs for the stop bookkeeping
    size = 36
    u_actual = np.sin(np.array(range(int(size))) * 3.0 / size).T
    A = operators.poisson((size,))
    b = tools.flexibleMmult(A, u_actual)
    parameters = {'problemShape': (size,), 'gridLevels': 2, 'threshold': 8e-3, 'giveInfo': True}
    u_mmg, info = openmg.mgSolve(A, b, parameters)
    assert parameters['threshold'] > np.linalg.norm(tools.flexibleMmult(A, u_mmg) - b)
    assert parameters['coarsestLevel'] == 2 and parameters['minSize'] == 8      # dict mutated (:96,:106)
    assert openmg.defaults['coarsestLevel'] == 1                                # module global mutated (:95)
    parameters = {'problemShape': (size,), 'gridLevels': 2, 'cycles': 3, 'threshold': 1e-10, 'giveInfo': True}
    u_mmg, info = openmg.mgSolve(A, b, parameters)
    assert info['cycle'] == parameters['cycles']
    # lexicographic GS on the device reproduces the reference's own mgSolve outputs
    z, meta = gold_cycles
    for N, gl, cycles, thr, cyc_done, norm, coarsest in meta["mgsolve"]:
        if N == 100:
            A = operators.poisson(N, sparse=True)
            u_true = np.array([np.sin(x / 10.0) for x in np.linspace(0, 20, N)])
        else:
            A = operators.poisson((N,))
            u_true = np.sin(np.array(range(int(N))) * 3.0 / N).T
        b = tools.flexibleMmult(A, u_true)
        params = {'problemShape': (N,), 'gridLevels': gl, 'cycles': cycles, 'threshold': thr, 'giveInfo': True,
                  'smoother': 'gs'}
        x, info = openmg.mgSolve(A, b, params)
        assert info['cycle'] == cyc_done and params['coarsestLevel'] == coarsest
        assert np.isclose(info['norm'], norm, rtol=1e-8)
        np.testing.assert_allclose(x, z["mgsolve/%d_%d_%d_%g/x" % (N, gl, cycles, thr)], rtol=1e-9, atol=1e-12)


def test_threshold_stop_runs_on_the_device():
    """A thresholded solve tests its stop rule on the device (every cycle sits behind an IF node of its CUDA graph):
    same cycle count, final norm, history and iterate as the oracle's mgSolve loop, with fewer host
    synchronisations than cycles; the direct-launch path (one read-back per cycle) agrees bit for bit."""
    from openmg_b200 import _lib
    from openmg_b200.hierarchy import Hierarchy
    for shape, gl, smoother, thr in (((32, 32, 32), 2, "jacobi", 1e-7), ((64, 64), 2, "rbgs", 2e-2),
                                     ((4096,), 5, "jacobi", 1e-9)):
        s1 = len(shape) == 1
        A_in = orc.poisson_csr(shape, sparse_1d=s1)
        _, b = seeded_problem(A_in)
        params = {'problemShape': shape, 'gridLevels': gl, 'cycles': 60, 'threshold': thr, 'preIterations': 1,
                  'postIterations': 1, 'smoother': smoother, 'giveInfo': True}
        xo, io = orc.mgSolve(A_in, b, dict(params))
        assert 4 < io['cycle'] < 60, io['cycle']            # the threshold, not the cap, ends the loop
        h = Hierarchy(A_in, shape, gl - 1, 8)
        x, cyc, norm, hist = h.solve(b, None, 1, 1, smoother, 0.8, 60, thr, want_history=True)
        stats = h.solve_stats()
        assert cyc == io['cycle'] and len(hist) == cyc
        assert np.isclose(norm, io['norm'], rtol=1e-7) and norm < thr <= hist[-2]
        np.testing.assert_allclose(x, xo, rtol=1e-9, atol=1e-12)
        assert stats['host_syncs'] < cyc, stats
        # the cycle cap also lives on the device: same history, stops at 3
        x3, cyc3, norm3, hist3 = h.solve(b, None, 1, 1, smoother, 0.8, 3, thr, want_history=True)
        assert cyc3 == 3 and np.array_equal(hist3, hist[:3])
        h.close()
        h2 = Hierarchy(A_in, shape, gl - 1, 8, flags=_lib.FLAG_NO_GRAPH)
        x2, cyc2, norm2, hist2 = h2.solve(b, None, 1, 1, smoother, 0.8, 60, thr, want_history=True)
        assert cyc2 == cyc and h2.solve_stats()['host_syncs'] == cyc
        np.testing.assert_array_equal(x2, x)
        np.testing.assert_array_equal(hist2, hist)
        h2.close()


def test_error_behaviour():
    with pytest.raises(ValueError):
        operators.poisson((1, 2, 3, 4))                           # openmg/tests.py:533-536
    for alpha in range(1, 4):
        for dense in (True, False):
            operators.restriction((4,) * alpha, dense=dense)      # :538-542
    with pytest.raises(ValueError):
        operators.restriction((4, 4, 4, 4))                       # :544-548
    A = operators.poisson((512,))
    b = np.ones(512)
    p = base_parameters()
    p['cycles'] = 0
    p['threshold'] = 0
    with pytest.raises(ValueError):                               # :550-556
        openmg.mgSolve(A, b, p)
    with pytest.raises(KeyError):
        openmg.mgSolve(A, b, {'gridLevels': 2})
    with pytest.raises(KeyError):
        openmg.mgSolve(A, b, {'problemShape': (512,)})
    with pytest.raises(ValueError):                               # coarse set would have 1 point
        openmg.mgSolve(operators.poisson((4,)), np.ones(4), {'problemShape': (4,), 'gridLevels': 3, 'cycles': 1})


def test_min_size():
    # openmg/tests.py:558-570
    parameters = base_parameters()
    shape = parameters["problemShape"] = (1024,)
    parameters["gridLevels"] = 24
    parameters["minSize"] = 23
    u_actual = np.random.RandomState(9).random_sample(shape).ravel()
    A_in = operators.poisson(shape)
    b = tools.flexibleMmult(A_in, u_actual)
    soln, info = openmg.mgSolve(A_in, b, parameters)
    assert min(info['R'][-1].shape) > parameters["minSize"]
    assert [r.shape for r in info['R']] == [r.shape for r in orc.restrictionList(shape, 23, 23)]


def test_smoother_plug_in_point():
    """`openmg.smooth = f` swaps the smoother (openmg/__init__.py:201,218): drive the device
    cycle with the oracle's Jacobi and compare with the built-in device Jacobi."""
    shape = (16, 16)
    A = operators.poisson(shape, sparse=True)
    _, b = seeded_problem(A)
    params = {'problemShape': shape, 'gridLevels': 2, 'cycles': 3, 'threshold': 0, 'preIterations': 1,
              'postIterations': 1, 'smoother': 'jacobi', 'giveInfo': True}
    x_dev, info_dev = openmg.mgSolve(A, b, dict(params))
    saved = openmg.smooth
    try:
        openmg.smooth = lambda A_, b_, x_, it, verbose=False: orc.jacobi(A_, b_, x_, it, 0.8)
        x_plug, info_plug = openmg.mgSolve(A, b, dict(params))
    finally:
        openmg.smooth = saved
    np.testing.assert_allclose(x_plug, x_dev, rtol=1e-12, atol=1e-14)
    assert info_plug['cycle'] == 3 and np.isclose(info_plug['norm'], info_dev['norm'], rtol=1e-9)


def test_mgcycle_api_with_lists():
    shape = (32,)
    A_in = operators.poisson(shape, sparse=True)
    _, b = seeded_problem(A_in)
    R = operators.restrictionList(shape, 1, 8)
    A = operators.coeffecientList(A_in, R)
    params = {'coarsestLevel': len(R), 'preIterations': 1, 'postIterations': 0, 'verbose': False,
              'problemShape': shape, 'smoother': 'jacobi'}
    u1, info = openmg.mgCycle(A, b, 0, R, params)
    Ro = orc.restrictionList(shape, 1, 8)
    Ao = orc.coeffecientList(A_in, Ro)
    u2, info2 = orc.mgCycle(Ao, b, 0, Ro, params, smooth=orc.make_smoother('jacobi', shape, 0.8))
    np.testing.assert_allclose(u1, u2, rtol=1e-12)
    assert np.isclose(info['norm'], info2['norm'], rtol=1e-10)
    # entering at level 1 with a coarse right-hand side
    bc = np.random.RandomState(2).random_sample(A[1].shape[0])
    v1, i1 = openmg.mgCycle(A, bc, 1, R, params)
    v2, i2 = orc.mgCycle(Ao, bc, 1, Ro, params, smooth=orc.make_smoother('jacobi', shape, 0.8))
    np.testing.assert_allclose(v1, v2, rtol=1e-12)
    assert np.isclose(i1['norm'], i2['norm'], rtol=1e-10, atol=1e-14)


def test_mgcycle_lists_semantics():
    """The reference's mgCycle runs on whatever A / R lists it is handed (openmg/__init__.py:199-214).  The device
    path accepts lists that equal its own hierarchy (also a shallower coarsestLevel than the lists allow, and the
    lists mgSolve returned), refuses anything else loudly, and never serves a stale device copy after the caller
    edited the matrix in place."""
    shape = (16, 16)
    A_in = sp.csr_matrix(operators.poisson(shape))
    _, b = seeded_problem(A_in)
    R = operators.restrictionList(shape, 1, 2)
    A = operators.coeffecientList(A_in, R)
    assert len(R) == 2
    sm = orc.make_smoother('rbgs', shape, 0.8)
    base = {'preIterations': 1, 'postIterations': 1, 'verbose': False, 'problemShape': shape, 'smoother': 'rbgs'}
    Ro = orc.restrictionList(shape, 1, 2)
    Ao = orc.coeffecientList(A_in, Ro)
    for depth in (2, 1):              # depth 1: direct solve on level 1 although the lists go deeper
        params = dict(base, coarsestLevel=depth)
        u1, info = openmg.mgCycle(A, b, 0, R, params)
        u2, info2 = orc.mgCycle(Ao, b, 0, Ro, params, smooth=sm)
        np.testing.assert_allclose(u1, u2, rtol=1e-10)
        assert np.isclose(info['norm'], info2['norm'], rtol=1e-9)
    # lists that are not the Galerkin hierarchy of A[0]: refused, not silently replaced
    A_bad = list(A)
    A_bad[1] = A[1] * 1.5
    with pytest.raises(NotImplementedError):
        openmg.mgCycle(A_bad, b, 0, R, dict(base, coarsestLevel=2))
    R_bad = list(R)
    R_bad[0] = R[0] * 2.0
    with pytest.raises(NotImplementedError):
        openmg.mgCycle(A, b, 0, R_bad, dict(base, coarsestLevel=2))
    # the lists mgSolve hands back drive mgCycle on the same device hierarchy
    x, info = openmg.mgSolve(A_in, b, dict(base, gridLevels=2, minSize=2, cycles=2, threshold=0, giveInfo=True))
    params = dict(base, coarsestLevel=len(info['R']))
    y, _ = openmg.mgCycle(info['A'], b, 0, info['R'], params, initial=x)
    z, _ = orc.mgCycle(Ao, b, 0, Ro, params, initial=np.array(x), smooth=sm)
    np.testing.assert_allclose(y, z, rtol=1e-10)
    # in-place edit of the matrix between calls: the cached device copies must not be reused
    A_in2 = A_in.copy()
    xr = np.random.RandomState(5).random_sample(A_in2.shape[0])
    r1 = tools.getresidual(b, A_in2, xr, xr.size).ravel()
    A_in2.data *= 2.0
    r2 = tools.getresidual(b, A_in2, xr, xr.size).ravel()
    np.testing.assert_allclose(r2, b - A_in2.dot(xr), rtol=1e-13)
    assert np.abs(r2 - r1).max() > 1e-3
    lists = [A_in2] + list(operators.coeffecientList(A_in2, R)[1:])
    u1, _ = openmg.mgCycle(lists, b, 0, R, dict(base, coarsestLevel=2))
    A_in2.data *= 0.5                      # lists[0] edited in place: the levels below no longer match
    with pytest.raises(NotImplementedError):
        openmg.mgCycle(lists, b, 0, R, dict(base, coarsestLevel=2))


def test_simple_demo_trace():
    # openmg_usage_demo.py:27-41 with the reference's smoother: residual trace of today's code
    N = 100
    u_true = np.array([np.sin(x / 10.0) for x in np.linspace(0, 20, N)])
    A = operators.poisson(N, sparse=True)
    b = tools.flexibleMmult(A, u_true)
    params = {'problemShape': (N,), 'gridLevels': 3, 'cycles': 10, 'iterations': 2, 'verbose': False,
              'dense': True, 'threshold': 1e-2, 'giveInfo': True, 'smoother': 'gs'}
    u_mg, info = openmg.mgSolve(A, b, params)
    assert info['cycle'] == 4
    np.testing.assert_allclose(info['norms'], [0.805593, 0.108255, 0.018646, 0.003409], rtol=3e-4)   # 6 printed digits
