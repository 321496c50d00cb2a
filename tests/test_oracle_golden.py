"""CPU: the oracle restatement against the golden vectors produced by the
reference itself (tools/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import assert_same_csr, csr_from, key, seeded_problem
import oracle.openmg_oracle as orc


def test_restriction_patterns_bit_exact(gold_operators):
    z, meta = gold_operators
    for shape in meta["restriction_shapes"]:
        shape = tuple(shape)
        R = orc.restriction(shape)
        assert R.dtype == np.float64
        assert R.has_sorted_indices
        assert_same_csr(R, z, "R/" + key(shape))
        if orc.product(shape) <= 600:
            assert_same_csr(orc.restriction_loop(shape), z, "R/" + key(shape))
        np.testing.assert_array_equal(orc.restriction(shape, dense=True), R.toarray())


def test_restriction_errors(gold_operators):
    _, meta = gold_operators
    for shape, err in meta["restriction_errors"]:
        shape = tuple(shape)
        if err == "none":
            orc.restriction(shape)
            continue
        exc = {"ValueError": ValueError, "IndexError": IndexError}[err]
        with pytest.raises(exc):
            orc.restriction(shape)


def test_poisson_generators_bit_exact(gold_operators):
    z, meta = gold_operators
    for shape, sparse_flag in meta["poisson"]:
        shape = tuple(shape)
        pre = "P/%s/%d" % (key(shape), int(sparse_flag))
        assert_same_csr(sp.csr_matrix(orc.poisson(shape, sparse=sparse_flag)), z, pre)
        assert_same_csr(orc.poisson_csr(shape, sparse_1d=sparse_flag), z, pre)
    with pytest.raises(ValueError):
        orc.poisson((1, 2, 3, 4))
    with pytest.raises(NotImplementedError):
        orc.poisson((4, 4), sparse=True)
    # int shape accepted (openmg/operators.py:264-265)
    assert_same_csr(orc.poisson(8, sparse=True), z, "P/8/1")


def test_restriction_list_depth_rule(gold_operators):
    _, meta = gold_operators
    for shape, cl, ms, shapes in meta["rlist"]:
        Rl = orc.restrictionList(tuple(shape), cl, ms)
        assert [list(r.shape) for r in Rl] == shapes, (shape, cl, ms)


def test_galerkin_levels_bit_exact(gold_galerkin):
    z, meta = gold_galerkin
    for name, shape, sparse_flag, pshape, gl, nlev, sizes in meta:
        A_in = orc.poisson_csr(tuple(shape), sparse_1d=sparse_flag)
        R = orc.restrictionList(tuple(pshape), gl - 1, 8)
        A = orc.coeffecientList(A_in, R)
        assert len(A) == nlev and [a.shape[0] for a in A] == sizes
        for l, Al in enumerate(A):
            pre = "%s/A%d" % (name, l)
            if pre + "/indptr" in z.files:
                assert_same_csr(Al, z, pre)


def test_gauss_seidel_matches_reference(gold_smoothers):
    z, meta = gold_smoothers
    for shape, sparse_flag, tag, it in meta:
        k = "gs/%s/%d/%s/%d" % (key(tuple(shape)), int(sparse_flag), tag, it)
        A = orc.poisson_csr(tuple(shape), sparse_1d=sparse_flag)
        As = A if tag == "csr" else A.toarray()
        x = z[k + "/x0"].copy()
        out = orc.gaussSeidel(As, z[k + "/b"], x, iterations=it)
        assert out is x                                    # in place (openmg/solvers.py:68,75)
        np.testing.assert_allclose(x, z[k + "/x"], rtol=1e-13, atol=1e-15)
        if tag == "csr":
            xc = z[k + "/x0"].copy()
            orc.gaussSeidel_c(A, z[k + "/b"], xc, it)
            np.testing.assert_allclose(xc, z[k + "/x"], rtol=1e-13, atol=1e-15)
    A = orc.poisson_csr((12, 12))
    x = orc.gaussSeidel(A, z["gs_thresh/b"], np.zeros(144), threshold=1e-4)
    np.testing.assert_allclose(x, z["gs_thresh/x"], rtol=1e-12, atol=1e-14)
    assert np.linalg.norm(z["gs_thresh/b"] - A.dot(x)) < 1e-4


def test_coarse_solve_matches_reference(gold_smoothers):
    z, _ = gold_smoothers
    for shape, sparse_flag in (((64,), True), ((8, 8), False), ((4, 4, 4), False)):
        A = orc.poisson_csr(shape, sparse_1d=sparse_flag)
        pre = "coarse/%s/%d" % (key(shape), int(sparse_flag))
        x = orc.coarseSolve(A, z[pre + "/b"].reshape(-1, 1))
        assert x.shape == (A.shape[0],)
        np.testing.assert_allclose(x, z[pre + "/x"], rtol=1e-12, atol=0)


def test_vcycles_match_reference(gold_cycles):
    """4 V-cycles: norms per cycle and final iterate, every smoother, against
    the reference's own mgCycle (lexicographic GS untouched; Jacobi / 2-colour
    through its `openmg.smooth` plug-in point).  Tolerance 1e-12 relative."""
    z, meta = gold_cycles
    for name, shape, sparse_flag, pshape, gl, smoother, pre, post, ncyc in meta["cycles"]:
        A_in = orc.poisson_csr(tuple(shape), sparse_1d=sparse_flag)
        _, b = seeded_problem(A_in)
        params = {'problemShape': tuple(pshape), 'gridLevels': gl, 'preIterations': pre,
                  'postIterations': post, 'verbose': False, 'minSize': 8}
        R = orc.restrictionList(tuple(pshape), gl - 1, 8)
        params['coarsestLevel'] = len(R)
        A = orc.coeffecientList(A_in, R)
        smooth = orc.make_smoother(smoother, tuple(pshape), 0.8)
        x, norms = None, []
        for _ in range(ncyc):
            x, info = orc.mgCycle(A, b, 0, R, params, initial=x, smooth=smooth)
            norms.append(info['norm'])
        tag = "%s/%s/%d%d" % (name, smoother, pre, post)
        # norms: 1e-10 relative, plus rounding noise relative to the starting residual
        np.testing.assert_allclose(norms, z[tag + "/norms"], rtol=1e-10,
                                   atol=1e-13 * max(np.linalg.norm(b), 1.0), err_msg=tag)
        scale = np.abs(z[tag + "/x"]).max()
        np.testing.assert_allclose(x, z[tag + "/x"], rtol=0, atol=1e-12 * scale, err_msg=tag)


def test_mgsolve_stop_rules_match_reference(gold_cycles):
    z, meta = gold_cycles
    for N, gl, cycles, thr, cyc_done, norm, coarsest in meta["mgsolve"]:
        if N == 100:
            A = orc.poisson(N, sparse=True)
            u_true = np.array([np.sin(x / 10.0) for x in np.linspace(0, 20, N)])
        else:
            A = orc.poisson((N,))
            u_true = np.sin(np.array(range(int(N))) * 3.0 / N).T
        b = np.asarray(orc.flexibleMmult(A, u_true)).ravel()
        params = {'problemShape': (N,), 'gridLevels': gl, 'cycles': cycles, 'threshold': thr, 'giveInfo': True}
        x, info = orc.mgSolve(A, b, params)
        assert info['cycle'] == cyc_done
        assert params['coarsestLevel'] == coarsest          # mutated (openmg/__init__.py:106)
        assert np.isclose(info['norm'], norm, rtol=1e-9)
        np.testing.assert_allclose(x, z["mgsolve/%d_%d_%d_%g/x" % (N, gl, cycles, thr)], rtol=1e-10, atol=1e-13)
    with pytest.raises(ValueError):
        A = orc.poisson((64,))
        orc.mgSolve(A, np.ones(64), {'problemShape': (64,), 'gridLevels': 2, 'cycles': 0, 'threshold': 0})


def test_band_matvec_equals_csr():
    for shape, s1 in (((37,), True), ((12, 12), False), ((6, 6, 6), False), ((4, 6, 8), False)):
        A = orc.poisson_csr(shape, sparse_1d=s1)
        d, bands = orc.poisson_bands(shape, sparse_1d=s1)
        x = np.random.RandomState(5).random_sample(A.shape[0])
        np.testing.assert_allclose(orc.band_matvec(d, bands, x), A.dot(x), rtol=1e-14, atol=1e-14)


def test_colouring_rule():
    c = orc.colouring((8, 8), 0, 64)
    np.testing.assert_array_equal(c, np.arange(64) & 1)
    c = orc.colouring((8, 8, 8), 0, 512).reshape(8, 8, 8)
    i, j, k = np.indices((8, 8, 8))
    np.testing.assert_array_equal(c, (i + j + k) & 1)
    c1 = orc.colouring((8, 8), 1, 16).reshape(4, 4)
    i, j = np.indices((4, 4))
    np.testing.assert_array_equal(c1, (i + j) & 1)
