"""GPU: the CUDA path, called through the C-ABI (ctypes), against the oracle and the
reference goldens.  Integer/index work bit-exact; fp64 within the tolerances the
north star states (Jacobi 1e-12 relative, two-colour GS 1e-10 relative)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import assert_same_csr, key, seeded_problem
import oracle.openmg_oracle as orc

pytestmark = pytest.mark.gpu

import openmg_b200 as omg                      # noqa: E402
from openmg_b200 import _lib                   # noqa: E402
from openmg_b200.hierarchy import Hierarchy    # noqa: E402

JAC_RTOL = 1e-12
RB_RTOL = 1e-10


def close(a, b, rtol, what=""):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, "%s: max rel err %.3e > %.1e" % (what, err, rtol)


# ------------------------------------------------------------------ index work: bit exact

def test_restriction_bit_exact_vs_reference(gold_operators):
    z, meta = gold_operators
    for shape in meta["restriction_shapes"]:
        shape = tuple(shape)
        R = omg.operators.restriction(shape)
        assert R.dtype == np.float64 and R.indices.dtype == np.int32
        assert_same_csr(R, z, "R/" + key(shape))
    np.testing.assert_array_equal(omg.operators.restriction((4, 4), dense=True),
                                  orc.restriction((4, 4), dense=True))
    P = omg.operators.interpolation((8, 8))
    assert (P != orc.restriction((8, 8)).T).nnz == 0


def test_restriction_errors_match_reference(gold_operators):
    _, meta = gold_operators
    for shape, err in meta["restriction_errors"]:
        exc = {"ValueError": ValueError, "IndexError": IndexError}[err]
        with pytest.raises(exc):
            omg.operators.restriction(tuple(shape))


def test_restriction_list_depth_rule(gold_operators):
    _, meta = gold_operators
    for shape, cl, ms, shapes in meta["rlist"]:
        Rl = omg.operators.restrictionList(tuple(shape), cl, ms)
        assert [list(r.shape) for r in Rl] == shapes, (shape, cl, ms)


@pytest.mark.parametrize("via", ["csr", "band", "force_csr"])
def test_galerkin_bit_exact_vs_reference(gold_galerkin, via):
    z, meta = gold_galerkin
    for name, shape, sparse_flag, pshape, gl, nlev, sizes in meta:
        shape, pshape = tuple(shape), tuple(pshape)
        if via == "band":
            A_in = omg.operators.poisson_band(shape, sparse_1d=sparse_flag)
        else:
            A_in = orc.poisson_csr(shape, sparse_1d=sparse_flag)
        flags = _lib.FLAG_FORCE_CSR if via == "force_csr" else 0
        h = Hierarchy(A_in, pshape, gl - 1, 8, flags=flags)
        assert h.nlevels == nlev
        for l in range(nlev):
            assert h.level_info(l)["n"] == sizes[l]
            pre = "%s/A%d" % (name, l)
            if pre + "/indptr" in z.files:
                assert_same_csr(h.export_A(l), z, pre)
        for l in range(nlev - 1):
            Rref = orc.restriction(orc.level_shape(pshape, l))
            assert (h.export_R(l) != Rref).nnz == 0
        h.close()


def test_coeffecient_list_api(gold_galerkin):
    z, _ = gold_galerkin
    A_in = omg.operators.poisson((16, 16))
    R = omg.operators.restrictionList((16, 16), 1, 8)
    A = omg.operators.coeffecientList(A_in, R)
    assert len(A) == len(R) + 1
    Aref = orc.coeffecientList(sp.csr_matrix(A_in), orc.restrictionList((16, 16), 1, 8))
    for a, b in zip(A, Aref):
        assert (orc.canonical_csr(a) != orc.canonical_csr(b)).nnz == 0


def test_non_regular_shapes_match_oracle():
    """Shapes where the reference's NX-offset quirk makes aggregates overlap or the
    grid is odd: explicit-R path."""
    for pshape, gl in (((4, 6), 1), ((9,), 1), ((25,), 1), ((4, 6, 8), 1), ((6, 6), 2), ((10, 10), 2)):
        N = int(np.prod(pshape))
        A_in = orc.poisson_csr(pshape)
        try:
            R = orc.restrictionList(pshape, gl - 1, 2)
            Aref = orc.coeffecientList(A_in, R)
        except ValueError:
            with pytest.raises(ValueError):
                Hierarchy(A_in, pshape, gl - 1, 2)
            continue
        h = Hierarchy(A_in, pshape, gl - 1, 2)
        assert h.nlevels == len(Aref)
        for l, a in enumerate(Aref):
            got, want = h.export_A(l), orc.canonical_csr(a)
            assert (got != want).nnz == 0, (pshape, l)
            np.testing.assert_array_equal(got.indices, want.indices)
        u, b = seeded_problem(A_in)
        x = np.random.RandomState(3).random_sample(N)
        rc = h.residual_restrict(0, b, x)
        close(rc, R[0].dot(b - A_in.dot(x)), 1e-13, "residual_restrict %r" % (pshape,))
        e = np.random.RandomState(4).random_sample(R[0].shape[0])
        close(h.prolong_correct(0, e, x), x + R[0].T.dot(e), 1e-14, "prolong %r" % (pshape,))
        h.close()


# ------------------------------------------------------------------ per-kernel parity

CASES = [((256,), True, 3), ((64,), False, 2), ((32, 32), False, 2), ((64, 64), False, 2),
         ((8, 8, 8), False, 1), ((16, 16, 16), False, 2), ((32, 32, 32), False, 3)]


@pytest.mark.parametrize("flags", [0, _lib.FLAG_FORCE_CSR, _lib.FLAG_NO_FUSED])
@pytest.mark.parametrize("case", CASES, ids=lambda c: key(c[0]))
def test_kernels_vs_oracle(case, flags):
    shape, s1, gl = case
    A0 = orc.poisson_csr(shape, sparse_1d=s1)
    R = orc.restrictionList(shape, gl - 1, 8)
    A = orc.coeffecientList(A0, R)
    h = Hierarchy(A0, shape, gl - 1, 8, flags=flags)
    assert h.nlevels == len(A)
    rs = np.random.RandomState(7)
    for l in range(len(A)):
        Al = sp.csr_matrix(A[l])
        n = Al.shape[0]
        info = h.level_info(l)
        if flags & _lib.FLAG_FORCE_CSR:
            assert info["kind"] == "csr"
        x = rs.random_sample(n)
        b = rs.random_sample(n)
        close(h.matvec(x, l), Al.dot(x), 1e-14, "matvec L%d" % l)
        close(h.residual(l, b, x), b - Al.dot(x), 1e-14, "residual L%d" % l)
        assert abs(h.residual_norm(l, b, x) - np.linalg.norm(b - Al.dot(x))) <= 1e-13 * np.linalg.norm(b)
        for sweeps in (1, 3):
            close(h.smooth(l, b, x, sweeps, "jacobi", 0.8), orc.jacobi(Al, b, x.copy(), sweeps, 0.8),
                  JAC_RTOL, "jacobi L%d" % l)
            col = orc.colouring(shape, l, n)
            close(h.smooth(l, b, x, sweeps, "rbgs"), orc.rbgs(Al, b, x.copy(), sweeps, col),
                  RB_RTOL, "rbgs L%d" % l)
        if n <= 4096:
            close(h.smooth(l, b, x, 2, "gs"), orc.gaussSeidel_c(Al, b, x.copy(), 2), 1e-12, "lexgs L%d" % l)
        if l < len(A) - 1:
            Rl = R[l]
            close(h.residual_restrict(l, b, x), Rl.dot(b - Al.dot(x)), 1e-13, "residual_restrict L%d" % l)
            e = rs.random_sample(Rl.shape[0])
            close(h.prolong_correct(l, e, x), x + Rl.T.dot(e), 1e-14, "prolong L%d" % l)
            y = x + Rl.T.dot(e)
            close(h.prolong_correct_smooth(l, b, e, x, 2, "jacobi", 0.8), orc.jacobi(Al, b, y.copy(), 2, 0.8),
                  JAC_RTOL, "prolong+jacobi L%d" % l)
            close(h.prolong_correct_smooth(l, b, e, x, 1, "rbgs"),
                  orc.rbgs(Al, b, y.copy(), 1, orc.colouring(shape, l, n)), RB_RTOL, "prolong+rbgs L%d" % l)
        else:
            close(h.coarse_solve(b), orc.coarseSolve(Al, b), 1e-11, "coarse solve")
    h.close()


# shapes on which the structured (TMA-staged) kernels apply: many y-segments, several x-chunks, flat-index wraps
@pytest.mark.parametrize("shape,gl", [((256, 256), 2), ((1 << 15,), 3), ((1024, 1024), 3), ((4096, 4096), 1)])
def test_structured_kernels_vs_oracle(shape, gl):
    s1 = len(shape) == 1
    A0 = orc.poisson_csr(shape, sparse_1d=s1)
    big = A0.shape[0] > (1 << 21)
    if big:         # level 0 only: skip the CPU Galerkin product
        R, A = [orc.restriction(shape)], [A0]
    else:
        R = orc.restrictionList(shape, gl, 8)
        A = orc.coeffecientList(A0, R)
    h = Hierarchy(omg.operators.poisson_band(shape, sparse_1d=s1), shape, 6 if big else gl, 8)
    rs = np.random.RandomState(11)
    for l in range(1 if big else len(A) - 1):
        Al = sp.csr_matrix(A[l])
        n = Al.shape[0]
        x, b = rs.random_sample(n), rs.random_sample(n)
        col = orc.colouring(shape, l, n)
        for sweeps in (1, 2):
            close(h.smooth(l, b, x, sweeps, "rbgs"), orc.rbgs(Al, b, x.copy(), sweeps, col), RB_RTOL,
                  "rbgs %d sweeps L%d" % (sweeps, l))
        e = rs.random_sample(R[l].shape[0])
        y = x + R[l].T.dot(e)
        for sweeps in (1, 2):
            close(h.prolong_correct_smooth(l, b, e, x, sweeps, "rbgs"), orc.rbgs(Al, b, y.copy(), sweeps, col),
                  RB_RTOL, "prolong+rbgs %d sweeps L%d" % (sweeps, l))
        # the Jacobi-side structured kernels on the same levels (coarse 2-D levels: in-kernel column corrections)
        close(h.smooth(l, b, x, 2, "jacobi", 0.8), orc.jacobi(Al, b, x.copy(), 2, 0.8), JAC_RTOL, "jacobi L%d" % l)
        close(h.residual_restrict(l, b, x), R[l].dot(b - Al.dot(x)), 1e-13, "residual_restrict L%d" % l)
        close(h.prolong_correct_smooth(l, b, e, x, 1, "jacobi", 0.8), orc.jacobi(Al, b, y.copy(), 1, 0.8),
              JAC_RTOL, "prolong+jacobi L%d" % l)
    h.close()


@pytest.mark.parametrize("shape,gl", [((64, 64, 64), 2), ((64, 32, 64), 2), ((32, 16, 64), 2), ((128, 128, 128), 3),
                                      ((256, 8, 256), 3)])
def test_3d_single_pass_two_colour_sweep(shape, gl, monkeypatch):
    """k_rb3 (both colour half-sweeps of a 3-D level in one pass, plain and fused with the prolongation) against the
    oracle and against the two half-sweep launches (OMG_NO_RB3), on level 0 (pure band) and on Galerkin levels
    (class-corrected boundary points)."""
    A0 = orc.poisson_csr(shape)
    R = orc.restrictionList(shape, gl - 1, 8)
    A = orc.coeffecientList(A0, R)
    rs = np.random.RandomState(13)
    cases = []
    for l in range(len(A) - 1):
        n = A[l].shape[0]
        if n < (1 << 15):
            continue
        cases.append((l, rs.random_sample(n), rs.random_sample(n), rs.random_sample(R[l].shape[0])))
    outs = {}
    for on in (False, True):
        if on:
            monkeypatch.delenv("OMG_NO_RB3", raising=False)
        else:
            monkeypatch.setenv("OMG_NO_RB3", "1")
        h = Hierarchy(omg.operators.poisson_band(shape), shape, gl - 1, 8, flags=_lib.FLAG_NO_GRAPH)
        outs[on] = [[h.smooth(l, b, x, 1, "rbgs"), h.smooth(l, b, x, 2, "rbgs"),
                     h.prolong_correct_smooth(l, b, e, x, 1, "rbgs"), h.prolong_correct_smooth(l, b, e, x, 2, "rbgs")]
                    for (l, x, b, e) in cases]
        h.close()
    for ci, (l, x, b, e) in enumerate(cases):
        Al = sp.csr_matrix(A[l])
        col = orc.colouring(shape, l, Al.shape[0])
        y = x + R[l].T.dot(e)
        want = [orc.rbgs(Al, b, x.copy(), 1, col), orc.rbgs(Al, b, x.copy(), 2, col),
                orc.rbgs(Al, b, y.copy(), 1, col), orc.rbgs(Al, b, y.copy(), 2, col)]
        for i in range(4):
            close(outs[True][ci][i], want[i], RB_RTOL, "single-pass vs oracle, L%d case %d" % (l, i))
            close(outs[True][ci][i], outs[False][ci][i], 1e-13, "single-pass vs half-sweeps, L%d case %d" % (l, i))


@pytest.mark.parametrize("shape,gl", [((64, 64, 64), 2), ((64, 32, 64), 2), ((128, 128, 128), 3), ((256, 16, 256), 3),
                                      ((256, 32, 256), 3), ((512, 16, 512), 3)])
def test_3d_jacobi_sweep_and_restricted_residual_single_pass(shape, gl, monkeypatch):
    """k_jr3 (the last pre-smoothing Jacobi sweep and the restricted residual of openmg/__init__.py:201,209-210 in
    one pass over x) against the oracle and against the two separate kernels (OMG_NO_JR3): first/last chunks and
    segments (flat-index wraps into the neighbouring planes), one and two patch slots per thread, 64- to 512-wide
    rows, with and without preceding sweeps."""
    _single_pass_descent_case(shape, gl, monkeypatch)


@pytest.mark.parametrize("shape,gl,maxw", [((256, 256), 3, None), ((1024, 1024), 5, None), ((1024, 1024), 5, 256),
                                           ((2048, 2048), 6, 1024), ((1 << 15,), 3, None), ((1 << 20,), 8, None),
                                           (((1 << 21) + 4096,), 10, 512)])
def test_2d_1d_jacobi_sweep_and_restricted_residual_single_pass(shape, gl, maxw, monkeypatch):
    """k_jr2, the row-marching member of the same family (2-D level 0: offsets 1 and N+1, openmg/operators.py:221-241;
    1-D vectors viewed as rows of 2048), one and several x-chunks (OMG_RB_MAXW) and y-segments, row ends continuing
    into the next row."""
    if maxw:
        monkeypatch.setenv("OMG_RB_MAXW", str(maxw))
    _single_pass_descent_case(shape, gl, monkeypatch)


def _single_pass_descent_case(shape, gl, monkeypatch):
    s1 = len(shape) == 1
    A0 = sp.csr_matrix(orc.poisson_csr(shape, sparse_1d=s1))
    R = orc.restrictionList(shape, gl - 1, 8)
    n = A0.shape[0]
    rs = np.random.RandomState(29)
    x, b = rs.random_sample(n), rs.random_sample(n)
    outs = {}
    for on in (False, True):
        if on:
            monkeypatch.delenv("OMG_NO_JR3", raising=False)
        else:
            monkeypatch.setenv("OMG_NO_JR3", "1")
        h = Hierarchy(omg.operators.poisson_band(shape, sparse_1d=s1), shape, gl - 1, 8, flags=_lib.FLAG_NO_GRAPH)
        outs[on] = [h.smooth_residual_restrict(0, b, x, sweeps, "jacobi", 0.8) for sweeps in (1, 2, 3)]
        outs[on].append(h.smooth_residual_restrict(0, b, np.zeros(n), 1, "jacobi", 0.8))
        h.close()
    wants = [orc.jacobi(A0, b, x.copy(), sweeps, 0.8) for sweeps in (1, 2, 3)] + [orc.jacobi(A0, b, np.zeros(n), 1, 0.8)]
    for i, xw in enumerate(wants):
        rw = R[0].dot(b - A0.dot(xw))
        close(outs[True][i][0], xw, JAC_RTOL, "x, case %d" % i)
        close(outs[True][i][1], rw, 1e-13, "R r, case %d" % i)
        close(outs[True][i][0], outs[False][i][0], 1e-15, "x: single pass vs two kernels, case %d" % i)
        close(outs[True][i][1], outs[False][i][1], 1e-13, "R r: single pass vs two kernels, case %d" % i)
    # small residuals: the patch-sum form of the residual must not lose accuracy when b - A x cancels
    xs = rs.random_sample(n)
    bs = A0.dot(xs)
    h = Hierarchy(omg.operators.poisson_band(shape, sparse_1d=s1), shape, gl - 1, 8, flags=_lib.FLAG_NO_GRAPH)
    xg, rg = h.smooth_residual_restrict(0, bs, xs, 1, "jacobi", 0.8)
    h.close()
    close(xg, xs, 1e-14, "x already solves A x = b: the sweep leaves it alone")
    assert np.abs(rg).max() <= 1e-12 * np.abs(bs).max(), "x already solves A x = b: restricted residual at rounding level"


@pytest.mark.parametrize("shape,gl", [((64, 64, 64), 3), ((128, 32, 128), 3)])
def test_zero_start_descent_with_and_without_transform_pass(shape, gl, monkeypatch):
    """Levels >= 1 start from the zero iterate (openmg/__init__.py:191-192).  On levels with a uniform diagonal the
    fused sweep + residual + restriction kernel runs its 7-point pass on the staged b planes and scales the result
    (k_st3 MODE 4); OMG_NO_MODE4 selects the older form that first turns the staged planes into x = omega b / a_ii.
    Both against the oracle's cycles and against each other."""
    A_in = orc.poisson_csr(shape)
    _, b = seeded_problem(A_in)
    params = {'problemShape': shape, 'gridLevels': gl, 'preIterations': 1, 'postIterations': 1,
              'verbose': False, 'minSize': 8}
    R = orc.restrictionList(shape, gl - 1, 8)
    params['coarsestLevel'] = len(R)
    A = orc.coeffecientList(A_in, R)
    smooth = orc.make_smoother("jacobi", shape, 0.8)
    x = None
    for _ in range(2):
        x, info = orc.mgCycle(A, b, 0, R, params, initial=x, smooth=smooth)
    outs = {}
    for on in (False, True):
        if on:
            monkeypatch.delenv("OMG_NO_MODE4", raising=False)
        else:
            monkeypatch.setenv("OMG_NO_MODE4", "1")
        h = Hierarchy(omg.operators.poisson_band(shape), shape, gl - 1, 8, flags=_lib.FLAG_NO_GRAPH)
        outs[on] = h.solve(b, None, 1, 1, "jacobi", 0.8, 2, 0.0)[0]
        h.close()
    close(outs[True], x, JAC_RTOL, "scaled 7-point pass on b vs oracle")
    close(outs[False], x, JAC_RTOL, "transform pass vs oracle")
    close(outs[True], outs[False], 1e-14, "both forms")


def test_band_detection_reports_structure():
    h = Hierarchy(orc.poisson_csr((32, 32, 32)), (32, 32, 32), 2, 8)
    i0, i1 = h.level_info(0), h.level_info(1)
    assert i0["kind"] == "band" and i0["nexc"] == 0
    assert i1["kind"] == "band+exc" and 0 < i1["nexc"] < 0.3 * i1["n"]
    d, offs, coef = h.level_band(0)
    assert d == -12.0 and sorted(offs.tolist()) == [-1024, -32, -1, 1, 32, 1024] and set(coef.tolist()) == {1.0}
    d1, offs1, coef1 = h.level_band(1)
    assert d1 == -1.125 and set(coef1.tolist()) == {0.0625}
    h.close()


# ------------------------------------------------------------------ whole V-cycles vs the reference's own mgCycle

def test_vcycles_match_reference_goldens(gold_cycles):
    z, meta = gold_cycles
    for name, shape, sparse_flag, pshape, gl, smoother, pre, post, ncyc in meta["cycles"]:
        shape, pshape = tuple(shape), tuple(pshape)
        A_in = orc.poisson_csr(shape, sparse_1d=sparse_flag)
        _, b = seeded_problem(A_in)
        h = Hierarchy(A_in, pshape, gl - 1, 8)
        x, cyc, norm, hist = h.solve(b, None, pre, post, smoother, 0.8, ncyc, 0.0, want_history=True)
        tag = "%s/%s/%d%d" % (name, smoother, pre, post)
        tol = {"jacobi": JAC_RTOL, "rbgs": RB_RTOL, "gs": 1e-11}[smoother]
        assert cyc == ncyc
        np.testing.assert_allclose(hist, z[tag + "/norms"], rtol=100 * tol,
                                   atol=1e-12 * max(np.linalg.norm(b), 1.0), err_msg=tag)
        close(x, z[tag + "/x"], tol, tag)
        # one cycle at a time, chained through `initial`, must give the same iterate
        xi = None
        for _ in range(ncyc):
            xi, nrm = h.cycle(b, xi, 0, pre, post, smoother, 0.8)
        close(xi, x, 1e-14, tag + " chained")
        h.close()


def test_band_and_csr_inputs_agree_bitwise():
    for shape in ((128,), (32, 32), (16, 16, 16)):
        A_csr = orc.poisson_csr(shape)
        _, b = seeded_problem(A_csr)
        outs = []
        for A_in in (A_csr, omg.operators.poisson_band(shape)):
            h = Hierarchy(A_in, shape, 2, 8)
            outs.append(h.solve(b, None, 1, 1, "jacobi", 0.8, 3, 0.0)[0])
            h.close()
        np.testing.assert_array_equal(outs[0], outs[1])


def test_graph_and_direct_launch_agree_bitwise():
    shape = (16, 16, 16)
    A = orc.poisson_csr(shape)
    _, b = seeded_problem(A)
    outs = []
    for flags in (0, _lib.FLAG_NO_GRAPH):
        h = Hierarchy(A, shape, 2, 8, flags=flags)
        outs.append(h.solve(b, None, 2, 1, "rbgs", 0.8, 5, 0.0)[0])
        h.close()
    np.testing.assert_array_equal(outs[0], outs[1])


@pytest.mark.parametrize("shape,gl,smoother", [((64, 64, 64), 3, "jacobi"), ((64, 64, 64), 3, "rbgs"),
                                              ((256, 256), 4, "jacobi"), ((1 << 16,), 10, "jacobi"),
                                              ((1 << 16,), 10, "rbgs")])
def test_midsize_cycles_vs_oracle(shape, gl, smoother):
    s1 = len(shape) == 1
    A_in = orc.poisson_csr(shape, sparse_1d=s1)
    _, b = seeded_problem(A_in)
    params = {'problemShape': shape, 'gridLevels': gl, 'preIterations': 1, 'postIterations': 1,
              'verbose': False, 'minSize': 8}
    R = orc.restrictionList(shape, gl - 1, 8)
    params['coarsestLevel'] = len(R)
    A = orc.coeffecientList(A_in, R)
    smooth = orc.make_smoother(smoother, shape, 0.8)
    x, norms = None, []
    for _ in range(3):
        x, info = orc.mgCycle(A, b, 0, R, params, initial=x, smooth=smooth)
        norms.append(info['norm'])
    Ab = omg.operators.poisson_band(shape, sparse_1d=s1)
    h = Hierarchy(Ab, shape, gl - 1, 8)
    xg, cyc, norm, hist = h.solve(b, None, 1, 1, smoother, 0.8, 3, 0.0, want_history=True)
    tol = JAC_RTOL if smoother == "jacobi" else RB_RTOL
    close(xg, x, tol, "x")
    np.testing.assert_allclose(hist, norms, rtol=1e-9, atol=1e-12 * np.linalg.norm(b))
    h.close()


def test_converged_solution_matches_reference_gs():
    """Two-colour GS replaces the reference's lexicographic GS: compare the converged
    solution (1e-10 relative) and that the residual-vs-cycle curve decays like the reference's."""
    shape = (16, 16, 16)
    A = orc.poisson_csr(shape)
    u, b = seeded_problem(A)
    h = Hierarchy(A, shape, 1, 8)
    x, cyc, norm, hist = h.solve(b, None, 1, 1, "rbgs", 0.8, 40, 0.0, want_history=True)
    close(x, u, 1e-10, "converged rbgs vs true solution")
    xr = orc.mgSolve(A, b, {'problemShape': shape, 'gridLevels': 2, 'cycles': 40, 'threshold': 0,
                            'preIterations': 1, 'postIterations': 1, 'smoother': 'gs'})
    close(x, xr, 1e-10, "converged rbgs vs reference lexicographic GS")
    rate = (hist[7] / hist[2]) ** (1 / 5.0)
    assert rate < 0.2                    # reference lex GS: ~0.04 per cycle; 2-colour: ~0.03
    h.close()


def test_split_row_and_tiling_variants_agree(monkeypatch):
    """The stencil kernel's tiling knobs (debug environment variables read at launch time) select other
    code paths — split rows with per-row TMA copies, 512-thread CTAs, deeper rings — which must all give
    the same iterates as the default tiling."""
    shape, gl = (128, 128, 128), 4
    A = omg.operators.poisson_band(shape)
    u = np.random.RandomState(0).random_sample(A.n)
    ref = None
    for env in ({}, {"OMG_ST_XW": "64"}, {"OMG_ST_NT": "512"}, {"OMG_ST_NT": "128", "OMG_ST_TY": "4"},
                {"OMG_ST_NS": "4", "OMG_ST_ZL": "6"}):
        for k in ("OMG_ST_XW", "OMG_ST_NT", "OMG_ST_TY", "OMG_ST_NS", "OMG_ST_ZL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        h = Hierarchy(A, shape, gl - 1, 8)
        b = h.matvec(u, 0)
        outs = [h.solve(b, None, 1, 1, sm, 0.8, 3, 0.0)[0] for sm in ("jacobi", "rbgs")]
        h.close()
        if ref is None:
            ref = outs
        else:
            for a, r in zip(outs, ref):
                close(a, r, 1e-13, "tiling %r" % (env,))


def test_structured_2d_and_1d_paths_vs_generic():
    """2-D and 1-D band levels run on the row-marching TMA kernel; OMG_FLAG_NO_FUSED forces the generic
    kernels: same iterates, Jacobi and two-colour."""
    for shape, gl, s1 in (((512, 512), 4, False), ((1 << 18,), 9, True)):
        A = omg.operators.poisson_band(shape, sparse_1d=s1)
        u = np.random.RandomState(1).random_sample(A.n)
        res = {}
        for flags in (0, _lib.FLAG_NO_FUSED):
            h = Hierarchy(A, shape, gl - 1, 8, flags=flags)
            b = h.matvec(u, 0)
            res[flags] = [h.solve(b, None, pre, post, sm, 0.8, 3, 0.0)[0]
                          for sm in ("jacobi", "rbgs") for (pre, post) in ((1, 1), (2, 0))]
            h.close()
        for a, r in zip(res[0], res[_lib.FLAG_NO_FUSED]):
            close(a, r, 1e-13, "structured vs generic %r" % (shape,))


def test_general_sparse_matrix_csr_path():
    """An arbitrary (non-Poisson, non-dyadic, unsorted, with explicit zeros and duplicates) coefficient
    matrix: every level is generic CSR.  Galerkin patterns bit-exact, values to rounding, cycles vs oracle."""
    rs = np.random.RandomState(11)
    for pshape, gl in (((64, 64), 2), ((512,), 3), ((8, 8, 8), 1)):
        n = int(np.prod(pshape))
        M = sp.random(n, n, density=6.0 / n, random_state=rs, format="coo")
        M = M + M.T + sp.diags(np.full(n, 20.0) + rs.random_sample(n))
        M = sp.coo_matrix(M)
        perm = rs.permutation(M.nnz)                       # unsorted COO with a few explicit zeros / duplicates
        rows = np.concatenate([M.row[perm], [0, 1, 1]])
        cols = np.concatenate([M.col[perm], [n - 1, 2, 2]])
        vals = np.concatenate([M.data[perm], [0.0, 0.25, -0.25]])
        A_in = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))   # duplicates summed by scipy on conversion
        A_ref = sp.csr_matrix(A_in)
        R = orc.restrictionList(pshape, gl - 1, 8)
        Aor = orc.coeffecientList(A_ref, R)
        h = Hierarchy(A_in, pshape, gl - 1, 8)
        assert h.nlevels == len(Aor)
        for l in range(1, h.nlevels):
            got, want = h.export_A(l), orc.canonical_csr(Aor[l])
            assert h.level_info(l)["kind"] == "csr"
            np.testing.assert_array_equal(got.indptr, want.indptr)
            np.testing.assert_array_equal(got.indices, want.indices)
            np.testing.assert_allclose(got.data, want.data, rtol=1e-13, atol=1e-15)
        u, b = seeded_problem(A_ref)
        params = {'coarsestLevel': len(R), 'preIterations': 1, 'postIterations': 1, 'verbose': False}
        for smoother, tol in (("jacobi", JAC_RTOL), ("rbgs", RB_RTOL), ("gs", 1e-11)):
            sm = orc.make_smoother(smoother, pshape, 0.8)
            xo = None
            for _ in range(3):
                xo, info = orc.mgCycle(Aor, b, 0, R, params, initial=xo, smooth=sm)
            x, cyc, norm, hist = h.solve(b, None, 1, 1, smoother, 0.8, 3, 0.0, want_history=True)
            close(x, xo, tol, "general matrix %r %s" % (pshape, smoother))
            assert abs(norm - info['norm']) <= 1e-9 * max(info['norm'], 1e-30) + 1e-12
        h.close()


def test_dense_and_matrix_inputs():
    """ndarray / np.matrix / COO / CSC inputs and (N,1) right-hand sides give the same result as CSR."""
    shape = (16, 16)
    A = orc.poisson_csr(shape)
    u, b = seeded_problem(A)
    ref = omg.mgSolve(A, b, {'problemShape': shape, 'gridLevels': 2, 'cycles': 3, 'smoother': 'jacobi'})
    for Ain, bb in ((A.toarray(), b), (np.asmatrix(A.toarray()), b.reshape(-1, 1)), (A.tocoo(), b), (A.tocsc(), b)):
        x = omg.mgSolve(Ain, bb, {'problemShape': shape, 'gridLevels': 2, 'cycles': 3, 'smoother': 'jacobi'})
        assert x.shape == (256,)
        np.testing.assert_array_equal(x, ref)
