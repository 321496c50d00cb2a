"""CPU, build container only: the oracle against the LIVE reference (/root/reference loaded in memory by
tools/ref_shim.py).  Skipped where the reference does not exist (the GPU box); the committed goldens of
tests/test_oracle_golden.py are the portable form of the same check."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shim  # noqa: E402
import oracle.openmg_oracle as orc  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


def same(a, b):
    a, b = orc.canonical_csr(a), orc.canonical_csr(b)
    return a.shape == b.shape and (a != b).nnz == 0 and np.array_equal(a.indices, b.indices)


def test_restriction_and_errors_live(ref):
    rs = np.random.RandomState(0)
    shapes = [(int(rs.randint(4, 40)),) for _ in range(6)] + [(2 * int(rs.randint(2, 9)),) * 2 for _ in range(4)] + \
             [(2 * int(rs.randint(2, 5)),) * 3 for _ in range(3)] + [(4, 6), (6, 4), (3, 5), (5, 3), (4, 4, 6), (2, 2), (7, 7, 7)]
    for shape in shapes:
        try:
            want = ref.operators.restriction(shape)
        except Exception as e:  # noqa: BLE001
            with pytest.raises(type(e)):
                orc.restriction(shape)
            continue
        assert same(orc.restriction(shape), want), shape


def test_galerkin_and_cycles_live(ref):
    for shape, gl, sparse_flag in (((48,), 3, True), ((12, 12), 2, False), ((6, 6, 6), 2, False)):
        A_ref = ref.operators.poisson(shape, sparse=sparse_flag)
        A = orc.poisson_csr(shape, sparse_1d=sparse_flag)
        assert same(A, sp.csr_matrix(A_ref))
        R_ref = ref.operators.restrictionList(shape, gl - 1, 4)
        R = orc.restrictionList(shape, gl - 1, 4)
        assert len(R) == len(R_ref)
        for a, b in zip(orc.coeffecientList(A, R), ref.operators.coeffecientList(A_ref, R_ref)):
            assert same(a, b)
        b = np.random.RandomState(3).random_sample(A.shape[0])
        p1 = {'problemShape': shape, 'gridLevels': gl, 'cycles': 3, 'threshold': 0, 'minSize': 4, 'giveInfo': True}
        p2 = dict(p1)
        x_ref, i_ref = ref.mgSolve(A_ref, b.copy(), p1)
        x, i = orc.mgSolve(A, b.copy(), p2, smooth=orc.make_smoother('gs', fast=False))
        np.testing.assert_allclose(x, np.asarray(x_ref).ravel(), rtol=1e-12, atol=1e-14)
        assert i['cycle'] == i_ref['cycle'] and np.isclose(i['norm'], i_ref['norm'], rtol=1e-10)
        assert p1['coarsestLevel'] == p2['coarsestLevel']
