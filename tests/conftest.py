import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load(name):
    z = np.load(os.path.join(GOLD, name), allow_pickle=False)
    return z, json.loads(str(z["meta"]))


@pytest.fixture(scope="session")
def gold_operators():
    return _load("operators.npz")


@pytest.fixture(scope="session")
def gold_galerkin():
    return _load("galerkin.npz")


@pytest.fixture(scope="session")
def gold_smoothers():
    return _load("smoothers.npz")


@pytest.fixture(scope="session")
def gold_cycles():
    return _load("cycles.npz")


def key(shape):
    return "x".join(str(s) for s in shape)


def csr_from(z, prefix):
    import scipy.sparse as sp
    shape = tuple(int(v) for v in z[prefix + "/shape"])
    return sp.csr_matrix((z[prefix + "/data"], z[prefix + "/indices"].astype(np.int32),
                          z[prefix + "/indptr"].astype(np.int32)), shape=shape)


def assert_same_csr(M, z, prefix, values_exact=True):
    """Bit-exact pattern (and, by default, values) against a golden CSR."""
    import scipy.sparse as sp
    M = sp.csr_matrix(M).copy()
    M.sum_duplicates()
    M.eliminate_zeros()
    M.sort_indices()
    assert tuple(M.shape) == tuple(int(v) for v in z[prefix + "/shape"]), prefix
    np.testing.assert_array_equal(M.indptr.astype(np.int64), z[prefix + "/indptr"], err_msg=prefix)
    np.testing.assert_array_equal(M.indices.astype(np.int64), z[prefix + "/indices"], err_msg=prefix)
    if values_exact:
        np.testing.assert_array_equal(M.data, z[prefix + "/data"], err_msg=prefix)
    else:
        np.testing.assert_allclose(M.data, z[prefix + "/data"], rtol=1e-13, atol=0, err_msg=prefix)


def seeded_problem(A, seed=0):
    """u = RandomState(seed).random_sample(N); b = A @ u."""
    N = A.shape[0]
    u = np.random.RandomState(seed).random_sample(N)
    return u, np.asarray(A.dot(u)).ravel()
