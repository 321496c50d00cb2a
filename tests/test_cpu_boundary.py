"""CPU: the C-ABI library loads and exports every symbol include/omg_b200.h declares
(no compute calls), the product path fails loudly without a device, and the host-side
problem generators match the reference goldens."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, assert_same_csr, key

import openmg_b200
from openmg_b200 import _lib


def _header_functions():
    src = open(os.path.join(ROOT, "include", "omg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(omg_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), "libomg_b200.so does not export %s" % name
    # and the ctypes table covers the whole header
    assert sorted(_lib.SIGNATURES) == names


def test_ctypes_table_matches_header_prototypes():
    """Every prototype of include/omg_b200.h has the same number of parameters as its ctypes signature."""
    src = open(os.path.join(ROOT, "include", "omg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(?:int|void|const char \*)\s*(omg_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(protos) >= 30
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert len(_lib.SIGNATURES[name][1]) == n, (name, n, len(_lib.SIGNATURES[name][1]))


def test_product_path_fails_loudly_without_device():
    import ctypes
    lib = _lib.load()
    has_gpu = lib.omg_init(-1) == 0
    if has_gpu:
        pytest.skip("a CUDA device is present")
    assert b"no CPU fallback" in lib.omg_last_error()
    A = openmg_b200.operators.poisson((16,), sparse=True)
    with pytest.raises(RuntimeError):
        openmg_b200.mgSolve(A, np.ones(16), {'problemShape': (16,), 'gridLevels': 2, 'cycles': 1})
    with pytest.raises(RuntimeError):
        openmg_b200.smooth(A, np.ones(16), np.zeros(16), 1)
    with pytest.raises(RuntimeError):
        openmg_b200.operators.restriction((8,))
    h = ctypes.c_void_p()
    shape = np.array([16], dtype=np.int64)
    rc = lib.omg_hierarchy_create_band(ctypes.byref(h), 1, _lib.i64(shape), 1, 8, 16, 4.0, 0, None, None, 0)
    assert rc == _lib.ENODEV


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "openmg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_poisson_generators_match_reference(gold_operators):
    z, meta = gold_operators
    ops = openmg_b200.operators
    for shape, sparse_flag in meta["poisson"]:
        shape = tuple(shape)
        pre = "P/%s/%d" % (key(shape), int(sparse_flag))
        A = ops.poisson(shape, sparse=sparse_flag)
        if not sparse_flag:
            assert isinstance(A, np.ndarray) and A.shape == (np.prod(shape),) * 2
        assert_same_csr(A, z, pre)
        if len(shape) > 1 or not sparse_flag:
            assert_same_csr(ops.poisson_band(shape).tocsr(), z, pre)
    assert_same_csr(ops.poisson(8, sparse=True), z, "P/8/1")           # int shape (openmg/operators.py:264-265)
    assert_same_csr(ops.poisson((4, 4), sparse=True), z, "P/4x4/0")    # extension: sparse 2-D/3-D
    with pytest.raises(ValueError):
        ops.poisson((1, 2, 3, 4))
    assert openmg_b200.tools.poisson is ops.poisson


def test_dict_helpers_and_defaults():
    t = openmg_b200.tools
    d = {'a': 1}
    assert t.dictUpdateNoClobber({'a': 2, 'b': 3}, d) is d and d == {'a': 1, 'b': 3}
    assert t.product((2, 3, 4)) == 24
    ref_defaults = {'problemShape': (200,), 'gridLevels': 2, 'verbose': False, 'threshold': 0.1, 'cycles': 0,
                    'preIterations': 1, 'postIterations': 0, 'dense': False, 'giveInfo': False, 'minSize': 8}
    for k, v in ref_defaults.items():
        assert openmg_b200.defaults[k] == v
    assert openmg_b200.mg_solve is openmg_b200.mgSolve


def test_openmg_alias_package():
    import openmg
    assert openmg.mgSolve is openmg_b200.mgSolve
    import openmg.operators as ops
    assert ops is openmg_b200.operators


def test_stop_rule_matches_the_reference_semantics():
    """The package's stop test against a literal restatement of openmg/__init__.py:121-130 on a parameter sweep."""
    import itertools
    import openmg_b200 as omg

    def reference_rule(parameters, cycle, norm):
        cap = 'cycles' in parameters and parameters['cycles'] > 0 and cycle >= parameters['cycles']
        thr = 'threshold' in parameters and parameters['threshold'] > 0 and norm < parameters['threshold']
        return cap or thr

    for cycles, threshold, cycle, norm in itertools.product((None, -1, 0, 1, 3), (None, -1.0, 0.0, 0.5),
                                                            (1, 2, 3, 4), (0.1, 0.5, 0.7)):
        prm = {}
        if cycles is not None:
            prm['cycles'] = cycles
        if threshold is not None:
            prm['threshold'] = threshold
        assert omg._stop_rule(prm, cycle, norm) == reference_rule(prm, cycle, norm), (prm, cycle, norm)


def test_restriction_list_depth_rule_host_logic(gold_operators, monkeypatch):
    """restrictionList's depth rule is host code around the device-built R: check it on the CPU with the oracle's
    restriction() standing in for the device call (the -m gpu suite checks the real thing)."""
    import oracle.openmg_oracle as orc
    import openmg_b200.operators as ops
    monkeypatch.setattr(ops, "restriction", orc.restriction)
    _, meta = gold_operators
    for shape, cl, ms, shapes in meta["rlist"]:
        Rl = ops.restrictionList(tuple(shape), cl, ms)
        assert [list(r.shape) for r in Rl] == shapes, (shape, cl, ms)
        assert Rl.problemShape == tuple(shape)
