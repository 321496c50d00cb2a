"""Generate tests/golden/*.npz from the REFERENCE ITSELF.

Runs only in the build container (needs /root/reference, loaded in memory by
tools/ref_shim.py — nothing from the reference is written into the repo).
The committed .npz files are the pin for oracle/ (and, through it, for the
CUDA path) on the GPU box, where the reference does not exist.

    python tools/make_golden.py

Inputs are always regenerated from a seed by the tests:
    u = np.random.RandomState(seed).random_sample(N);  b = A @ u
so only outputs are stored.
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

RESTRICTION_SHAPES = [(8,), (9,), (25,), (5,), (64,), (4, 4), (6, 6), (5, 5), (3, 3), (4, 6), (16, 16),
                      (4, 4, 4), (6, 6, 6), (3, 3, 3), (4, 6, 8), (8, 8, 8), (6, 4, 6)]
RESTRICTION_ERRORS = [(6, 4), (8, 6, 4), (2, 2), (2,), (3,), (4, 4, 4, 4), (1, 1, 1)]
POISSON_CASES = [((8,), False), ((8,), True), ((100,), True), ((4, 4), False), ((4, 6), False),
                 ((6, 4), False), ((16, 16), False), ((4, 4, 4), False), ((2, 3, 4), False),
                 ((4, 3, 2), False), ((8, 8, 8), False)]
RLIST_CASES = [((1024,), 23, 23), ((512,), 2, 8), ((200,), 2, 30), ((8, 8, 8), 2, 8), ((16, 16, 16), 2, 8),
               ((32, 32), 2, 8), ((64, 64), 1, 8), ((64, 64), 2, 8), ((12, 12, 12), 3, 8), ((36,), 1, 8),
               ((16, 16), 5, 8), ((1 << 14,), 19, 8)]
# (name, matrix spec, problemShape, gridLevels, extra parameters)
CYCLE_CASES = [
    ("1d_sparse_512", ("poisson", (512,), True), (512,), 3, {}),
    ("1d_dense_36", ("poisson", (36,), False), (36,), 2, {}),
    ("2d_64", ("poisson", (64, 64), False), (64, 64), 2, {}),
    ("2d_32", ("poisson", (32, 32), False), (32, 32), 3, {}),
    ("3d_8", ("poisson", (8, 8, 8), False), (8, 8, 8), 3, {}),
    ("3d_16", ("poisson", (16, 16, 16), False), (16, 16, 16), 2, {}),
    ("testa_1d_as_3d", ("poisson", (1728,), False), (12, 12, 12), 4, {}),
]


def csr_parts(M):
    M = sparse.csr_matrix(M).copy()
    M.sum_duplicates()
    M.eliminate_zeros()
    M.sort_indices()
    return dict(indptr=M.indptr.astype(np.int64), indices=M.indices.astype(np.int64),
                data=M.data.astype(np.float64), shape=np.array(M.shape, dtype=np.int64))


def put(store, prefix, parts):
    for k, v in parts.items():
        store[prefix + "/" + k] = v


def key(shape):
    return "x".join(str(s) for s in shape)


def gen_operators(ref):
    store = {}
    meta = {"restriction_shapes": [], "restriction_errors": [], "poisson": [], "rlist": []}
    for shape in RESTRICTION_SHAPES:
        R = ref.operators.restriction(shape)
        assert R.dtype == np.float64
        put(store, "R/" + key(shape), csr_parts(R))
        # raw (as returned) index order must already be sorted: check & record
        store["R/" + key(shape) + "/raw_sorted"] = np.array(bool(R.has_sorted_indices))
        meta["restriction_shapes"].append(list(shape))
    for shape in RESTRICTION_ERRORS:
        try:
            ref.operators.restriction(shape)
            err = "none"
        except Exception as e:  # noqa: BLE001
            err = type(e).__name__
        meta["restriction_errors"].append([list(shape), err])
    for shape, sp in POISSON_CASES:
        A = ref.operators.poisson(shape, sparse=sp)
        put(store, "P/%s/%d" % (key(shape), int(sp)), csr_parts(A))
        meta["poisson"].append([list(shape), bool(sp)])
    for err_shape in [(1, 2, 3, 4)]:
        try:
            ref.operators.poisson(err_shape)
            raise AssertionError
        except ValueError:
            pass
    for shape, cl, ms in RLIST_CASES:
        Rl = ref.operators.restrictionList(shape, cl, ms)
        meta["rlist"].append([list(shape), cl, ms, [list(r.shape) for r in Rl]])
    store["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "operators.npz"), **store)
    print("operators.npz: %d arrays" % len(store))


def build_matrix(ref, spec):
    kind, shape, sp = spec
    assert kind == "poisson"
    return ref.operators.poisson(shape, sparse=sp)


def gen_galerkin(ref):
    store = {}
    meta = []
    for name, spec, pshape, gl, _ in CYCLE_CASES:
        A_in = build_matrix(ref, spec)
        R = ref.operators.restrictionList(pshape, gl - 1, 8)
        A = ref.operators.coeffecientList(A_in, R)
        for l, Al in enumerate(A):
            if l == 0 and Al.shape[0] > 2048:
                continue  # level 0 is the generator output, covered by operators.npz
            put(store, "%s/A%d" % (name, l), csr_parts(Al))
        meta.append([name, list(spec[1]), bool(spec[2]), list(pshape), gl, len(A),
                     [int(a.shape[0]) for a in A]])
    store["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "galerkin.npz"), **store)
    print("galerkin.npz: %d arrays" % len(store))


def run_cycles(ref, A_in, b, pshape, gl, pre, post, ncycles, smooth_fn=None):
    """The loop of openmg/__init__.py:93-138 with per-cycle norms recorded,
    driving the reference's own restrictionList/coeffecientList/mgCycle."""
    params = {'problemShape': pshape, 'gridLevels': gl, 'preIterations': pre, 'postIterations': post,
              'verbose': False, 'minSize': 8, 'coarsestLevel': gl - 1}
    R = ref.operators.restrictionList(pshape, params['coarsestLevel'], params['minSize'])
    params['coarsestLevel'] = len(R)
    A = ref.operators.coeffecientList(A_in, R)
    saved = ref.smooth
    if smooth_fn is not None:
        sizes = [a.shape[0] for a in A]

        def patched(Al, bl, xl, iterations, verbose=False):
            return smooth_fn(Al, bl, xl, iterations, level=sizes.index(Al.shape[0]))
        ref.__dict__['smooth'] = patched
    try:
        norms = []
        x = None
        for _ in range(ncycles):
            x, info = ref.mgCycle(A, b, 0, R, params, initial=x)
            norms.append(float(info['norm']))
    finally:
        ref.__dict__['smooth'] = saved
    return np.asarray(x).ravel(), np.array(norms)


def gen_cycles(ref):
    import oracle.openmg_oracle as orc
    store = {}
    meta = []
    ncycles = 4
    for name, spec, pshape, gl, _ in CYCLE_CASES:
        A_in = build_matrix(ref, spec)
        A_csr = sparse.csr_matrix(A_in)
        N = A_csr.shape[0]
        u = np.random.RandomState(0).random_sample(N)
        b = np.asarray(A_csr.dot(u)).ravel()
        for smoother in ("gs", "jacobi", "rbgs"):
            for (pre, post) in ((1, 0), (1, 1), (2, 1)):
                if smoother == "gs" and (pre, post) == (2, 1) and N > 2000:
                    continue
                if smoother == "gs":
                    fn = None  # the reference's own lexicographic GS, untouched
                else:
                    fn = orc.make_smoother(smoother, pshape, 0.8)
                x, norms = run_cycles(ref, A_in, b.copy(), pshape, gl, pre, post, ncycles, fn)
                tag = "%s/%s/%d%d" % (name, smoother, pre, post)
                store[tag + "/x"] = x
                store[tag + "/norms"] = norms
                meta.append([name, list(spec[1]), bool(spec[2]), list(pshape), gl, smoother, pre, post, ncycles])
                print("  %-28s norms %s" % (tag, " ".join("%.3e" % v for v in norms)))
    # full mgSolve semantics (stop rule, info dict) on the reference
    msolve = []
    for (N, gl, cycles, thr) in ((36, 2, 3, 1e-10), (36, 2, 0, 8e-3), (100, 3, 10, 1e-2)):
        if N == 100:
            A = ref.operators.poisson(N, sparse=True)
            u_true = np.array([np.sin(x / 10.0) for x in np.linspace(0, 20, N)])
        else:
            A = ref.operators.poisson((N,))
            u_true = np.sin(np.array(range(int(N))) * 3.0 / N).T
        b = np.asarray(ref.tools.flexibleMmult(A, u_true)).ravel()
        params = {'problemShape': (N,), 'gridLevels': gl, 'cycles': cycles, 'threshold': thr, 'giveInfo': True}
        x, info = ref.mgSolve(A, b, params)
        tag = "mgsolve/%d_%d_%d_%g" % (N, gl, cycles, thr)
        store[tag + "/x"] = np.asarray(x).ravel()
        msolve.append([N, gl, cycles, thr, int(info['cycle']), float(info['norm']),
                       int(params['coarsestLevel'])])
    store["meta"] = np.array(json.dumps({"cycles": meta, "mgsolve": msolve}))
    np.savez_compressed(os.path.join(GOLD, "cycles.npz"), **store)
    print("cycles.npz: %d arrays" % len(store))


def gen_smoothers(ref):
    """Reference gaussSeidel (openmg/solvers.py:34-75) on small seeded systems."""
    store = {}
    meta = []
    for shape, sp in (((64,), True), ((12, 12), False), ((6, 6, 6), False), ((12,), False)):
        A = ref.operators.poisson(shape, sparse=sp)
        N = A.shape[0]
        rs = np.random.RandomState(1)
        b = rs.random_sample(N)
        for As, tag in ((sparse.csr_matrix(A), "csr"), (np.asarray(sparse.csr_matrix(A).todense()), "dense")):
            for it in (1, 3):
                x = rs.random_sample(N) if it == 3 else np.zeros(N)
                x0 = x.copy()
                out = ref.solvers.gaussSeidel(As, b, x, iterations=it)
                assert out is x
                k = "gs/%s/%d/%s/%d" % (key(shape), int(sp), tag, it)
                store[k + "/x0"] = x0
                store[k + "/b"] = b
                store[k + "/x"] = x.copy()
                meta.append([list(shape), bool(sp), tag, it])
    # threshold mode (test_gs_thresh)
    A = ref.operators.poisson((12, 12))
    b = np.random.RandomState(2).random_sample(144)
    x = ref.solvers.smoothToThreshold(A, b, np.zeros(144), 1e-4)
    store["gs_thresh/b"] = b
    store["gs_thresh/x"] = x
    # coarseSolve
    for shape, sp in (((64,), True), ((8, 8), False), ((4, 4, 4), False)):
        A = ref.operators.poisson(shape, sparse=sp)
        b = np.random.RandomState(3).random_sample(A.shape[0])
        xs = ref.solvers.coarseSolve(sparse.csr_matrix(A), b.reshape(-1, 1))
        store["coarse/%s/%d/x" % (key(shape), int(sp))] = xs
        store["coarse/%s/%d/b" % (key(shape), int(sp))] = b
    store["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "smoothers.npz"), **store)
    print("smoothers.npz: %d arrays" % len(store))


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_shim.load()
    gen_operators(ref)
    gen_galerkin(ref)
    gen_smoothers(ref)
    gen_cycles(ref)
    with open(os.path.join(GOLD, "README.md"), "w") as f:
        f.write("# Golden vectors\n\nGenerated by `python tools/make_golden.py` in the build container from the\n"
                "UNMODIFIED reference (`/root/reference/openmg`, loaded in memory through `tools/ref_shim.py`),\n"
                "numpy %s / scipy %s.\n\n"
                "* `operators.npz` — `restriction(shape)` CSR, `poisson(shape)` CSR, `restrictionList` shapes, error types.\n"
                "* `galerkin.npz`  — `coeffecientList` levels in canonical (sorted, zero-free) CSR.\n"
                "* `smoothers.npz` — reference `gaussSeidel` / `smoothToThreshold` / `coarseSolve` outputs.\n"
                "* `cycles.npz`    — per-cycle residual norms and final iterate of 4 V-cycles of the reference's `mgCycle`:\n"
                "  `gs` = reference smoother untouched; `jacobi`/`rbgs` = reference `mgCycle` with its plug-in point\n"
                "  `openmg.smooth` replaced by `oracle.make_smoother(...)`.  Inputs: `u=RandomState(0).random_sample(N)`, `b=A@u`.\n"
                % (np.__version__, __import__('scipy').__version__))


if __name__ == "__main__":
    main()
