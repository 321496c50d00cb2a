"""One-GPU check of the wide-row code paths the 8-GPU 1024^3 run uses (1024-wide level 0, 512-wide level 1 with
in-kernel boundary corrections): structured kernels vs the generic ones (OMG_FLAG_NO_FUSED), same iterates.

    python tools/wide_row_check.py [--shape 1024 64 1024] [--gl 5]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openmg_b200 as omg                     # noqa: E402
from openmg_b200 import _lib                  # noqa: E402
from openmg_b200.hierarchy import Hierarchy   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", type=int, nargs="+", default=[1024, 64, 1024])
ap.add_argument("--gl", type=int, default=5)
a = ap.parse_args()
shape = tuple(a.shape)
A = omg.operators.poisson_band(shape)
u = np.random.RandomState(0).random_sample(A.n)
res = {}
for flags in (0, _lib.FLAG_NO_FUSED):
    h = Hierarchy(A, shape, a.gl, 8, flags=flags)
    if flags == 0:
        print("levels:", [(h.level_info(l)["n"], h.level_info(l)["kind"]) for l in range(h.nlevels)])
    b = h.matvec(u, 0)
    res[flags] = [h.solve(b, None, pre, post, sm, 0.8, 3, 0.0)[0]
                  for sm in ("jacobi", "rbgs") for (pre, post) in ((1, 1), (2, 0))]
    h.close()
bad = 0
for i, (x, r) in enumerate(zip(res[0], res[_lib.FLAG_NO_FUSED])):
    err = np.abs(x - r).max() / np.abs(r).max()
    print("case %d: structured vs generic max rel err %.2e%s" % (i, err, "  FAIL" if err > 1e-12 else ""))
    bad += err > 1e-12
sys.exit(1 if bad else 0)
