"""Workload for compute-sanitizer (tools/sanitize.sh): every structured kernel family once, on the smallest shapes
that still select them — the TMA/mbarrier rings of k_st3 / k_st2 / k_st2rb / k_rb3 / k_jr3 (plain and prolongation-fused,
band and class-corrected levels) and whole V-cycles with both smoothers, direct launches (no CUDA graph)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openmg_b200 as omg                     # noqa: E402
from openmg_b200 import _lib                  # noqa: E402
from openmg_b200.hierarchy import Hierarchy   # noqa: E402

cases = [((64, 32, 64), 2), ((128, 64, 128), 3), ((256, 256), 2), ((1 << 14,), 4)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
rs = np.random.RandomState(3)
for shape, gl in cases:
    A = omg.operators.poisson_band(shape, sparse_1d=(len(shape) == 1))
    h = Hierarchy(A, shape, gl - 1, 8, flags=_lib.FLAG_NO_GRAPH)
    for l in range(h.nlevels - 1):
        n, nc = h.n(l), h.n(l + 1)
        x, b, e = rs.random_sample(n), rs.random_sample(n), rs.random_sample(nc)
        h.smooth(l, b, x, 1, "jacobi", 0.8)
        h.smooth(l, b, x, 1, "rbgs")
        h.residual_restrict(l, b, x)
        h.smooth_residual_restrict(l, b, x, 1, "jacobi", 0.8)      # level 0: k_jr3 (sweep + residual + restriction)
        h.prolong_correct_smooth(l, b, e, x, 1, "jacobi", 0.8)
        h.prolong_correct_smooth(l, b, e, x, 1, "rbgs")
        h.residual_norm(l, b, x)
    b = h.matvec(rs.random_sample(A.n), 0)
    for sm in ("jacobi", "rbgs"):
        x, cyc, norm, _ = h.solve(b, None, 1, 1, sm, 0.8, 2, 0.0)
        print(shape, sm, "norm after 2 cycles %.3e" % norm, flush=True)
    h.close()
print("sanitize_driver: done")
