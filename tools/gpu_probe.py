"""Small driver for ncu / experiments: a few direct-launched V-cycles on one shape.

    python tools/gpu_probe.py [--shape 512 512 512] [--gl 5] [--cycles 2] [--smoother jacobi] [--graph]
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
ap.add_argument("--shape", type=int, nargs="+", default=[512, 512, 512])
ap.add_argument("--gl", type=int, default=5)
ap.add_argument("--cycles", type=int, default=2)
ap.add_argument("--smoother", default="jacobi")
ap.add_argument("--pre", type=int, default=1)
ap.add_argument("--post", type=int, default=1)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--profile", action="store_true")
a = ap.parse_args()
shape = tuple(a.shape)
A = omg.operators.poisson_band(shape, sparse_1d=(len(shape) == 1))
h = Hierarchy(A, shape, a.gl - 1, 8, flags=0 if a.graph else _lib.FLAG_NO_GRAPH)
print("levels:", [(h.level_info(l)["n"], h.level_info(l)["kind"], h.level_info(l)["nexc"]) for l in range(h.nlevels)])
print("setup:", h.setup_times())
u = np.random.RandomState(0).random_sample(A.n)
h.set_rhs(h.matvec(u, 0))
ms, launches = h.bench_cycles(a.cycles, a.pre, a.post, a.smoother, 0.8)
print("%d cycles: %.3f ms/cycle, %d launches, norm %.3e" % (a.cycles, ms / a.cycles, launches, h.current_norm()))
if a.profile:
    for p in h.profile_cycle(3, a.pre, a.post, a.smoother, 0.8):
        print("  %-20s L%d  x%-3d %8.4f ms  %8.1f GB/s" % (p["name"], p["level"], p["launches"] // 3, p["ms"],
                                                       p["bytes"] / p["ms"] / 1e6 if p["ms"] > 0 else 0))
