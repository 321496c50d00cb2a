"""Debug helper: where does the single-pass two-colour sweep differ from the two half-sweep launches?"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openmg_b200 as omg                     # noqa: E402
from openmg_b200 import _lib                  # noqa: E402
from openmg_b200.hierarchy import Hierarchy   # noqa: E402

shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 128, 128)
lvl = int(sys.argv[4]) if len(sys.argv) > 4 else 1
outs = {}
for on in (False, True):
    if on:
        os.environ.pop("OMG_NO_RB3", None)
    else:
        os.environ["OMG_NO_RB3"] = "1"
    h = Hierarchy(omg.operators.poisson_band(shape), shape, 2, 8, flags=_lib.FLAG_NO_GRAPH)
    n = h.n(lvl)
    rs = np.random.RandomState(13)
    x, b = rs.random_sample(n), rs.random_sample(n)
    outs[on] = h.smooth(lvl, b, x, 1, "rbgs")
    h.close()
sh = tuple(s >> lvl for s in shape)
d = np.abs(outs[True] - outs[False]).reshape(sh)
bad = np.argwhere(d > 1e-12)
print("level shape", sh, "mismatches", len(bad), "of", d.size, "max", d.max())
if len(bad):
    for ax, name in enumerate("zyx"):
        vals, cnt = np.unique(bad[:, ax], return_counts=True)
        print(name, "values:", dict(zip(vals.tolist()[:12], cnt.tolist()[:12])), "..." if len(vals) > 12 else "")
    print("parity (x+y+z)&1 of mismatches:", np.unique(bad.sum(1) & 1, return_counts=True))
    print(bad[:10])
