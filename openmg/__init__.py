"""`import openmg` drop-in alias of the B200 implementation (openmg_b200).

User code written against tsbertalan/openmg keeps working unchanged:
`openmg.mgSolve`, `openmg.mgCycle`, `openmg.defaults`, `openmg.smooth`,
`openmg.operators.*`, `openmg.solvers.*`, `openmg.tools.*`.
Attribute assignment (e.g. the smoother plug-in point `openmg.smooth = f`) is
forwarded to the implementing module.
"""
import sys

import openmg_b200 as _impl

sys.modules[__name__ + ".operators"] = _impl.operators
sys.modules[__name__ + ".solvers"] = _impl.solvers
sys.modules[__name__ + ".tools"] = _impl.tools
sys.modules[__name__] = _impl
