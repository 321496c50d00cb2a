"""Load the UNMODIFIED reference (/root/reference/openmg, Python 2) in this
Python 3 interpreter by transliterating it IN MEMORY.

Nothing is written to disk: the reference sources are read where they lie,
~20 mechanical py2->py3 edits are applied to the text, and
the result is exec'd into synthetic modules `openmg_ref`, `openmg_ref.tools`,
`openmg_ref.operators`, `openmg_ref.solvers`.  Used ONLY in this container by
tools/make_golden.py (to generate tests/golden/*) and by the optional
cross-check tests that skip when /root/reference is absent (it does not exist
on the GPU box).
"""
import os
import re
import sys
import types

REF_ROOT = os.environ.get("OPENMG_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "openmg", "__init__.py"))


def _py3(src, name):
    # print statements -> functions (single-line forms only)
    def fix_print(m):
        indent, pre, body = m.group(1), m.group(2), m.group(3).rstrip()
        end = ""
        if body.endswith(","):
            body = body[:-1]
            end = ", end=' '"
        return "%s%sprint(%s%s)" % (indent, pre, body, end)

    src = re.sub(r"^(\s*)((?:if [^\n:]*: )?)print (.*)$", fix_print, src, flags=re.M)
    src = re.sub(r"^(\s*)print\s*$", r"\1print()", src, flags=re.M)
    if name == "__init__":
        src = src.replace("from solvers import", "from .solvers import")
        src = src.replace("import tools\n", "from . import tools\n")
        src = src.replace("import operators\n", "from . import operators\n")
        src = src.replace("scipy.sparse.base.np.linalg.norm", "np.linalg.norm")
    elif name == "operators":
        src = src.replace("import tools\n", "from . import tools\n")
        src = src.replace("def poisson2D((NX, NY), sparse=False):",
                          "def poisson2D(shape, sparse=False):\n    NX, NY = shape")
        src = src.replace("def poisson3D((NX, NY, NZ), sparse=False):",
                          "def poisson3D(shape, sparse=False):\n    NX, NY, NZ = shape")
        src = src.replace("xrange", "range")
        src = src.replace("n = N / (2 ** alpha)", "n = N // (2 ** alpha)")
        src = src.replace("np.array(problemShape) / (2 ** level)",
                          "np.array(problemShape) // (2 ** level)")
    elif name == "solvers":
        src = src.replace("from tools import", "from .tools import")
    elif name == "tools":
        src = src.replace(".iteritems()", ".items()")
    return src


def load(pkgname="openmg_ref"):
    """Return the transliterated reference package (cached in sys.modules)."""
    if pkgname in sys.modules:
        return sys.modules[pkgname]
    if not available():
        raise ImportError("reference not present at %s" % REF_ROOT)
    pkg = types.ModuleType(pkgname)
    pkg.__path__ = []  # mark as package
    pkg.__package__ = pkgname
    sys.modules[pkgname] = pkg
    try:
        for sub in ("tools", "operators", "solvers"):
            path = os.path.join(REF_ROOT, "openmg", sub + ".py")
            with open(path) as f:
                src = _py3(f.read(), sub)
            mod = types.ModuleType(pkgname + "." + sub)
            mod.__package__ = pkgname
            mod.__file__ = path
            sys.modules[pkgname + "." + sub] = mod
            exec(compile(src, path, "exec"), mod.__dict__)
            setattr(pkg, sub, mod)
        path = os.path.join(REF_ROOT, "openmg", "__init__.py")
        with open(path) as f:
            src = _py3(f.read(), "__init__")
        pkg.__file__ = path
        exec(compile(src, path, "exec"), pkg.__dict__)
    except Exception:
        for k in [k for k in sys.modules if k == pkgname or k.startswith(pkgname + ".")]:
            del sys.modules[k]
        raise
    return pkg


if __name__ == "__main__":
    import numpy as np
    ref = load()
    A = ref.operators.poisson(100, sparse=True)
    u = np.sin(np.linspace(0, 20, 100) / 10.0)
    b = np.asarray(A * u).ravel()
    x, info = ref.mgSolve(A, b, {'problemShape': (100,), 'gridLevels': 3, 'cycles': 10,
                                 'threshold': 1e-2, 'giveInfo': True})
    print("reference shim OK: cycles=%d norm=%g" % (info['cycle'], info['norm']))
