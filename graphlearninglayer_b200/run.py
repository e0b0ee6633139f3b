"""Launcher for the reference scripts that carry a PASTED copy of the layer instead of importing GLL
(train_and_adversarial.py:26-263, adversarial.py:28-263).

    python -m graphlearninglayer_b200.run train_and_adversarial.py gl natural mnist

The script is executed unchanged except that its module-level definitions of `LaplaceLearningSparseHard`,
`knn_sym_dist` and `stable_conjgrad` are dropped (AST filter) and the three names are bound to this package's
implementations before anything else runs.  Imports of packages that only the pasted code or plotting needs and that are
not installed (`graphlearning`, `umap`, `matplotlib`) are satisfied with empty stand-in modules, so the script's own
training / attack loops run as written on the B200 path.
"""
from __future__ import annotations

import ast
import importlib
import importlib.util
import os
import sys
import types

REPLACED = ("LaplaceLearningSparseHard", "knn_sym_dist", "stable_conjgrad")
OPTIONAL_MODULES = ("graphlearning", "umap", "matplotlib", "matplotlib.pyplot")


def rewrite(source: str, filename: str = "<script>"):
    """Returns (code object, list of dropped definitions)."""
    tree = ast.parse(source, filename)
    dropped, body = [], []
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in REPLACED:
            dropped.append(f"{type(node).__name__} {node.name} (line {node.lineno})")
            continue
        body.append(node)
    inject = ast.parse("from graphlearninglayer_b200 import LaplaceLearningSparseHard, knn_sym_dist, stable_conjgrad").body
    # keep a module docstring / __future__ imports first
    k = 0
    while k < len(body) and ((isinstance(body[k], ast.Expr) and isinstance(getattr(body[k], "value", None), ast.Constant)
                              and isinstance(body[k].value.value, str))
                             or (isinstance(body[k], ast.ImportFrom) and body[k].module == "__future__")):
        k += 1
    tree.body = body[:k] + inject + body[k:]
    ast.fix_missing_locations(tree)
    return compile(tree, filename, "exec"), dropped


def stub_missing_modules(names=OPTIONAL_MODULES):
    made = []
    for nm in names:
        if nm in sys.modules:
            continue
        try:
            if importlib.util.find_spec(nm) is not None:
                continue
        except (ImportError, ValueError):
            pass
        mod = types.ModuleType(nm)
        mod.__dict__["__getattr__"] = lambda attr, _nm=nm: (_ for _ in ()).throw(
            AttributeError(f"'{_nm}' is a stand-in created by graphlearninglayer_b200.run; '{attr}' is not available"))
        sys.modules[nm] = mod
        if "." in nm:
            parent, child = nm.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)
        made.append(nm)
    return made


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = argv[0]
    with open(script, "r") as fh:
        src = fh.read()
    code, dropped = rewrite(src, script)
    stubs = stub_missing_modules()
    print(f"[gll-b200] {script}: replaced {dropped or 'nothing (the script imports GLL: use the drop-in module instead)'}"
          + (f"; stand-in modules: {stubs}" if stubs else ""), file=sys.stderr)
    sys.argv = argv
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    glb = {"__name__": "__main__", "__file__": os.path.abspath(script), "__builtins__": __builtins__}
    exec(code, glb)
    return 0


if __name__ == "__main__":
    sys.exit(main())
