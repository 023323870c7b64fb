"""Importable alias of the package directory `scalable-meta-learning-with-gaussian-processes_b200/`.

The mandated directory name contains hyphens and cannot appear in an `import` statement;
this shim points the package search path at it, so `import scamlgp_b200` (and
`scamlgp_b200.model`, `.optimizer`, `.utils`, ...) resolve to the modules that live there.
"""
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
_PKG = _os.path.join(_ROOT, "scalable-meta-learning-with-gaussian-processes_b200")
__path__.insert(0, _PKG)

with open(_os.path.join(_PKG, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG, "__init__.py"), "exec"))
