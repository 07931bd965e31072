"""Importable alias of the package directory `protograd.jl_b200/` (a dotted directory name cannot
be imported directly).  `import protograd_b200 as pg`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "protograd.jl_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
