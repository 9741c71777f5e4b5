"""Import shim: `import mcphylo_jl_b200` loads the package whose sources live in the
directory `mcphylo.jl_b200/` (a name Python's import system cannot spell).  The shim only
redirects the package search path; all code is in `mcphylo.jl_b200/`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mcphylo.jl_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _fh
