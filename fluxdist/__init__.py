"""Import shim: ``import fluxdist`` loads the package that lives in
``fluxdistributed.jl_b200/`` (a directory name Python cannot import directly
because of the dot).  The shim points ``__path__`` at the real directory and
executes its ``__init__``, so ``fluxdist.parallel`` etc. resolve to
``fluxdistributed.jl_b200/parallel`` and pickling by module name works in
spawned worker processes.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "fluxdistributed.jl_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
