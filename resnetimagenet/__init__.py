"""Legacy package name (reference C1': ``src/ResNetImageNet.jl`` is a byte-identical copy of the module under its
old name).  ``import resnetimagenet`` is ``import fluxdist``."""
import sys as _sys

import fluxdist as _fd

_sys.modules[__name__] = _fd
