"""Import alias for the package directory ``reflectance-filtering_b200/``.

The package directory carries the hyphenated project name, which Python cannot
import directly.  This stub points ``__path__`` at the real directory and runs
its ``__init__`` so ``import reflectance_filtering_b200`` (and every
``reflectance_filtering_b200.<submodule>``) resolves to the files there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "reflectance-filtering_b200")
__path__ = [_real]
_init = _os.path.join(_real, "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
del _f, _init
