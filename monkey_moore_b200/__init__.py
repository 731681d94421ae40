"""Import shim: the product package lives in ``monkey-moore_b200/`` (a directory name Python cannot
import directly).  ``import monkey_moore_b200`` executes that package under this name."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "monkey-moore_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    _src = _f.read()
__file__ = _os.path.join(_real, "__init__.py")
exec(compile(_src, __file__, "exec"))
