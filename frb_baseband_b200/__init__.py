"""Import shim: the product package lives in the directory ``frb-baseband_b200/`` (the name
the project layout asks for), which is not a valid Python identifier.  This package points
its search path there so ``import frb_baseband_b200.process_vdif`` works from the repo root.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "frb-baseband_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
