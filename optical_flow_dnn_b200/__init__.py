"""Importable alias of the ``optical-flow-dnn_b200/`` package directory (a hyphen is not a valid
Python identifier, so the real package lives next to this stub and is exposed under this name)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "optical-flow-dnn_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
