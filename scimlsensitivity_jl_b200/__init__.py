"""Importable alias of the package directory `scimlsensitivity.jl_b200/` (a dot cannot appear in a module name)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "scimlsensitivity.jl_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
