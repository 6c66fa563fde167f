"""Import alias: the package directory `sp-gan-tip2025_b200/` is not a valid Python identifier, so this module
makes it importable as `spgan_b200` (it becomes a package by carrying `__path__`)."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "sp-gan-tip2025_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
