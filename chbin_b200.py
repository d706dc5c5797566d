"""Loader: makes the directory `ch-bin_b200/` importable as the package `chbin_b200` (a hyphen cannot appear in
a Python module name).  `import chbin_b200` executes this file, which replaces itself in sys.modules with the
real package."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "ch-bin_b200")
_spec = _ilu.spec_from_file_location(
    "chbin_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = _ilu.module_from_spec(_spec)
_sys.modules["chbin_b200"] = _mod
_spec.loader.exec_module(_mod)
