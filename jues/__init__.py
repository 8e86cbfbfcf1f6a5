"""Import shim: the product package lives in the directory ``jues.jl_b200/`` at the repo root
(the layout the build contract names).  A directory with a dot in its name is not importable
by the default finder, so ``import jues.jl_b200`` is wired up here explicitly."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "jues.jl_b200")
if "jues.jl_b200" not in _sys.modules:
    _spec = _ilu.spec_from_file_location(
        "jues.jl_b200", _os.path.join(_pkg_dir, "__init__.py"),
        submodule_search_locations=[_pkg_dir])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules["jues.jl_b200"] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = _sys.modules["jues.jl_b200"]
