"""Import shim: the product package lives in `euclidiannormalizingflows.jl_b200/`
(a directory name Python cannot import directly because of the dot).  Importing
`enf_b200` loads that directory as the package `enf_b200`."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg_dir = _os.path.join(_here, "euclidiannormalizingflows.jl_b200")
_spec = _ilu.spec_from_file_location(
    "enf_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["enf_b200"] = _mod
_spec.loader.exec_module(_mod)
