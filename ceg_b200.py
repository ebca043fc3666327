"""Import shim: `import ceg_b200` loads the package directory `crystalenergygrids.jl_b200/`
(its name contains a dot, so it cannot be imported by name)."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "crystalenergygrids.jl_b200")
_spec = _u.spec_from_file_location("ceg_b200", _os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = _u.module_from_spec(_spec)
_sys.modules["ceg_b200"] = _mod
_spec.loader.exec_module(_mod)
