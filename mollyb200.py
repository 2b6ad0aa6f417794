"""Import shim: the package directory is `molly.jl_b200/` (named after the reference), which is not a
valid Python identifier, so it is loaded here under the module name `molly_jl_b200` and re-exported."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "molly.jl_b200")
if "molly_jl_b200" not in _sys.modules:
    _spec = _ilu.spec_from_file_location("molly_jl_b200", _os.path.join(_pkg_dir, "__init__.py"),
                                         submodule_search_locations=[_pkg_dir])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules["molly_jl_b200"] = _mod
    _spec.loader.exec_module(_mod)
_mod = _sys.modules["molly_jl_b200"]
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
capi = _mod._capi
