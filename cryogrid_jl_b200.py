"""Import shim: the package directory is literally named ``cryogrid.jl_b200`` (not a valid Python
identifier), so this module loads it under the importable name ``cryogrid_jl_b200``."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cryogrid.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "cryogrid_jl_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cryogrid_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
