"""Import shim: the package directory is `dynamicsparsearrays.jl_b200/` (named after the reference package); the dot makes
it unimportable by name, so `import dsa_b200` loads it from that directory under this alias."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dynamicsparsearrays.jl_b200")
_spec = importlib.util.spec_from_file_location("dsa_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dsa_b200"] = _mod
_spec.loader.exec_module(_mod)
